#!/usr/bin/env bash
# Round-2 (second session) experiment B: class-token query rows of the long-sequence attention on CUDA-core warps.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" -p no:cacheprovider > gpurun_out/r3b_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -3 gpurun_out/r3b_kernels.log
timeout 300 python tools/profile_attn_l64.py time > gpurun_out/r3b_l64.jsonl 2> gpurun_out/r3b_l64.err; echo "l64 time rc=$?"; cat gpurun_out/r3b_l64.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_retrieval.py -q -x -p no:cacheprovider > gpurun_out/r3b_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/r3b_parity.log
for c in 1 0; do
  VITED_L64_CLS_WARPS=$c timeout 600 python bench.py --workload hisfrag --items 128 --steps 1 --warmup 1 > gpurun_out/r3b_hisfrag128_clsw$c.json 2> gpurun_out/r3b_hisfrag128_clsw$c.err; echo "hisfrag clsw=$c rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r3b_hisfrag128_clsw$c.json'))
c=d.get('classes_rank0',{})
print('clsw=$c', round(d['value'],1), d['clocks']['sm_mhz'], {k:round(v['ms'],1) for k,v in c.items() if v['share']>0.02})
PY
done
