#!/usr/bin/env bash
# Round-2 (second session) experiment I: long-sequence attention with two threads per query row (VITED_L64_SPLIT=1 default)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" -p no:cacheprovider > gpurun_out/r3i_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -12 gpurun_out/r3i_kernels.log | cut -c1-250
timeout 300 python tools/profile_attn_l64.py time > gpurun_out/r3i_l64.jsonl 2> gpurun_out/r3i_l64.err; echo "l64 rc=$?"; cut -c1-140 gpurun_out/r3i_l64.jsonl; tail -3 gpurun_out/r3i_l64.err
VITED_LIB=$PWD/tools/bin/jitter/libvited_b200.so timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" -p no:cacheprovider > gpurun_out/r3i_jitter.log 2>&1; echo "jitter rc=$?"; tail -2 gpurun_out/r3i_jitter.log | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_retrieval.py -q -x -p no:cacheprovider > gpurun_out/r3i_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/r3i_parity.log
for rep in 1 2; do
for c in 1 0; do
  VITED_L64_SPLIT=$c timeout 600 python bench.py --workload hisfrag --items 128 --steps 1 --warmup 1 > gpurun_out/r3i_hisfrag128_split${c}_$rep.json 2> gpurun_out/r3i_hisfrag128_split${c}_$rep.err; echo "hisfrag split=$c rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r3i_hisfrag128_split${c}_$rep.json'))
    c=d.get('classes_rank0',{})
    print('split=$c', round(d['value'],1), d['clocks']['sm_mhz'], {k:round(v['ms'],1) for k,v in c.items() if v['share']>0.04})
except Exception as ex:
    print('no line', ex)
PY
done
done
