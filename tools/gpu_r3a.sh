#!/usr/bin/env bash
# Round-2 (second session) experiment A: 16 epilogue warps in the two full-row kernels (VITED_EPI_WARPS=16) against 8.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "resid_ln" -p no:cacheprovider > gpurun_out/r3a_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -3 gpurun_out/r3a_kernels.log
OPS=fused timeout 300 python tools/bench_ops.py > gpurun_out/r3a_ops.jsonl 2> gpurun_out/r3a_ops.err; echo "ops rc=$?"; cut -c1-200 gpurun_out/r3a_ops.jsonl
VITED_EPI_WARPS=16 timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider > gpurun_out/r3a_parity16.log 2>&1; echo "parity(16) rc=$?"; tail -2 gpurun_out/r3a_parity16.log
for rep in 1 2; do
  for e in 8 16; do
    VITED_EPI_WARPS=$e timeout 600 python bench.py --no-cpu --no-extras > gpurun_out/r3a_bench_e${e}_$rep.json 2> gpurun_out/r3a_bench_e${e}_$rep.err; echo "bench e=$e rep=$rep rc=$?"
    python - <<PY
import json
d=json.load(open('gpurun_out/r3a_bench_e${e}_$rep.json'))
c=d['roofline'].get('classes',{})
print('e=$e', round(d['value']), d['clocks']['sm_mhz'], {k:round(v['ms'],1) for k,v in c.items() if v['share']>0.04})
PY
  done
done
