#!/usr/bin/env bash
# pass-1 breakdown of the fused MLP epilogue + more polynomial share in the softmax exponentials (4, 6, 8 of 8)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python tools/trace_mlp_ln.py > gpurun_out/r3h_trace_mlp.txt 2>&1; tail -26 gpurun_out/r3h_trace_mlp.txt
for v in 2 4 6 8; do
  export VITED_LIB=$PWD/tools/bin/sm_poly$v/libvited_b200.so
  timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" -p no:cacheprovider > gpurun_out/r3h_kernels_poly$v.log 2>&1; echo "[poly$v] kernel tests rc=$? $(tail -1 gpurun_out/r3h_kernels_poly$v.log)"
  OPS=attn timeout 300 python tools/bench_ops.py > gpurun_out/r3h_ops_poly$v.jsonl 2> gpurun_out/r3h_ops_poly$v.err
  timeout 300 python tools/profile_attn_l64.py time >> gpurun_out/r3h_ops_poly$v.jsonl 2>> gpurun_out/r3h_ops_poly$v.err
  python - <<PY
import json
r={}
for l in open('gpurun_out/r3h_ops_poly$v.jsonl'):
    d=json.loads(l); r[d['op']]=round(d['ms'],4)
print('[poly$v]', {k:r[k] for k in r if 'impl0' in k})
PY
done
