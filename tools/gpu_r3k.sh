#!/usr/bin/env bash
# de-synchronising the epilogue warps of the two full-row kernels (-DVITED_EPI_STAGGER=cycles): isolated timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for rep in 1 2; do
for v in 0 300 600 1000; do
  if [ $v = 0 ]; then unset VITED_LIB; else export VITED_LIB=$PWD/tools/bin/stg$v/libvited_b200.so; fi
  OPS=fused timeout 200 python tools/bench_ops.py > gpurun_out/r3k_ops_stg${v}_$rep.jsonl 2> gpurun_out/r3k_ops_stg${v}_$rep.err
  python - <<PY
import json
r={}
for l in open('gpurun_out/r3k_ops_stg${v}_$rep.jsonl'):
    d=json.loads(l); r[d['op']]=round(d['ms'],4)
print('[stagger $v, run $rep]', r)
PY
done
done
