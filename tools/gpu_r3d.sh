#!/usr/bin/env bash
# Round-2 (second session) experiment D: softmax exponentials -- packed FFMA2 / FADD2 and a share of the exponentials as
# a polynomial on the FMA pipe (variant libraries built by tools/build_softmax_variants.sh).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in scalar packed poly1 poly2 poly3; do
  lib=$PWD/tools/bin/sm_$v/libvited_b200.so
  [ -f $lib ] || { echo "missing $lib"; continue; }
  export VITED_LIB=$lib
  timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" -p no:cacheprovider > gpurun_out/r3d_kernels_$v.log 2>&1; echo "[$v] kernel tests rc=$? $(tail -1 gpurun_out/r3d_kernels_$v.log)"
  OPS=attn timeout 300 python tools/bench_ops.py > gpurun_out/r3d_ops_$v.jsonl 2> gpurun_out/r3d_ops_$v.err
  timeout 300 python tools/profile_attn_l64.py time >> gpurun_out/r3d_ops_$v.jsonl 2>> gpurun_out/r3d_ops_$v.err
  python - <<PY
import json
r={}
for l in open('gpurun_out/r3d_ops_$v.jsonl'):
    d=json.loads(l); r[d['op']]=round(d['ms'],4)
print('[$v]', {k:r[k] for k in r if 'impl0' in k})
PY
done
for rep in 1 2; do
for v in scalar packed poly1 poly2 poly3; do
  export VITED_LIB=$PWD/tools/bin/sm_$v/libvited_b200.so
  timeout 600 python bench.py --no-cpu --no-extras > gpurun_out/r3d_bench_${v}_$rep.json 2> gpurun_out/r3d_bench_${v}_$rep.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r3d_bench_${v}_$rep.json'))
    c=d['roofline'].get('classes',{})
    print('[$v] puzzle', round(d['value']), d['clocks']['sm_mhz'], {k:round(x['ms'],1) for k,x in c.items() if 'attn' in k})
except Exception as ex:
    print('[$v] no bench line', ex)
PY
done
done
for v in scalar packed poly1 poly2 poly3; do
  export VITED_LIB=$PWD/tools/bin/sm_$v/libvited_b200.so
  timeout 600 python bench.py --workload hisfrag --items 128 --steps 1 --warmup 1 > gpurun_out/r3d_hisfrag_$v.json 2> gpurun_out/r3d_hisfrag_$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r3d_hisfrag_$v.json'))
    c=d.get('classes_rank0',{})
    print('[$v] hisfrag128', round(d['value'],1), d['clocks']['sm_mhz'], {k:round(x['ms'],1) for k,x in c.items() if 'attn' in k})
except Exception as ex:
    print('[$v] no hisfrag line', ex)
PY
done
