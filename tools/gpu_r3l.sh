#!/usr/bin/env bash
# L2 prefetch of the next h tile in the fused MLP (A = tools/bin/nopf without, B = product library)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "mlp_resid_ln" -p no:cacheprovider 2>&1 | tail -2
for rep in 1 2 3; do
for v in nopf pf; do
  if [ $v = pf ]; then unset VITED_LIB; else export VITED_LIB=$PWD/tools/bin/nopf/libvited_b200.so; fi
  OPS=fused timeout 200 python tools/bench_ops.py > gpurun_out/r3l_ops_${v}_$rep.jsonl 2> gpurun_out/r3l_ops_${v}_$rep.err
  python - <<PY
import json
r={}
for l in open('gpurun_out/r3l_ops_${v}_$rep.jsonl'):
    d=json.loads(l); r[d['op']]=round(d['ms'],4)
print('[$v, run $rep]', r)
PY
done
done
unset VITED_LIB
timeout 120 python tools/trace_mlp_ln.py > gpurun_out/r3l_trace_mlp.txt 2>&1; sed -n 8,13p gpurun_out/r3l_trace_mlp.txt
for rep in 1 2; do
for v in nopf pf; do
  if [ $v = pf ]; then unset VITED_LIB; else export VITED_LIB=$PWD/tools/bin/nopf/libvited_b200.so; fi
  timeout 600 python bench.py --no-cpu --no-extras > gpurun_out/r3l_bench_${v}_$rep.json 2> gpurun_out/r3l_bench_${v}_$rep.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r3l_bench_${v}_$rep.json'))
    c=d['roofline'].get('classes',{})
    print('[$v] puzzle', round(d['value']), d['clocks']['sm_mhz'], {k:round(x['ms'],1) for k,x in c.items() if x['share']>0.2})
except Exception as ex:
    print('[$v] no bench line', ex)
PY
done
done
