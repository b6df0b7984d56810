#!/usr/bin/env bash
# puzzle attention: the MMA issuer serves whichever group is ready (A = tools/bin/inorder: strict unit order)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" -p no:cacheprovider 2>&1 | tail -2
VITED_LIB=$PWD/tools/bin/jitter/libvited_b200.so timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" -p no:cacheprovider 2>&1 | tail -2
for rep in 1 2; do
for v in inorder ooo; do
  if [ $v = ooo ]; then unset VITED_LIB; else export VITED_LIB=$PWD/tools/bin/inorder/libvited_b200.so; fi
  OPS=attn timeout 200 python tools/bench_ops.py > gpurun_out/r3m_ops_${v}_$rep.jsonl 2> gpurun_out/r3m_ops_${v}_$rep.err
  python - <<PY
import json
r={}
for l in open('gpurun_out/r3m_ops_${v}_$rep.jsonl'):
    d=json.loads(l); r[d['op']]=round(d['ms'],4)
print('[$v, run $rep]', {k:v for k,v in r.items() if 'impl0' in k})
PY
done
done
for rep in 1 2; do
for v in inorder ooo; do
  if [ $v = ooo ]; then unset VITED_LIB; else export VITED_LIB=$PWD/tools/bin/inorder/libvited_b200.so; fi
  timeout 600 python bench.py --no-cpu --no-extras > gpurun_out/r3m_bench_${v}_$rep.json 2> gpurun_out/r3m_bench_${v}_$rep.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r3m_bench_${v}_$rep.json'))
    c=d['roofline'].get('classes',{})
    print('[$v] puzzle', round(d['value']), d['clocks']['sm_mhz'], {k:round(x['ms'],1) for k,x in c.items() if 'attn' in k})
except Exception as ex:
    print('[$v] no bench line', ex)
PY
done
done
