#!/usr/bin/env bash
# Round-2 (second session) experiment E: packed FFMA2 / FADD2 in the GELU of the fused MLP and in both full-row epilogue
# passes. A = tools/bin/sm_poly2 (the previous library: same attention, scalar epilogues), B = the product library.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "resid_ln or gelu or saturate" -p no:cacheprovider > gpurun_out/r3e_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -3 gpurun_out/r3e_kernels.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_retrieval.py -q -x -p no:cacheprovider > gpurun_out/r3e_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/r3e_parity.log
for v in old new; do
  if [ $v = old ]; then export VITED_LIB=$PWD/tools/bin/sm_poly2/libvited_b200.so; else unset VITED_LIB; fi
  OPS=fused timeout 300 python tools/bench_ops.py > gpurun_out/r3e_ops_$v.jsonl 2> gpurun_out/r3e_ops_$v.err; echo "[$v] ops rc=$?"; cut -c1-170 gpurun_out/r3e_ops_$v.jsonl
done
for rep in 1 2; do
for v in old new; do
  if [ $v = old ]; then export VITED_LIB=$PWD/tools/bin/sm_poly2/libvited_b200.so; else unset VITED_LIB; fi
  timeout 600 python bench.py --no-cpu --no-extras > gpurun_out/r3e_bench_${v}_$rep.json 2> gpurun_out/r3e_bench_${v}_$rep.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r3e_bench_${v}_$rep.json'))
    c=d['roofline'].get('classes',{})
    print('[$v] puzzle', round(d['value']), d['clocks']['sm_mhz'], {k:round(x['ms'],1) for k,x in c.items() if x['share']>0.04})
except Exception as ex:
    print('[$v] no bench line', ex)
PY
done
done
for v in old new; do
  if [ $v = old ]; then export VITED_LIB=$PWD/tools/bin/sm_poly2/libvited_b200.so; else unset VITED_LIB; fi
  timeout 600 python bench.py --workload hisfrag --items 128 --steps 1 --warmup 1 > gpurun_out/r3e_hisfrag_$v.json 2> gpurun_out/r3e_hisfrag_$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r3e_hisfrag_$v.json'))
    c=d.get('classes_rank0',{})
    print('[$v] hisfrag128', round(d['value'],1), d['clocks']['sm_mhz'], {k:round(x['ms'],1) for k,x in c.items() if x['share']>0.04})
except Exception as ex:
    print('[$v] no hisfrag line', ex)
PY
done
