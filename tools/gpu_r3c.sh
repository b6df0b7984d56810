#!/usr/bin/env bash
# Round-2 (second session) experiment C: puzzle attention with two units per 128-row tile (attention_pair.cu).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" -p no:cacheprovider > gpurun_out/r3c_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -15 gpurun_out/r3c_kernels.log | cut -c1-300
OPS=attn timeout 300 python tools/bench_ops.py > gpurun_out/r3c_ops.jsonl 2> gpurun_out/r3c_ops.err; echo "ops rc=$?"; cut -c1-220 gpurun_out/r3c_ops.jsonl; tail -3 gpurun_out/r3c_ops.err
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider > gpurun_out/r3c_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/r3c_parity.log
for rep in 1 2; do
  for e in 1 0; do
    VITED_P64_PAIR=$e timeout 600 python bench.py --no-cpu --no-extras > gpurun_out/r3c_bench_pair${e}_$rep.json 2> gpurun_out/r3c_bench_pair${e}_$rep.err; echo "bench pair=$e rep=$rep rc=$?"
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r3c_bench_pair${e}_$rep.json'))
    c=d['roofline'].get('classes',{})
    print('pair=$e', round(d['value']), d['clocks']['sm_mhz'], {k:round(v['ms'],1) for k,v in c.items() if v['share']>0.04})
except Exception as ex:
    print('no bench line', ex)
PY
  done
done
