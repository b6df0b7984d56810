#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary9.txt
: > $S
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "cta_pair" --timeout 800 -p no:cacheprovider > gpurun_out/t9_pair.log 2>&1
echo "pair exit=$?" | tee -a $S; tail -n 30 gpurun_out/t9_pair.log | cut -c1-300 | tee -a $S
VITED_GEMM_PAIR=1 timeout 600 python tools/bench_ops.py > gpurun_out/bench_ops_pair.jsonl 2> gpurun_out/bench_ops_pair.err
echo "bench_ops pair exit=$?" | tee -a $S
grep gemm_ gpurun_out/bench_ops_pair.jsonl | cut -c1-150 | tee -a $S
tail -5 gpurun_out/bench_ops_pair.err | tee -a $S
