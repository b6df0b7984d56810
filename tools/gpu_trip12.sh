#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary12.txt
: > $S
run() { local name=$1; local to=$2; shift 2
  timeout "$to" python -m pytest "$@" -q --timeout 800 -p no:cacheprovider > "gpurun_out/t12_${name}.log" 2>&1
  echo "$name exit=$?" | tee -a $S; tail -n 6 "gpurun_out/t12_${name}.log" | cut -c1-300 | tee -a $S; }
run attn 900 tests/test_gpu_kernels.py -m gpu -k "attention and mma"
timeout 600 python tools/bench_ops.py > gpurun_out/bench_ops_v8.jsonl 2> gpurun_out/bench_ops_v8.err
grep attn gpurun_out/bench_ops_v8.jsonl | cut -c1-150 | tee -a $S
timeout 900 python bench.py --no-cpu > gpurun_out/bench_n1_v8.json 2> gpurun_out/bench_n1_v8.err; echo "bench n1 exit=$?" | tee -a $S
cut -c1-300 gpurun_out/bench_n1_v8.json | tee -a $S
