#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary11.txt
: > $S
run() { local name=$1; local to=$2; shift 2
  timeout "$to" python -m pytest "$@" -q --timeout 800 -p no:cacheprovider > "gpurun_out/t11_${name}.log" 2>&1
  echo "$name exit=$?" | tee -a $S; tail -n 6 "gpurun_out/t11_${name}.log" | cut -c1-300 | tee -a $S; }
run attn 900 tests/test_gpu_kernels.py -m gpu -k "attention and mma"
run parity  1500 tests/test_gpu_parity.py -m gpu
timeout 600 python tools/bench_ops.py > gpurun_out/bench_ops_v7.jsonl 2> gpurun_out/bench_ops_v7.err
grep attn gpurun_out/bench_ops_v7.jsonl | cut -c1-150 | tee -a $S
timeout 900 python bench.py --no-cpu > gpurun_out/bench_n1_v7.json 2> gpurun_out/bench_n1_v7.err; echo "bench n1 exit=$?" | tee -a $S
cut -c1-300 gpurun_out/bench_n1_v7.json | tee -a $S
timeout 900 python bench.py --workload hisfrag --items 32 --steps 1 --warmup 1 > gpurun_out/bench_hisfrag_v7.json 2> gpurun_out/bench_hisfrag_v7.err; echo "hisfrag exit=$?" | tee -a $S
cut -c1-700 gpurun_out/bench_hisfrag_v7.json | tee -a $S
