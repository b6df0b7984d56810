#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TRIP:-18}
S=gpurun_out/summary$T.txt
: > $S
run() { local name=$1; local to=$2; shift 2
  timeout "$to" python -m pytest "$@" -q --timeout 600 -p no:cacheprovider -x > "gpurun_out/t${T}_${name}.log" 2>&1
  echo "$name exit=$?" | tee -a $S; tail -n 14 "gpurun_out/t${T}_${name}.log" | cut -c1-300 | tee -a $S; }
run attn 600 tests/test_gpu_kernels.py -m gpu -k "attention and fast"
run parity 900 tests/test_gpu_parity.py -m gpu
timeout 600 python bench.py --workload hisfrag --items 32 --steps 1 --warmup 1 > gpurun_out/bench_hisfrag_v$T.json 2> gpurun_out/bench_hisfrag_v$T.err; echo "hisfrag exit=$?" | tee -a $S
cut -c1-900 gpurun_out/bench_hisfrag_v$T.json | tee -a $S
