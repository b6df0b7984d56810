#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TRIP:-22}
S=gpurun_out/summary$T.txt
: > $S
run() { local name=$1; local to=$2; shift 2
  timeout "$to" python -m pytest "$@" -q --timeout 600 -p no:cacheprovider -x > "gpurun_out/t${T}_${name}.log" 2>&1
  echo "$name exit=$?" | tee -a $S; tail -n 6 "gpurun_out/t${T}_${name}.log" | cut -c1-300 | tee -a $S; }
run gemm 900 tests/test_gpu_kernels.py -m gpu -k "gemm and not simt"
timeout 600 python tools/bench_ops.py > gpurun_out/bench_ops_v$T.jsonl 2> gpurun_out/bench_ops_v$T.err
cut -c1-120 gpurun_out/bench_ops_v$T.jsonl | tee -a $S
