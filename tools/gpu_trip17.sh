#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TRIP:-17}
S=gpurun_out/summary$T.txt
: > $S
run() { local name=$1; local to=$2; shift 2
  timeout "$to" python -m pytest "$@" -q --timeout 800 -p no:cacheprovider > "gpurun_out/t${T}_${name}.log" 2>&1
  echo "$name exit=$?" | tee -a $S; tail -n 12 "gpurun_out/t${T}_${name}.log" | cut -c1-400 | tee -a $S; }
run gemmln 300 tests/test_gpu_kernels.py -m gpu -k "gemm_resid_ln"
timeout 600 python tools/bench_ops.py > gpurun_out/bench_ops_v$T.jsonl 2> gpurun_out/bench_ops_v$T.err
grep "gemm_ln\|resid_ln\|gemm_proj\|gemm_fc2" gpurun_out/bench_ops_v$T.jsonl | cut -c1-200 | tee -a $S
run parity 900 tests/test_gpu_parity.py -m gpu
timeout 900 python bench.py --no-cpu > gpurun_out/bench_n1_v$T.json 2> gpurun_out/bench_n1_v$T.err; echo "bench n1 exit=$?" | tee -a $S
cut -c1-300 gpurun_out/bench_n1_v$T.json | tee -a $S
