#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary7.txt
: > $S
run() { local name=$1; local to=$2; shift 2
  timeout "$to" python -m pytest "$@" -q --timeout 600 -p no:cacheprovider > "gpurun_out/t7_${name}.log" 2>&1
  echo "$name exit=$?" | tee -a $S; tail -n 3 "gpurun_out/t7_${name}.log" | tee -a $S; }
run gemm_tc  600 tests/test_gpu_kernels.py -m gpu -k "test_gemm and tcgen05"
for ord in 1 0; do
  VITED_GEMM_ORDER=$ord timeout 600 python tools/bench_ops.py > gpurun_out/bench_ops_v5_ord$ord.jsonl 2> gpurun_out/bench_ops_v5_ord$ord.err
  echo "order=$ord" | tee -a $S
  grep gemm_ gpurun_out/bench_ops_v5_ord$ord.jsonl | cut -c1-150 | tee -a $S
done
for bn in 128 256; do
  VITED_GEMM_BN=$bn timeout 600 python tools/bench_ops.py > gpurun_out/bench_ops_v5_bn$bn.jsonl 2> gpurun_out/bench_ops_v5_bn$bn.err
  echo "order=1 bn=$bn" | tee -a $S
  grep gemm_ gpurun_out/bench_ops_v5_bn$bn.jsonl | cut -c1-150 | tee -a $S
done
