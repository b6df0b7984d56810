#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary5.txt
: > $S
run() { local name=$1; local to=$2; shift 2
  timeout "$to" python -m pytest "$@" -q --timeout 600 -p no:cacheprovider > "gpurun_out/t5_${name}.log" 2>&1
  echo "$name exit=$?" | tee -a $S; tail -n 3 "gpurun_out/t5_${name}.log" | tee -a $S; }
run gemm_tc  600 tests/test_gpu_kernels.py -m gpu -k "test_gemm and tcgen05"
run attn_mma 600 tests/test_gpu_kernels.py -m gpu -k "attention and mma"
run parity   1500 tests/test_gpu_parity.py -m gpu
timeout 600 python tools/bench_ops.py > gpurun_out/bench_ops_v4.jsonl 2> gpurun_out/bench_ops_v4.err
echo "bench_ops exit=$?" | tee -a $S
grep -v cublas gpurun_out/bench_ops_v4.jsonl | cut -c1-200 | tee -a $S
VITED_GEMM_RESIDENT=0 timeout 600 python tools/bench_ops.py > gpurun_out/bench_ops_v4_nores.jsonl 2> gpurun_out/bench_ops_v4_nores.err
grep gemm_ gpurun_out/bench_ops_v4_nores.jsonl | cut -c1-200 | tee -a $S
timeout 900 python bench.py --no-cpu > gpurun_out/bench_n1_v4.json 2> gpurun_out/bench_n1_v4.err; echo "bench exit=$?" | tee -a $S
cat gpurun_out/bench_n1_v4.json | tee -a $S
