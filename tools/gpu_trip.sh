#!/usr/bin/env bash
# One GPU-box visit: every test group in its own process (a trapped kernel poisons only its group), then op timings.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() {  # name, timeout, pytest args...
  local name=$1; local to=$2; shift 2
  timeout "$to" python -m pytest "$@" -q --timeout 600 -p no:cacheprovider > "gpurun_out/t_${name}.log" 2>&1
  echo "$name exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 3 "gpurun_out/t_${name}.log" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run rowops   300 tests/test_gpu_kernels.py -m gpu -k "resid_ln or im2col"
run gemm_simt 600 tests/test_gpu_kernels.py -m gpu -k "test_gemm and simt"
run gemm_tc  600 tests/test_gpu_kernels.py -m gpu -k "test_gemm and tcgen05"
run gemm_bn  600 tests/test_gpu_kernels.py -m gpu -k "tile_shapes"
run attn_simt 600 tests/test_gpu_kernels.py -m gpu -k "attention and simt"
run attn_mma 600 tests/test_gpu_kernels.py -m gpu -k "attention and mma"
run parity_debug 900 tests/test_gpu_parity.py -m gpu -k "ref-ref"
run parity   1500 tests/test_gpu_parity.py -m gpu -k "not ref-ref"
timeout 600 python tools/bench_ops.py > gpurun_out/bench_ops.jsonl 2> gpurun_out/bench_ops.err
echo "bench_ops exit=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_ops.jsonl | tee -a gpurun_out/summary.txt
