#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TRIP:-34}
S=gpurun_out/summary$T.txt
: > $S
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -k "gemm and not simt or gelu" -q --timeout 600 -p no:cacheprovider > gpurun_out/t${T}_gemm.log 2>&1; echo "gemm exit=$?" | tee -a $S
tail -n 4 gpurun_out/t${T}_gemm.log | cut -c1-300 | tee -a $S
timeout 600 python tools/bench_ops.py 2>/dev/null | grep "fc1\|fc2\|qkv" | cut -c1-120 | tee -a $S
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/t${T}_parity.log 2>&1; echo "parity exit=$?" | tee -a $S
tail -n 3 gpurun_out/t${T}_parity.log | cut -c1-300 | tee -a $S
