#!/usr/bin/env bash
# GEMM kernels: tests + isolated timings
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -k "gemm and not simt or gelu" -q --timeout 600 -p no:cacheprovider 2>&1 | tail -3
timeout 600 python tools/bench_ops.py 2>/dev/null | grep "gemm" | cut -c1-125
