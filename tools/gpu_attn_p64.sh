#!/usr/bin/env bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -k "attention and fast" -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -3
timeout 600 python tools/bench_ops.py 2>/dev/null | grep attn | cut -c1-130
