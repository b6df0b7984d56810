#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TRIP:-28}
S=gpurun_out/summary$T.txt
: > $S
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -k "attention and fast" -q --timeout 600 -p no:cacheprovider -x > gpurun_out/t${T}_attn.log 2>&1; echo "attn exit=$?" | tee -a $S
tail -n 3 gpurun_out/t${T}_attn.log | cut -c1-200 | tee -a $S
timeout 600 python tools/bench_ops.py 2>/dev/null | grep attn | cut -c1-130 | tee -a $S
