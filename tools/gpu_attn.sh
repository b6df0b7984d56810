#!/usr/bin/env bash
# attention kernels: tests + isolated timings (puzzle and Hisfrag shapes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -k "attention and fast" -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -3
timeout 300 python tools/profile_attn_l64.py time 2>&1 | grep impl0
VITED_ATTN_L64_CLS=inline timeout 300 python tools/profile_attn_l64.py time 2>&1 | grep impl0 | sed 's/^/inline-cls: /'
