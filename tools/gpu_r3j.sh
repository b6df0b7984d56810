#!/usr/bin/env bash
# warp-uniform lazy rescale in the long-sequence attention: the new rescale test, the attention tests under the timing
# fuzzer, parity + retrieval, kernel timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" -p no:cacheprovider > gpurun_out/r3j_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -4 gpurun_out/r3j_kernels.log | cut -c1-250
VITED_LIB=$PWD/tools/bin/jitter/libvited_b200.so timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" -p no:cacheprovider > gpurun_out/r3j_jitter.log 2>&1; echo "jitter rc=$?"; tail -2 gpurun_out/r3j_jitter.log | cut -c1-250
timeout 300 python tools/profile_attn_l64.py time > gpurun_out/r3j_l64.jsonl 2> gpurun_out/r3j_l64.err; echo "l64 rc=$?"; cut -c1-140 gpurun_out/r3j_l64.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_retrieval.py -q -x -p no:cacheprovider > gpurun_out/r3j_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/r3j_parity.log
