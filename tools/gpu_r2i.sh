#!/usr/bin/env bash
mkdir -p gpurun_out
VITED_LIB=$PWD/tools/bin/jitter/libvited_b200.so timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_train.py -q -x > gpurun_out/r02i_jitter.log 2>&1; echo "jitter rc=$?"; tail -2 gpurun_out/r02i_jitter.log
VITED_LIB=$PWD/tools/bin/bf16/libvited_b200.so timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_train.py -q > gpurun_out/r02i_bf16.log 2>&1; echo "bf16 rc=$?"; tail -4 gpurun_out/r02i_bf16.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k quad > gpurun_out/r02i_experimental.log 2>&1; echo "experimental rc=$?"; tail -2 gpurun_out/r02i_experimental.log
