#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -s > gpurun_out/r02_pytest_train.log 2>&1; echo "train tests rc=$?"
tail -40 gpurun_out/r02_pytest_train.log
