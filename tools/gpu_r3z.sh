#!/usr/bin/env bash
# Final check of the second session's code: full GPU suite, the kernel + parity suites under the timing fuzzer and with
# bf16 operands, smoke(), the default bench line and the Hisfrag grid on one GPU (512 fragments).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -2 gpurun_out/r02c_pytest_gpu.log | cut -c1-200
VITED_LIB=$PWD/tools/bin/jitter/libvited_b200.so timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_train.py -q -x -p no:cacheprovider > gpurun_out/r02c_pytest_jitter.log 2>&1; echo "jitter rc=$?"; tail -2 gpurun_out/r02c_pytest_jitter.log | cut -c1-200
VITED_LIB=$PWD/tools/bin/bf16/libvited_b200.so timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_train.py -q -p no:cacheprovider > gpurun_out/r02c_pytest_bf16.log 2>&1; echo "bf16 rc=$?"; tail -3 gpurun_out/r02c_pytest_bf16.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/r02c_bench_n1.json
timeout 600 python bench.py --workload hisfrag --items 512 --steps 1 --warmup 1 > gpurun_out/r02c_bench_hisfrag512_n1.json 2> gpurun_out/r02c_bench_hisfrag512_n1.err; echo "hisfrag rc=$?"; cut -c1-330 gpurun_out/r02c_bench_hisfrag512_n1.json
