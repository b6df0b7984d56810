#!/usr/bin/env bash
# round-2 first GPU pass: new fused MLP kernel alone, then the whole GPU suite, per-op timings, the bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "mlp_resid_ln or gemm_resid_ln" > gpurun_out/r02a_pytest_mlp.log 2>&1
echo "mlp kernel tests rc=$?" | tee -a gpurun_out/r02a_status.txt
timeout 300 python tools/bench_ops.py > gpurun_out/r02a_bench_ops.jsonl 2> gpurun_out/r02a_bench_ops.err
echo "bench_ops rc=$?" | tee -a gpurun_out/r02a_status.txt
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_retrieval.py::test_retrieval_top1_and_map_identical_to_three_decimals -s > gpurun_out/r02a_pytest_gpu.log 2>&1
echo "gpu suite rc=$?" | tee -a gpurun_out/r02a_status.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err
echo "bench rc=$?" | tee -a gpurun_out/r02a_status.txt
VITED_FUSE_MLP=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/r02a_bench_n1_nomlp.json 2> gpurun_out/r02a_bench_n1_nomlp.err
echo "bench (unfused MLP) rc=$?" | tee -a gpurun_out/r02a_status.txt
tail -3 gpurun_out/r02a_pytest_mlp.log; tail -5 gpurun_out/r02a_pytest_gpu.log; cat gpurun_out/r02a_bench_ops.jsonl | cut -c1-220; cut -c1-600 gpurun_out/r02a_bench_n1.json; cut -c1-300 gpurun_out/r02a_bench_n1_nomlp.json
