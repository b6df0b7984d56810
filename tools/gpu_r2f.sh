#!/usr/bin/env bash
# 2-GPU pass: NCCL test of the package's sharded entry points, bench at N=2 (puzzle through grid.score_puzzles, Hisfrag
# through grid.score_fragments), plus the full 1-GPU test suite on the current code
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02f_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "gpu suite rc=$?"
tail -4 gpurun_out/r02f_pytest_gpu.log
grep -E "mAP|hisfrag20 model|max \|logit" gpurun_out/r02f_pytest_gpu.log | head -20
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
VITED_BENCH_UNITS=${UNITS:-1000} timeout 900 $RUN bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err; echo "bench n2 rc=$?"
cut -c1-700 gpurun_out/r02f_bench_n2.json
timeout 900 $RUN bench.py --gpus 2 --workload hisfrag --items 256 --steps 1 --warmup 1 > gpurun_out/r02f_bench_hisfrag256_n2.json 2> gpurun_out/r02f_bench_hisfrag256_n2.err; echo "hisfrag n2 rc=$?"
cut -c1-1500 gpurun_out/r02f_bench_hisfrag256_n2.json
tail -5 gpurun_out/r02f_bench_hisfrag256_n2.err
