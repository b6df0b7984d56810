#!/usr/bin/env bash
# 2-GPU box: attention v5 tests, single-GPU bench, then the N=2 bench path (NCCL all-gather) on a reduced unit count
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary8.txt
: > $S
run() { local name=$1; local to=$2; shift 2
  timeout "$to" python -m pytest "$@" -q --timeout 600 -p no:cacheprovider > "gpurun_out/t8_${name}.log" 2>&1
  echo "$name exit=$?" | tee -a $S; tail -n 3 "gpurun_out/t8_${name}.log" | tee -a $S; }
run attn_mma 600 tests/test_gpu_kernels.py -m gpu -k "attention and mma"
run parity   1500 tests/test_gpu_parity.py -m gpu
timeout 600 python tools/bench_ops.py > gpurun_out/bench_ops_v5.jsonl 2> gpurun_out/bench_ops_v5.err
grep attn gpurun_out/bench_ops_v5.jsonl | cut -c1-150 | tee -a $S
timeout 900 python bench.py --no-cpu > gpurun_out/bench_n1_v5.json 2> gpurun_out/bench_n1_v5.err; echo "bench n1 exit=$?" | tee -a $S
cut -c1-600 gpurun_out/bench_n1_v5.json | tee -a $S
VITED_BENCH_UNITS=600 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 exit=$?" | tee -a $S
tail -c 1500 gpurun_out/bench_n2.json | tee -a $S; tail -n 5 gpurun_out/bench_n2.err | tee -a $S
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
echo "bench ref n2 exit=$?" | tee -a $S; cut -c1-300 gpurun_out/bench_ref_n2.json | tee -a $S
timeout 900 python bench.py --workload hisfrag --items 32 --steps 1 --warmup 1 > gpurun_out/bench_hisfrag.json 2> gpurun_out/bench_hisfrag.err; echo "hisfrag exit=$?" | tee -a $S
cat gpurun_out/bench_hisfrag.json | tee -a $S
