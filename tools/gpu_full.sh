#!/usr/bin/env bash
# full GPU check: all gpu tests, op bench, puzzle bench (N=1), Hisfrag side bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TRIP:-full}
S=gpurun_out/summary_$T.txt
: > $S
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/t_${T}_all.log 2>&1; echo "pytest gpu exit=$?" | tee -a $S
tail -n 6 gpurun_out/t_${T}_all.log | cut -c1-300 | tee -a $S
timeout 600 python tools/bench_ops.py > gpurun_out/bench_ops_$T.jsonl 2> gpurun_out/bench_ops_$T.err
timeout 900 python bench.py --no-cpu > gpurun_out/bench_n1_$T.json 2> gpurun_out/bench_n1_$T.err; echo "bench n1 exit=$?" | tee -a $S
cut -c1-260 gpurun_out/bench_n1_$T.json | tee -a $S
timeout 600 python bench.py --workload hisfrag --items 32 --steps 1 --warmup 1 > gpurun_out/bench_hisfrag_$T.json 2> gpurun_out/bench_hisfrag_$T.err; echo "hisfrag exit=$?" | tee -a $S
cut -c1-700 gpurun_out/bench_hisfrag_$T.json | tee -a $S
