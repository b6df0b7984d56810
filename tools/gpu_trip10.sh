#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary10.txt
: > $S
run() { local name=$1; local to=$2; shift 2
  timeout "$to" python -m pytest "$@" -q --timeout 800 -p no:cacheprovider > "gpurun_out/t10_${name}.log" 2>&1
  echo "$name exit=$?" | tee -a $S; tail -n 3 "gpurun_out/t10_${name}.log" | tee -a $S; }
run kernels 1500 tests/test_gpu_kernels.py -m gpu
run parity  1500 tests/test_gpu_parity.py -m gpu
timeout 600 python tools/bench_ops.py > gpurun_out/bench_ops_v6.jsonl 2> gpurun_out/bench_ops_v6.err
grep -v cublas gpurun_out/bench_ops_v6.jsonl | cut -c1-150 | tee -a $S
timeout 900 python bench.py > gpurun_out/bench_n1_v6.json 2> gpurun_out/bench_n1_v6.err; echo "bench n1 exit=$?" | tee -a $S
cut -c1-400 gpurun_out/bench_n1_v6.json | tee -a $S
CMD="python bench.py --steps 1 --warmup 1 --no-cpu"
$CMD > gpurun_out/plain10.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_v6.csv $CMD > gpurun_out/ncu10a.log 2>&1
echo "ncu launches exit=$?" | tee -a $S
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_pair|attn_mma" -s 400 -c 8 -o gpurun_out/prof_v6 -f $CMD > gpurun_out/ncu10b.log 2>&1
echo "ncu full exit=$?" | tee -a $S
