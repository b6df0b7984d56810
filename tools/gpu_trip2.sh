#!/usr/bin/env bash
# smoke + bench (N=1) + ncu launch list + one full capture of the dominant kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/summary2.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" | tee -a gpurun_out/summary2.txt
tail -n 2 gpurun_out/smoke.log | tee -a gpurun_out/summary2.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "errors_are_loud" -p no:cacheprovider > gpurun_out/t_err.log 2>&1; echo "errtest exit=$?" | tee -a gpurun_out/summary2.txt
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?" | tee -a gpurun_out/summary2.txt
cat gpurun_out/bench_n1.json | tee -a gpurun_out/summary2.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref exit=$?" | tee -a gpurun_out/summary2.txt
cat gpurun_out/bench_ref.json | tee -a gpurun_out/summary2.txt
CMD="python bench.py --steps 1 --warmup 1 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu launches exit=$?" | tee -a gpurun_out/summary2.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 300 -c 4 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu full exit=$?" | tee -a gpurun_out/summary2.txt
ls -la gpurun_out | tee -a gpurun_out/summary2.txt
