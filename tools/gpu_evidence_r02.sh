#!/usr/bin/env bash
# Round-2 evidence (1 GPU): GPU tests, bench lines (ours + reference arm), per-kernel timings, side benches of the
# consumer-side kernels, the training step, ncu launch list of one bench step + full capture of the hot kernels at the
# bench chunk shape (each command first runs clean without ncu).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=${V:-r02}
S=gpurun_out/evidence_$V.txt
: > $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_event_reasons.active --format=csv > gpurun_out/${V}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${V}_pytest_gpu.log 2>&1; echo "pytest gpu exit=$?" | tee -a $S
tail -n 3 gpurun_out/${V}_pytest_gpu.log | tee -a $S
timeout 900 python bench.py > gpurun_out/${V}_bench_n1.json 2> gpurun_out/${V}_bench_n1.err; echo "bench exit=$?" | tee -a $S
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${V}_bench_ref.json 2> gpurun_out/${V}_bench_ref.err; echo "bench ref exit=$?" | tee -a $S
timeout 600 python tools/bench_ops.py > gpurun_out/${V}_bench_ops.jsonl 2> gpurun_out/${V}_bench_ops.err
timeout 300 python tools/profile_attn_l64.py time >> gpurun_out/${V}_bench_ops.jsonl 2>> gpurun_out/${V}_bench_ops.err
timeout 600 python bench.py --workload train --steps 2 --warmup 1 > gpurun_out/${V}_bench_train_n1.json 2> gpurun_out/${V}_bench_train_n1.err; echo "train exit=$?" | tee -a $S
{ timeout 300 python tools/bench_tables.py; timeout 300 python tools/bench_prep.py; timeout 300 python tests/analysis/bench_metrics.py; } > gpurun_out/${V}_bench_consumers.txt 2>&1; echo "consumer benches exit=$?" | tee -a $S
# ncu: launch list of one bench step, then full captures of the hot kernels at the bench chunk shape
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu --no-extras > gpurun_out/${V}_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/${V}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-extras > gpurun_out/${V}_ncu1.log 2>&1
echo "ncu launches exit=$?" | tee -a $S
HOT=1 REPS=1 ROWS=511355 python tools/profile_ops.py > gpurun_out/${V}_plain2.log 2>&1 && \
HOT=1 REPS=1 ROWS=511355 MANIFEST=gpurun_out/${V}_manifest.json timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_pair|gemm_ln_pair|mlp_ln_pair|attn_p64' -c 8 -o gpurun_out/${V}_prof_ops python tools/profile_ops.py > gpurun_out/${V}_ncu2.log 2>&1
echo "ncu ops exit=$?" | tee -a $S
python tools/profile_attn_l64.py > gpurun_out/${V}_plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_l64 -c 2 -o gpurun_out/${V}_prof_l64 python tools/profile_attn_l64.py > gpurun_out/${V}_ncu3.log 2>&1
echo "ncu l64 exit=$?" | tee -a $S
cut -c1-400 gpurun_out/${V}_bench_n1.json | tee -a $S
cat gpurun_out/${V}_bench_consumers.txt | tee -a $S
