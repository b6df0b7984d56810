#!/usr/bin/env bash
# Round-2 (second session) checkpoint: full GPU suite, default bench line, per-kernel timings.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r3f_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/r3f_pytest_gpu.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r3f_bench_n1.json 2> gpurun_out/r3f_bench_n1.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r3f_bench_n1.json
timeout 600 python tools/bench_ops.py > gpurun_out/r3f_bench_ops.jsonl 2> gpurun_out/r3f_bench_ops.err; echo "ops rc=$?"; cut -c1-150 gpurun_out/r3f_bench_ops.jsonl
timeout 300 python tools/profile_attn_l64.py time >> gpurun_out/r3f_bench_ops.jsonl 2>> gpurun_out/r3f_bench_ops.err; tail -4 gpurun_out/r3f_bench_ops.jsonl | cut -c1-150
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
