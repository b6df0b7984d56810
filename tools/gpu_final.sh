#!/usr/bin/env bash
# final check of the committed state: all GPU tests, smoke, bench (ours + reference arm)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench exit=$?"; cut -c1-330 gpurun_out/final_bench_n1.json
timeout 600 python bench.py --workload hisfrag --items 48 --steps 1 --warmup 1 2>/dev/null | cut -c1-200
