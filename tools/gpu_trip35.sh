#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TRIP:-35}
S=gpurun_out/summary$T.txt
: > $S
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -k "gemm_resid_ln" -q --timeout 600 -p no:cacheprovider > gpurun_out/t${T}_ln.log 2>&1; echo "gemm_ln exit=$?" | tee -a $S
tail -n 3 gpurun_out/t${T}_ln.log | cut -c1-300 | tee -a $S
timeout 600 python tools/bench_ops.py 2>/dev/null | grep "gemm_ln" | cut -c1-140 | tee -a $S
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/t${T}_parity.log 2>&1; echo "parity exit=$?" | tee -a $S
tail -n 3 gpurun_out/t${T}_parity.log | cut -c1-300 | tee -a $S
timeout 900 python bench.py --no-cpu > gpurun_out/bench_n1_v$T.json 2> gpurun_out/bench_n1_v$T.err; echo "bench n1 exit=$?" | tee -a $S
cut -c1-200 gpurun_out/bench_n1_v$T.json | tee -a $S
