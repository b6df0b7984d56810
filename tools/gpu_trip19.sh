#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TRIP:-19}
S=gpurun_out/summary$T.txt
: > $S
run() { local name=$1; local to=$2; shift 2
  timeout "$to" python -m pytest "$@" -q --timeout 600 -p no:cacheprovider -x > "gpurun_out/t${T}_${name}.log" 2>&1
  echo "$name exit=$?" | tee -a $S; tail -n 14 "gpurun_out/t${T}_${name}.log" | cut -c1-300 | tee -a $S; }
run attn 600 tests/test_gpu_kernels.py -m gpu -k "attention and fast"
python tools/profile_attn_l64.py time 2>&1 | tee -a $S
