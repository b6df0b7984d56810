#!/usr/bin/env bash
# Round 2, third session, second (last) call: launch-shape knobs re-measured on the final kernels -- rows per chunk of
# pairs (the 524,288 default was chosen before the MLP fusion) and programmatic dependent launch, alternating with the
# default on one box; Hisfrag20 model with two chunk sizes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
bench() {  # $1 = tag; env decides the variant
  timeout 120 python bench.py --no-cpu --no-extras > gpurun_out/r02e_bench_$1.json 2> gpurun_out/r02e_bench_$1.err; local rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02e_bench_$1.json'))
    print('[$1] rc=$rc value', round(d['value']), 'e2e', round(d['e2e']['value']), 'MHz', d['clocks']['sm_mhz'], 'launches', d['gpu_launches'])
except Exception as ex:
    print('[$1] rc=$rc no bench line', ex)
PY
  el done
}
bench base_1
VITED_CHUNK_ROWS=786432 bench chunk768k
VITED_CHUNK_ROWS=1048576 bench chunk1m
VITED_CHUNK_ROWS=393216 bench chunk384k
VITED_PDL=1 bench pdl
bench base_2
hf() {
  timeout 150 python bench.py --workload hisfrag --items 128 --steps 1 --warmup 1 > gpurun_out/r02e_hisfrag128_$1.json 2> gpurun_out/r02e_hisfrag128_$1.err; local rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02e_hisfrag128_$1.json'))
    print('[hisfrag $1] rc=$rc value', round(d['value']), 'MHz', d['clocks']['sm_mhz'])
except Exception as ex:
    print('[hisfrag $1] rc=$rc no bench line', ex)
PY
  el done
}
hf base
VITED_CHUNK_ROWS=1048576 hf chunk1m
