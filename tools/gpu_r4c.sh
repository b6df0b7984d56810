#!/usr/bin/env bash
# Round 2, third session, the remaining 2.8 GPU-minutes: Hisfrag20 model with 524,288 / 1,048,576 rows per chunk once more
# (alternating), then the parity + kernel suites with the larger chunk (would the default be safe to change?).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
hf() {
  timeout 100 python bench.py --workload hisfrag --items 128 --steps 1 --warmup 1 > gpurun_out/r02e_hisfrag128_$1.json 2> gpurun_out/r02e_hisfrag128_$1.err; local rc=$?
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
VITED_CHUNK_ROWS=1048576 hf chunk1m_2
hf base_2
VITED_CHUNK_ROWS=1048576 hf chunk1m_3
hf base_3
VITED_CHUNK_ROWS=1048576 timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_gpu_retrieval.py -q -x -p no:cacheprovider > gpurun_out/r02e_pytest_chunk1m.log 2>&1; el "parity + retrieval suites at 1,048,576 rows per chunk rc=$?"; tail -2 gpurun_out/r02e_pytest_chunk1m.log | cut -c1-200
