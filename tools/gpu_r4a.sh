#!/usr/bin/env bash
# (ran on commit 2388076, which carried the FOLD_LN option; the option was measured and removed again: profiles/README.md item 46)
# Round 2, third session (10 GPU-minutes left): the whole GPU suite on the session's code (default path + the new
# FOLD_LN tests), A/B of FOLD_LN through bench.py (lean form), the parity suite with every engine folded
# (VITED_FOLD_LN=1), the M = 64 layout probe. Most valuable first: the call may be cut short.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02d_pytest_gpu.log 2>&1; el "pytest gpu rc=$?"; tail -3 gpurun_out/r02d_pytest_gpu.log | cut -c1-300
bench() {  # $1 = tag, env decides the variant
  timeout 150 python bench.py --no-cpu --no-extras > gpurun_out/r02d_bench_$1.json 2> gpurun_out/r02d_bench_$1.err; local rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02d_bench_$1.json'))
    c=d['roofline'].get('classes',{})
    print('[$1] rc=$rc value', round(d['value']), 'e2e', round(d['e2e']['value']), 'MHz', d['clocks']['sm_mhz'], {k:round(x['ms'],1) for k,x in c.items() if x['share']>0.04})
except Exception as ex:
    print('[$1] rc=$rc no bench line', ex)
PY
}
VITED_FOLD_LN=1 bench fold_1; el done
bench plain_1; el done
OPS=fused timeout 100 python tools/bench_ops.py > gpurun_out/r02d_ops_fused.jsonl 2> gpurun_out/r02d_ops_fused.err; el "ops rc=$?"; cut -c1-120 gpurun_out/r02d_ops_fused.jsonl
VITED_FOLD_LN=1 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_retrieval.py -q -p no:cacheprovider > gpurun_out/r02d_pytest_parity_folded.log 2>&1; el "parity suite, every engine folded rc=$?"; tail -3 gpurun_out/r02d_pytest_parity_folded.log | cut -c1-300
VITED_FOLD_LN=1 bench fold_2; el done
bench plain_2; el done
timeout 30 tools/bin/umma_probe m64 0 > gpurun_out/r02d_umma_probe_m64.txt 2>&1; el "probe m64 exit=$?"
