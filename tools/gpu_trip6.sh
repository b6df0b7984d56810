#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python tools/profile_ops.py"
$CMD > gpurun_out/plain6.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"attn_mma|gemm_tc|resid_ln" -c 14 -o gpurun_out/prof_ops_v4 -f $CMD > gpurun_out/ncu6.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/ncu6.log
