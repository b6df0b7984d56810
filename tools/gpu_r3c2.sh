#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/debug_attn_pair.py > gpurun_out/r3c_debug.log 2>&1; echo "debug rc=$?"; tail -12 gpurun_out/r3c_debug.log | cut -c1-250
if ! grep -q "cross=True" gpurun_out/r3c_debug.log; then
  timeout 600 compute-sanitizer --tool memcheck python tools/debug_attn_pair.py > gpurun_out/r3c_sanitizer.log 2>&1; grep -m 12 -A6 "Invalid\|Misaligned\|Error" gpurun_out/r3c_sanitizer.log | cut -c1-200
  exit 0
fi
bash tools/gpu_r3c.sh
