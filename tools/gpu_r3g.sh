#!/usr/bin/env bash
# traces of the hot kernels on the current code + the long-sequence attention without class tokens
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/profile_attn_l64.py time > gpurun_out/r3g_l64.jsonl 2>&1; cat gpurun_out/r3g_l64.jsonl | cut -c1-140
timeout 120 python tools/trace_attn_p64.py > gpurun_out/r3g_trace_p64.txt 2>&1; head -30 gpurun_out/r3g_trace_p64.txt
timeout 120 python tools/trace_mlp_ln.py > gpurun_out/r3g_trace_mlp.txt 2>&1; tail -14 gpurun_out/r3g_trace_mlp.txt
timeout 120 python tools/trace_attn_l64.py > gpurun_out/r3g_trace_l64.txt 2>&1; head -30 gpurun_out/r3g_trace_l64.txt
