#!/usr/bin/env bash
# round-2 second GPU pass: trace + ncu capture of the fused MLP kernel, K/V-once A/B of the puzzle cross-attention,
# Hisfrag workload smoke through grid.score_fragments, jitter build on the new kernel
mkdir -p gpurun_out
timeout 300 python tools/trace_mlp_ln.py > gpurun_out/r02b_trace_mlp.txt 2>&1; echo "trace rc=$?"
REPS=1 timeout 300 python tools/profile_ops.py > gpurun_out/r02b_plain.log 2>&1 && \
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mlp_ln_pair|gemm_ln_pair' -c 3 -o gpurun_out/r02b_prof_mlp python tools/profile_ops.py > gpurun_out/r02b_ncu.log 2>&1
echo "ncu rc=$?"
timeout 300 python tools/bench_ops.py 2>/dev/null | grep attn_cross > gpurun_out/r02b_attn_cross_base.jsonl
VITED_P64_KV_ONCE=1 timeout 300 python tools/bench_ops.py 2>/dev/null | grep attn_cross > gpurun_out/r02b_attn_cross_kvonce.jsonl
echo "kv-once A/B rc=$?"
timeout 900 python bench.py --workload hisfrag --items 192 --steps 1 --warmup 1 > gpurun_out/r02b_bench_hisfrag192.json 2> gpurun_out/r02b_bench_hisfrag192.err; echo "hisfrag rc=$?"
VITED_LIB=$PWD/tools/bin/jitter/libvited_b200.so timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "mlp_resid_ln or gemm_resid_ln" > gpurun_out/r02b_jitter.log 2>&1; echo "jitter rc=$?"
cat gpurun_out/r02b_trace_mlp.txt; cat gpurun_out/r02b_attn_cross_base.jsonl gpurun_out/r02b_attn_cross_kvonce.jsonl | cut -c1-160; cut -c1-900 gpurun_out/r02b_bench_hisfrag192.json; tail -3 gpurun_out/r02b_jitter.log
