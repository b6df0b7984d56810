#!/usr/bin/env bash
mkdir -p gpurun_out
T=${T:-r02c}
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "mlp_resid_ln or gemm_resid_ln" > gpurun_out/${T}_pytest_mlp.log 2>&1; echo "kernel tests rc=$?"
tail -2 gpurun_out/${T}_pytest_mlp.log
timeout 300 python tools/trace_mlp_ln.py > gpurun_out/${T}_trace_mlp.txt 2>&1; echo "trace rc=$?"
cat gpurun_out/${T}_trace_mlp.txt; timeout 300 python tools/trace_gemm_ln.py 2>&1 | head -14 | tee gpurun_out/${T}_trace_gemm_ln.txt
timeout 300 python tools/bench_ops.py 2>/dev/null | grep -E "mlp_ln|gemm_ln" > gpurun_out/${T}_bench_ops.jsonl
cut -c1-200 gpurun_out/${T}_bench_ops.jsonl
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/${T}_bench_n1.json
python - <<'PY'
import json,os
T=os.environ.get('T','r02c')
d=json.load(open(f'gpurun_out/{T}_bench_n1.json'))
print(d['value'], d['clocks'])
for k,v in list(d['roofline']['classes'].items())[:8]: print('   ',k,v)
PY
