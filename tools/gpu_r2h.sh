#!/usr/bin/env bash
# new property tests + smoke + a bench line that takes roofline.traffic from profiles/ncu_traffic.json
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -s -k "full_size or fragment_grid_properties" > gpurun_out/r02h_pytest_props.log 2>&1; echo "property tests rc=$?"; tail -3 gpurun_out/r02h_pytest_props.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02h_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02h_smoke.log
timeout 900 python bench.py > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench_n1.json'))
r=d['roofline']
print(d['value'], d['e2e']['value'], d['clocks'], d['step_tensor_frac'])
print({k:r[k] for k in ('bound','achieved','frac','traffic','traffic_source','share_of_step')})
for k,v in r['other_kernels'].items(): print(k, v['traffic'], v['traffic_source'])
PY
