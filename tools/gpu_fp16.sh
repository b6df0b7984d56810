#!/usr/bin/env bash
# fp16-vs-bf16 operand A/B on one box: tests, error against the fixtures / oracle, bench of both builds
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BF=$PWD/tools/bin/bf16/libvited_b200.so   # tools/build_variants.sh
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 -p no:cacheprovider -rP 2>&1 | grep -v "^$" | grep "passed\|failed\|Error\|error\|argmax agreement\|assert" | head -20
timeout 300 python tests/analysis/fixture_err.py 2>&1 | grep "^\["
VITED_LIB=$BF timeout 300 python tests/analysis/fixture_err.py 2>&1 | grep "^\["
VITED_LIB=$BF timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k north_star -p no:cacheprovider -rP 2>&1 | grep "argmax agreement\|passed\|failed"
for i in 1 2; do
  timeout 600 python bench.py --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fp16', round(d['value']), d['clocks'])"
  VITED_LIB=$BF timeout 600 python bench.py --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bf16', round(d['value']), d['clocks'])"
done
