#!/usr/bin/env bash
# 8-GPU pass: Hisfrag grid (configs[3] model, 512 fragments) through grid.score_fragments at 1 / 2 / 4 / 8 GPUs, the
# puzzle batch (configs[2]) through grid.score_puzzles at 8 GPUs, the retrieval-clause test
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02g_smi.txt
ITEMS=${ITEMS:-512}
timeout 600 python -m pytest tests/test_gpu_retrieval.py -q -x -s -k three_decimals > gpurun_out/r02g_pytest_retrieval.log 2>&1; echo "retrieval test rc=$?"
grep -E "mAP|max \|logit|passed|failed" gpurun_out/r02g_pytest_retrieval.log
timeout 900 python bench.py --gpus 1 --workload hisfrag --items $ITEMS --steps 1 --warmup 1 > gpurun_out/r02g_bench_hisfrag${ITEMS}_n1.json 2> gpurun_out/r02g_bench_hisfrag${ITEMS}_n1.err; echo "hisfrag n1 rc=$?"
for N in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + N)) bench.py --gpus $N --workload hisfrag --items $ITEMS --steps 1 --warmup 1 > gpurun_out/r02g_bench_hisfrag${ITEMS}_n$N.json 2> gpurun_out/r02g_bench_hisfrag${ITEMS}_n$N.err; echo "hisfrag n$N rc=$?"
done
python - <<'PY'
import json, os
items = os.environ.get('ITEMS', '512')
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.load(open(f'gpurun_out/r02g_bench_hisfrag{items}_n{n}.json'))
    except Exception as e:
        print(n, 'no line', e); continue
    base = base or d['value']
    print(f"N={n}: {d['value']:.0f} pairs/s, {d['ms_per_step']:.0f} ms/step, efficiency {d['value'] / (n * base):.3f}, imbalance {d['imbalance']:.3f}, "
          f"tensor frac {d['step_tensor_frac']:.3f}, clocks {d['clocks']['sm_mhz']}, ranks: " + ' '.join(f"{r['rows']}:{r['pairs']}p:{r['ms_local']:.0f}ms" for r in d['per_rank']))
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29560 bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/r02g_bench_n8.json 2> gpurun_out/r02g_bench_n8.err; echo "puzzle n8 rc=$?"
cut -c1-400 gpurun_out/r02g_bench_n8.json
