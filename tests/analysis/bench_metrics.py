"""GPU box: retrieval metrics for a 4096-fragment similarity matrix (configs[3] size): device evaluator vs the
reference recipe on the host (fp16 cast, 1 - sim, numpy argsort + cumulative sums)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vited_b200 import grid  # noqa: E402
from oracle import vited_oracle as orc  # noqa: E402  (the CPU baseline of this measurement)

n = 4096
rng = np.random.default_rng(0)
labels = rng.integers(0, 512, n)
sim = torch.from_numpy(rng.normal(size=(n, n)).astype(np.float32)).cuda()
sim = torch.triu(sim) + torch.triu(sim, 1).T
grid.retrieval_metrics(sim, labels); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    got = grid.retrieval_metrics(sim, labels)
torch.cuda.synchronize()
dev_ms = (time.perf_counter() - t0) / 5 * 1e3
t0 = time.perf_counter()
want = orc.wi19_metrics(grid.similarity_to_distance(sim), labels, kind='stable')
host_ms = (time.perf_counter() - t0) * 1e3
print(f'N={n}: device evaluator {dev_ms:.1f} ms end to end; host recipe (D2H + fp16 + argsort + cumsums) {host_ms:.0f} ms; '
      f'max |diff| {max(abs(a - b) for a, b in zip(got, want)):.2e}')
