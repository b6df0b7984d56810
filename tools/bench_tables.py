"""GPU: time vited_puzzle_tables (device + D2H + host ordering) for 540- and 1000-piece puzzles."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vited_b200 import solver_tables  # noqa: E402

for n in (540, 1000):
    logits = torch.randn(n, n, 4, device='cuda') * 3
    order = np.random.default_rng(0).permutation(n)
    solver_tables.build_tables(logits, order=order)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        t = solver_tables.build_tables(logits, order=order)
    dt = (time.perf_counter() - t0) / 5
    print(f'N={n}: build_tables {dt * 1e3:.1f} ms end to end (replaces {4 * n * (n - 1):,} Python callbacks); '
          f'best buddies {(t.best_buddy >= 0).sum()}')
