"""Fixture for the retrieval metrics: runs the REFERENCE's own misc/wi19_evaluate.get_metrics (numpy only, imports
unchanged from /root/reference) on seeded distance matrices and stores inputs' seeds + outputs.
Run in the build container: python tests/golden/make_golden_metrics.py"""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


def case_inputs(seed, n, n_classes):
    """Seeded symmetric fp16 distance matrix with a zero diagonal (the layout hisfrag.py:281-296 produces) and labels
    with every class present at least twice."""
    rng = np.random.default_rng(seed)
    labels = np.concatenate([np.arange(n_classes), np.arange(n_classes), rng.integers(0, n_classes, n - 2 * n_classes)])
    rng.shuffle(labels)
    sim = rng.normal(size=(n, n)).astype(np.float32)
    same = labels[None, :] == labels[:, None]
    sim = sim + 1.5 * same                       # same-writer fragments are somewhat closer
    sim = np.triu(sim) + np.triu(sim, 1).T
    sim = sim.astype(np.float16)
    dist = (1 - sim).astype(np.float16)
    np.fill_diagonal(dist, -np.inf)              # self is always the first column, as with a trained model
    return dist, labels


CASES = [(0, 40, 8), (1, 300, 25), (2, 150, 60)]

if __name__ == '__main__':
    spec = importlib.util.spec_from_file_location('ref_wi19', os.path.join(REF, 'misc', 'wi19_evaluate.py'))
    wi19 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(wi19)
    out = {'cases': np.array(CASES, dtype=np.int64)}
    for seed, n, k in CASES:
        dist, labels = case_inputs(seed, n, k)
        out[f'metrics_{seed}'] = np.array(wi19.get_metrics(dist, labels), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, 'metrics_wi19.npz'), **out)
    print({k: v for k, v in out.items()})
