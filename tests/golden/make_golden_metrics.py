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

# (seed, n, n_classes, self_first): similarity matrices whose fp16 distances are DISTINCT within every row, so that
# numpy's (unstable) argsort has exactly one answer and an implementation with another sort must match bit for bit
SIM_CASES = [(10, 64, 9, 1), (11, 500, 40, 1), (12, 1000, 333, 0), (13, 257, 3, 0), (14, 120, 100, 1)]


def sim_case_inputs(seed, n, n_classes, self_first):
    """fp32 similarity logits [n, n] (what grid.score_fragments returns; not symmetric here -- get_metrics does not
    need it) on the grid k / 1024, k < 1024, a different permutation per row: fp16(sim) and 1 - fp16(sim) are exact
    and distinct within a row. self_first: the diagonal holds the row maximum (self is the nearest item, the case
    get_metrics assumes when it drops the first column); otherwise self lands anywhere."""
    assert n <= 1024
    rng = np.random.default_rng(seed)
    labels = np.concatenate([np.arange(n_classes), rng.integers(0, n_classes, n - n_classes)])
    rng.shuffle(labels)
    sim = np.stack([rng.permutation(1024)[:n] for _ in range(n)]).astype(np.float32) / np.float32(1024)
    if self_first:
        for i in range(n):
            j = int(sim[i].argmax())
            sim[i, j], sim[i, i] = sim[i, i], sim[i, j]
    return sim, labels


def sim_to_distance(sim):
    """hisfrag.py:283-296: fp16 similarity, then 1 - similarity in fp16."""
    import torch
    return (1 - torch.from_numpy(sim).type(torch.float16)).numpy()

if __name__ == '__main__':
    spec = importlib.util.spec_from_file_location('ref_wi19', os.path.join(REF, 'misc', 'wi19_evaluate.py'))
    wi19 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(wi19)
    out = {'cases': np.array(CASES, dtype=np.int64)}
    for seed, n, k in CASES:
        dist, labels = case_inputs(seed, n, k)
        out[f'metrics_{seed}'] = np.array(wi19.get_metrics(dist, labels), dtype=np.float64)
    out['sim_cases'] = np.array(SIM_CASES, dtype=np.int64)
    for case in SIM_CASES:
        sim, labels = sim_case_inputs(*case)
        dist = sim_to_distance(sim)
        assert dist.dtype == np.float16 and all(len(np.unique(r)) == len(r) for r in dist)
        with np.errstate(invalid='ignore', divide='ignore'):
            out[f'sim_metrics_{case[0]}'] = np.array(wi19.get_metrics(dist, labels), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, 'metrics_wi19.npz'), **out)
    print({k: v for k, v in out.items()})
