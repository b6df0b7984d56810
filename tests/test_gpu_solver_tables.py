"""GPU: the solver distance tables computed on the device (C-ABI vited_puzzle_tables) against the fixture written by
the reference's own InterPieceDistance class and against the oracle restatement -- bit-exact, every table."""
import os
import sys

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN

sys.path.insert(0, GOLDEN)
import make_golden_tables as mg  # noqa: E402

pytestmark = pytest.mark.gpu


def _compare(tables, want):
    n = tables.n
    assert tables.asym_dist.dtype == np.uint32 and np.array_equal(tables.asym_dist, want['asym_dist'])
    assert np.array_equal(tables.min_dist, want['min_dist'])
    assert np.array_equal(tables.second_dist, want['second_dist'])
    assert tables.asym_compat.dtype == np.float32 and np.array_equal(tables.asym_compat, want['asym_compat'])
    assert tables.mutual_compat.dtype == np.float32 and np.array_equal(tables.mutual_compat, want['mutual_compat'])
    assert np.array_equal(tables.n_candidates, want['candidates'].sum(-1))
    first = np.where(want['candidates'].any(-1), want['candidates'].argmax(-1), -1)
    assert np.array_equal(tables.candidate, first)
    assert np.array_equal(tables.best_buddy, want['best_buddy'])
    assert [(a, b) for (a, b, _) in tables.start_piece_ordering] == [tuple(r) for r in want['start_order'].tolist()]
    assert np.array_equal(np.array([float(c) for (_, _, c) in tables.start_piece_ordering]), want['start_compat'])
    assert tables.asym_dist.shape == (n, 4, n)


@pytest.mark.parametrize('rules', ['numpy1', 'numpy2'])
@pytest.mark.parametrize('case', mg.CASES, ids=lambda c: f'seed{c[0]}_{c[1]}x{c[2]}')
def test_device_tables_match_reference_fixture(case, rules):
    """Both scalar-rule sets the reference's own class wrote: `pred[k] * 1000.` in float64 (NumPy 1.x, the reference's
    pinned stack; the default) and in float32 (NumPy >= 2)."""
    from vited_b200 import solver_tables
    z = np.load(os.path.join(GOLDEN, 'solver_tables.npz'))
    d, order = mg.case_inputs(*case)
    tables = solver_tables.build_tables(torch.from_numpy(d).cuda(), order=order, scores_are_logits=False,
                                        scalar_rules=rules)
    infix = '_np1' if rules == 'numpy1' else ''
    keys = ['asym_dist', 'asym_compat', 'mutual_compat', 'min_dist', 'second_dist', 'candidates', 'best_buddy',
            'start_order', 'start_compat']
    _compare(tables, {k: z[f'{k}{infix}_{case[0]}'] for k in keys})


@pytest.mark.parametrize('case', [(21, 10, 13, 0, 1), (22, 12, 12, 2, 5), (23, 1, 1, 0, 0), (24, 17, 16, 0, 0)],
                         ids=lambda c: f'{c[1]}x{c[2]}')
def test_device_tables_match_oracle(case):
    """Larger grids than the fixture holds (up to 272 pieces: rows longer than one pass of the 256-thread CTA), ties,
    exact zeros, and the 1-piece edge case."""
    from oracle import vited_oracle as orc
    from vited_b200 import solver_tables
    d, order = mg.case_inputs(*case)
    tables = solver_tables.build_tables(torch.from_numpy(d).cuda(), order=order, scores_are_logits=False)
    _compare(tables, orc.solver_tables(d, order))
    ident = solver_tables.build_tables(torch.from_numpy(d).cuda(), order=None, scores_are_logits=False,
                                       scalar_rules='numpy2')
    _compare(ident, orc.solver_tables(d, np.arange(len(order)), scalar_rules='numpy2'))


def test_logit_mode_equals_torch_sigmoid_on_device():
    """scores_are_logits=1 applies 1 - sigmoid exactly as evaluation.py:109-114 does on the GPU (torch.sigmoid, fp32)."""
    from vited_b200 import solver_tables
    n = 96
    g = torch.Generator(device='cuda').manual_seed(3)
    logits = torch.randn(n, n, 4, device='cuda', generator=g) * 4
    a = solver_tables.build_tables(logits, scores_are_logits=True)
    b = solver_tables.build_tables(1.0 - torch.sigmoid(logits), scores_are_logits=False)
    for name in ('asym_dist', 'min_dist', 'second_dist', 'asym_compat', 'mutual_compat', 'best_buddy', 'candidate'):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name


def test_tables_from_model_scores_and_errors():
    import vited_b200
    from tests import helpers
    from vited_b200 import grid, solver_tables, synthetic
    z, kw = helpers.load_model_case('small_hd32')
    model, _ = helpers.make_gpu_model(kw, 1)
    images = synthetic.synthetic_images(12, kw['img_size'], seed=4)
    logits = grid.score_puzzle(model, images.cuda())
    tables = solver_tables.build_tables(logits, order=np.random.default_rng(0).permutation(12))
    assert tables.asym_dist.shape == (12, 4, 12) and len(tables.start_piece_ordering) == 12
    assert (tables.asym_dist[np.arange(12), :, np.arange(12)] == 2 ** 31 - 1).all()
    with pytest.raises(vited_b200.VitedError):
        solver_tables.build_tables(logits.cpu())
    with pytest.raises(vited_b200.VitedError):
        solver_tables.build_tables(logits, order=[0] * 12)
