"""CPU: the oracle (oracle/vited_oracle.py) against the fixtures generated from the reference's own code."""
import os

import numpy as np
import pytest
import torch

from tests import helpers
from tests.conftest import GOLDEN
from oracle import vited_oracle as orc
from vited_b200 import synthetic


@pytest.mark.parametrize('name', helpers.MODEL_CASES)
def test_oracle_matches_reference_forward(name):
    if name == 'hisfrag20_patch16_512' and os.environ.get('VITED_SKIP_SLOW'):
        pytest.skip('slow case skipped')
    z, kw = helpers.load_model_case(name)
    shapes = helpers.shapes_from_kwargs(kw)
    assert sorted(shapes) == [str(k) for k in z['keys']], 'state_dict key set differs from the reference model'
    sd = synthetic.synthetic_state_dict(shapes, seed=int(z['weight_seed']))
    assert sum(int(np.prod(s)) for s in shapes.values()) == int(z['n_params'])
    x1, x2 = helpers.case_inputs(z, kw)
    h = kw['num_heads']
    tokens = orc.forward(sd, h, x1, first_part=True)
    two_phase = orc.forward(sd, h, tokens, x2)
    one_shot = orc.forward(sd, h, torch.stack([x1, x2], dim=1))
    # fp32 vs fp32 (same math, different op order: SDPA vs explicit softmax, conv vs im2col) -> tight tolerance
    np.testing.assert_allclose(tokens[:, :4].numpy(), z['tokens_head'], rtol=0, atol=2e-4)
    np.testing.assert_allclose(tokens.double().sum(dim=(1, 2)).numpy(), z['tokens_sum'], rtol=1e-4, atol=5e-2)
    np.testing.assert_allclose(two_phase.numpy(), z['two_phase'], rtol=0, atol=2e-4)
    np.testing.assert_allclose(one_shot.numpy(), z['one_shot'], rtol=0, atol=2e-4)
    # the property the reference's only test asserts (tests/hisfrag_evaluation_test.py:129-143): cached == one-shot
    assert float((one_shot - two_phase).abs().max()) == 0.0


def test_integer_paths_match_reference():
    z = np.load(os.path.join(GOLDEN, 'integer_paths.npz'))
    for n in (1, 2, 5, 9):
        np.testing.assert_array_equal(orc.ordered_pairs(n), z[f'entries_{n}'].reshape(-1, 2))
    for n in (1, 4, 7):
        np.testing.assert_array_equal(orc.upper_tri_pairs(n), z[f'combos_{n}'])
    for n, world in z['sampler_cases']:
        pairs = orc.upper_tri_pairs(int(n))
        sizes = orc.sampler_sizes(pairs[:, 0], int(world))
        ref = z[f'sampler_{n}_{world}']
        for rank in range(int(world)):
            if ref[rank, 0] == -2:
                assert rank + 1 >= len(sizes)
            else:
                assert [sizes[rank], sizes[rank + 1]] == ref[rank].tolist()
    img = z['puzzle_image']
    for erosion, tag in ((0.0, '0p0'), (0.07, '0p07'), (0.14, '0p14')):
        rows, cols, top, left, crop, off = orc.crop_geometry(img.shape[0], img.shape[1], 64, erosion)
        assert [rows, cols] == z[f'grid_{tag}'].tolist()
        assert [crop, crop, 3] == z[f'piece_shape_{tag}'].tolist()
    # documented values: 60px / offset 2 at 7 %, 56px / offset 4 at 14 %
    assert orc.crop_geometry(1152, 1920, 64, 0.07)[4:] == (60, 2)
    assert orc.crop_geometry(1600, 2560, 64, 0.14)[4:] == (56, 4)


def test_distance_lookup_semantics():
    logits = np.zeros((2, 2, 4), dtype=np.float32)
    logits[0, 1] = [2.0, -1.0, 0.5, 3.0]
    pred = 1 - 1 / (1 + np.exp(-logits[0, 1]))
    assert orc.puzzle_distance_lookup(logits, 0, 1, 1, 3) == pytest.approx(pred[0] * 1000, rel=1e-6)
    assert orc.puzzle_distance_lookup(logits, 0, 1, 2, 0) == pytest.approx(pred[1] * 1000, rel=1e-6)
    assert orc.puzzle_distance_lookup(logits, 0, 1, 3, 1) == pytest.approx(pred[2] * 1000, rel=1e-6)
    assert orc.puzzle_distance_lookup(logits, 0, 1, 0, 2) == pytest.approx(pred[3] * 1000, rel=1e-6)
    assert orc.puzzle_distance_lookup(logits, 0, 1, 0, 0) == float('inf')
