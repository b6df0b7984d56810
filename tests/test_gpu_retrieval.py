"""GPU: the retrieval metrics computed on the device (C-ABI vited_retrieval_rows) against the fixture written by the
reference's own wi19_evaluate.get_metrics (tie-free cases: exact same ranking) and against the oracle with ties."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _mk():
    spec = importlib.util.spec_from_file_location('mk_metrics', os.path.join(GOLDEN, 'make_golden_metrics.py'))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    return mk


@pytest.mark.parametrize('case', _mk().SIM_CASES, ids=lambda c: f'seed{c[0]}_n{c[1]}')
def test_device_metrics_match_reference_fixture(case):
    from vited_b200 import grid
    mk = _mk()
    z = np.load(os.path.join(GOLDEN, 'metrics_wi19.npz'))
    sim, labels = mk.sim_case_inputs(*case)
    got = np.array(grid.retrieval_metrics(torch.from_numpy(sim).cuda(), labels))
    np.testing.assert_allclose(got, z[f'sim_metrics_{case[0]}'], rtol=0, atol=1e-12)   # nan == nan (singleton queries)


@pytest.mark.parametrize('n,n_classes,scale', [(300, 20, 1.0), (2048, 150, 3.0), (4096, 400, 0.05), (33, 33, 1.0)])
def test_device_rows_match_oracle_with_ties(n, n_classes, scale):
    """Real-valued similarities: fp16 rounding makes many distances tie (scale 0.05: thousands per row). The device
    breaks ties by ascending index, as the oracle's stable argsort does; integer outputs must be identical and the
    precision sums equal to fp64 rounding."""
    import ctypes
    from oracle import vited_oracle as orc
    from vited_b200 import _lib
    rng = np.random.default_rng(n)
    labels = rng.integers(0, n_classes, n).astype(np.int32)
    sim = (rng.normal(size=(n, n)) * scale).astype(np.float32)
    sim[rng.random((n, n)) < 0.01] = -0.0
    want = orc.wi19_rows(orc.sim_to_distance(sim), labels)
    sim_d, lab_d = torch.from_numpy(sim).cuda(), torch.from_numpy(labels).cuda()
    outs = [torch.empty(n, dtype=dt, device='cuda') for dt in (torch.int32, torch.float64, torch.int32, torch.int32, torch.int32)]
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.lib.vited_retrieval_rows(p(sim_d), p(lab_d), n, *[p(t) for t in outs],
                                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), 'retrieval_rows')
    got = [t.cpu().numpy() for t in outs]
    for k in (0, 2, 3, 4):
        assert np.array_equal(got[k], want[k]), k
    np.testing.assert_allclose(got[1], want[1], rtol=1e-13, atol=0)
    # and the four headline numbers against the oracle's get_metrics with the same tie convention
    from vited_b200 import grid
    a = np.array(grid.retrieval_metrics(sim_d, labels))
    b = np.array(orc.wi19_metrics(orc.sim_to_distance(sim), labels, kind='stable'))
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-12)


def test_metrics_of_a_scored_fragment_grid():
    """End of the Hisfrag path on the device: score_fragments -> retrieval_metrics equals the reference recipe
    (fp16 cast, 1 - sim, get_metrics) applied to the same matrix on the host."""
    import vited_b200
    from oracle import vited_oracle as orc
    from tests import helpers
    from vited_b200 import grid, synthetic
    z, kw = helpers.load_model_case('small_hd64')
    model, _ = helpers.make_gpu_model(kw, 2)
    images = synthetic.synthetic_images(14, kw['img_size'], seed=8)
    sim = grid.score_fragments(model, images.cuda())
    labels = np.array([0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7])
    got = np.array(grid.retrieval_metrics(sim, labels))
    want = np.array(orc.wi19_metrics(grid.similarity_to_distance(sim), labels, kind='stable'))
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    with pytest.raises(vited_b200.VitedError):
        grid.retrieval_metrics(sim.cpu(), labels)
    with pytest.raises(vited_b200.VitedError):
        grid.retrieval_metrics(sim, labels[:-1])


def _retrieval_case(which):
    """Production-shaped model (embed_dim 384, 6 heads of 64, 256 patch tokens) with the weights of
    tests/golden/make_retrieval_fixture.py: seeded synthetic matrices + vectors trained with the reference's pair
    objective so that the similarity logits separate writers. 'heldout': 8 training writers x 6 new fragments (clean
    separation, all metrics 1.0); 'mixed': 8 training writers + 4 unseen ones x 4 fragments (mAP 0.79, top-1 0.73:
    every rank matters)."""
    import ast
    from tests.conftest import GOLDEN
    from vited_b200 import synthetic
    z = np.load(os.path.join(GOLDEN, 'retrieval_weights.npz'))
    kw = ast.literal_eval(str(z['__kwargs__']))
    weight_seed, writer_seed, n_writers = [int(v) for v in z['__seeds__']]
    sd = synthetic.synthetic_state_dict(synthetic.state_dict_shapes(**kw), seed=weight_seed)
    for k in z.files:
        if not k.startswith('__'):
            assert sd[k].shape == z[k].shape, k
            sd[k] = torch.from_numpy(z[k])
    if which == 'heldout':
        images, labels = synthetic.synthetic_fragments(n_writers, 6, kw['img_size'], seed=5, writer_seed=writer_seed)
    else:
        images, labels = synthetic.synthetic_fragments(12, 4, kw['img_size'], seed=6, writer_seed=writer_seed)
    return z, kw, sd, images, labels.numpy(), torch.from_numpy(z[f'__{which}_sim__']), z[f'__{which}_metrics__']


@pytest.mark.parametrize('which', ['heldout', 'mixed'])
def test_retrieval_top1_and_map_identical_to_three_decimals(which):
    """North-star clause: "identical retrieval top-1 / mAP to 3 decimals" between the CUDA-scored and the fp32
    reference-scored similarity matrix, through the consumer's own recipe (hisfrag.py:281-309: fp16 similarity,
    1 - sim, wi19_evaluate.get_metrics as misc/wi19_evaluate.py:12-56) -- the property the reference's only test of
    this path asserts between its two scoring modes (tests/hisfrag_evaluation_test.py:129-143). 48 fragments =
    1,176 pairs = 302k token rows: the fused Linear + residual + LayerNorm / fused MLP kernels and the tcgen05
    long-sequence attention score them. The oracle matrix was written by the fixture generator; a corner of it is
    recomputed here."""
    import vited_b200
    from oracle import vited_oracle as orc
    from vited_b200 import grid
    z, kw, sd, images, labels, sim_orc, m_stored = _retrieval_case(which)
    model = vited_b200.VisionTransformerCustom(mlp_ratio=4., qkv_bias=True, **kw)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    corner = orc.score_fragment_grid(sd, kw['num_heads'], images[:6])
    np.testing.assert_allclose(corner.numpy(), sim_orc[:6, :6].numpy(), rtol=0, atol=2e-4)   # fp32 op-order noise only
    sim = grid.score_fragments(model, images.cuda())
    err = (sim.cpu() - sim_orc).abs().max().item()
    m_cuda = orc.wi19_metrics(grid.similarity_to_distance(sim), labels, kind='stable')
    m_orc = orc.wi19_metrics(orc.sim_to_distance(sim_orc), labels, kind='stable')
    m_dev = grid.retrieval_metrics(sim, labels)
    same = torch.from_numpy(labels[:, None] == labels[None, :])
    print(f'[{vited_b200.ACT_NAME}] {which}: max |logit - fp32 oracle| {err:.5f}; logits same-writer mean {sim_orc[same].mean():.2f} '
          f'other-writer mean {sim_orc[~same].mean():.2f}')
    print('mAP / top-1 / Pr@10 / Pr@100   CUDA-scored:', np.round(m_cuda, 5), ' oracle-scored:', np.round(m_orc, 5),
          ' device evaluator:', np.round(m_dev, 5))
    assert err < 2e-2
    np.testing.assert_allclose(m_orc, m_stored, rtol=0, atol=1e-12)
    assert m_orc[0] > 0.7, 'the fixture weights no longer separate writers'
    for a, b, c, name in zip(m_cuda, m_orc, m_dev, ('mAP', 'top-1', 'Pr@10', 'Pr@100')):
        assert abs(a - b) < 5e-4 and round(a, 3) == round(b, 3), f'{name}: CUDA-scored {a} vs oracle-scored {b}'
        assert abs(c - a) < 1e-12, f'{name}: device evaluator {c} vs host recipe {a} on the same matrix'
