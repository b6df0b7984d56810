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
