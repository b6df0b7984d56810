"""GPU: the Hisfrag training step on this repo's kernels -- forward with saved activations,
BCE-with-logits, backward to every parameter -- against (a) the fixture written by the reference's own prepare_data and
model file under autograd (tests/golden/train_step.npz: loss, logits, the L2 norm of EVERY parameter gradient, raw slices
of five of them) and (b) the oracle's autograd gradients, tensor by tensor. 16-bit GEMM operands with fp32 accumulation
(forward, dgrad and wgrad all run on the tcgen05 GEMM): gradients agree to about a percent."""
import os

import numpy as np
import pytest
import torch

from tests import helpers
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu

IMAGE_SEED = {'test_patch32_64': 11, 'small_hd64': 12}     # tests/golden/make_golden.py CASES


def _run(name):
    import vited_b200
    from vited_b200 import synthetic, train
    z = np.load(os.path.join(GOLDEN, 'train_step.npz'))
    _, kw = helpers.load_model_case(name)
    zc, _ = helpers.load_model_case(name)
    model, sd = helpers.make_gpu_model(kw, int(zc['weight_seed']))
    targets = z[f'{name}_targets']
    samples = synthetic.synthetic_images(len(targets), kw['img_size'], seed=IMAGE_SEED[name] + 100)
    torch.manual_seed(int(z[f'{name}_perm_seed']))            # the only random draw: randperm over the negatives
    loss, logits, groups, labels = train.train_step(model, samples.cuda(), targets)
    torch.cuda.synchronize()
    return z, kw, model, sd, samples, targets, loss, logits, groups, labels


@pytest.mark.parametrize('name', ['test_patch32_64', 'small_hd64'])
def test_train_step_matches_reference_fixture(name):
    z, kw, model, sd, samples, targets, loss, logits, groups, labels = _run(name)
    assert np.array_equal(labels.numpy(), z[f'{name}_labels'])
    assert np.array_equal(groups[:, 0].numpy(), z[f'{name}_first'])
    np.testing.assert_allclose(logits.cpu().numpy(), z[f'{name}_logits'], rtol=0, atol=2e-2)
    assert abs(loss.item() - float(z[f'{name}_loss'])) < 5e-3
    grads = {k: p.grad.detach().float().cpu() for k, p in model.named_parameters()}
    keys = [str(k) for k in z[f'{name}_grad_keys']]
    assert sorted(grads) == keys
    norms = z[f'{name}_grad_norms']
    worst = 0.0
    for k, want in zip(keys, norms):
        got = float(grads[k].double().norm())
        rel = abs(got - want) / max(want, 1e-8)
        worst = max(worst, rel)
        assert rel < 4e-2 or abs(got - want) < 1e-6, f'{k}: |grad| {got} vs reference {want}'
    for k in ('head.weight', 'cls_token', 'cross_blocks.0.cross_attn.kv.weight', 'blocks.0.attn.qkv.bias',
              'patch_embed.proj.weight'):
        want = z[f'{name}_grad::{k}']
        got = grads[k].flatten()[:64].numpy()
        scale = max(np.abs(want).max(), 1e-8)
        assert np.abs(got - want).max() < 5e-2 * scale, f'{k}: slice differs by {np.abs(got - want).max() / scale:.3f} of its max'
    print(f'{name}: loss {loss.item():.5f} (reference {float(z[f"{name}_loss"]):.5f}), worst gradient-norm error {worst:.4f}, '
          f'{model.train_launches} kernel launches')


@pytest.mark.parametrize('name', ['test_patch32_64', 'small_hd64'])
def test_train_step_gradients_match_oracle_autograd(name):
    from oracle import vited_oracle as orc
    z, kw, model, sd, samples, targets, loss, logits, groups, labels = _run(name)
    o_loss, o_logits, o_groups, o_labels, o_grads = orc.train_step(sd, kw['num_heads'], samples, targets,
                                                                   int(z[f'{name}_perm_seed']))
    assert torch.equal(groups, o_groups) and torch.equal(labels, o_labels)
    assert abs(loss.item() - o_loss.item()) < 5e-3
    worst, worst_key = 0.0, None
    for k, p in model.named_parameters():
        got, want = p.grad.detach().float().cpu().flatten().double(), o_grads[k].flatten().double()
        rel = float((got - want).norm() / max(float(want.norm()), 1e-12))
        if float(want.norm()) > 1e-7 and rel > worst:
            worst, worst_key = rel, k
        assert rel < 5e-2 or float((got - want).norm()) < 1e-6, f'{k}: relative L2 error {rel:.4f}'
    print(f'{name}: worst relative L2 error of a parameter gradient {worst:.4f} ({worst_key})')


def test_train_step_is_loud_without_cuda():
    import vited_b200
    from vited_b200 import train
    with pytest.raises(vited_b200.VitedError):
        train.train_step(None, torch.zeros(2, 3, 64, 64), [0, 0])
