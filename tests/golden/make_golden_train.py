"""Fixture for the Hisfrag training step (the product side is NOT built yet -- this pins the oracle
that the backward kernels of a later round will be checked against).

Runs in the build container: python tests/golden/make_golden_train.py
  * pair construction: the reference's own ``prepare_data`` (hisfrag.py:117-155) is extracted from
    /root/reference/hisfrag.py at generation time and executed unchanged (with ``.cuda()`` as the identity, the
    reference's ``misc.utils.get_combinations``, a seeded ``torch.randperm``);
  * forward + backward: the reference's own models/vision_transformer.py (timm through tests/golden/_timm_shim.py) in
    train mode with drop_path_rate 0, ``BCEWithLogitsLoss`` (hisfrag.py:60-61), fp32 on the CPU, gradients by autograd.
Stored: the pair lists and labels, the loss, the logits, the L2 norm of every parameter gradient and a few raw slices.
"""
import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden  # noqa: E402  (reference model loader, CASES)
from vited_b200 import synthetic  # noqa: E402

# name -> (targets of the batch: 3 images per writer as MPerClassSampler(m=3) draws them (hisfrag.py:107), perm seed)
TRAIN_CASES = {
    'test_patch32_64': ([0, 0, 0, 1, 1, 1, 2, 2, 2], 5),
    'small_hd64': ([4, 4, 4, 7, 7, 7, 1, 1, 1, 9, 9, 9], 6),
}
SLICES = ['head.weight', 'cls_token', 'cross_blocks.0.cross_attn.kv.weight', 'blocks.0.attn.qkv.bias', 'patch_embed.proj.weight']


def reference_prepare_data():
    """The reference's HisfragTrainer.prepare_data as a plain function of (self, samples, targets)."""
    sys.path.insert(0, REF)
    from misc.utils import get_combinations
    src = open(os.path.join(REF, 'hisfrag.py')).read()
    tree = ast.parse(src)
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == 'prepare_data'
              and any(isinstance(m, ast.Name) and m.id == 'get_combinations' for m in ast.walk(n)))
    ns = {'torch': torch, 'get_combinations': get_combinations}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), 'hisfrag.py:prepare_data', 'exec'), ns)
    return ns['prepare_data']


class _Self:
    def __init__(self, model):
        self.model = model
        self.config = type('C', (), {'AMP_ENABLE': False})()


if __name__ == '__main__':
    ref_mod = make_golden.load_reference_model_module()
    prepare_data = reference_prepare_data()
    torch.Tensor.cuda = lambda self, *a, **k: self          # the reference moves masks to the GPU; identity here
    out = {}
    for name, (targets, perm_seed) in TRAIN_CASES.items():
        kw, _, wseed, iseed = make_golden.CASES[name]
        torch.manual_seed(0)
        model = ref_mod.VisionTransformerCustom(mlp_ratio=4., qkv_bias=True, drop_path_rate=0., **kw)
        sd = synthetic.synthetic_state_dict(model, seed=wseed)
        model.load_state_dict(sd, strict=True)
        model.train()
        samples = synthetic.synthetic_images(len(targets), kw['img_size'], seed=iseed + 100)
        t = torch.tensor(targets)
        torch.manual_seed(perm_seed)                          # the only random draw: torch.randperm over the negatives
        (x, x1), labels = prepare_data(_Self(model), samples, t)
        logits = model(x1, x)                                 # train_step (hisfrag.py:157-159)
        loss = torch.nn.BCEWithLogitsLoss()(logits, labels)
        loss.backward()
        # recover the pair list from what prepare_data returned: x = samples[groups[:, 0]]
        first = [int((samples == xi).flatten(1).all(1).nonzero()[0]) for xi in x]
        grads = {k: p.grad.detach() for k, p in model.named_parameters()}
        keys = sorted(grads)
        out[f'{name}_targets'] = np.array(targets, dtype=np.int64)
        out[f'{name}_perm_seed'] = perm_seed
        out[f'{name}_labels'] = labels.numpy()
        out[f'{name}_first'] = np.array(first, dtype=np.int64)
        out[f'{name}_logits'] = logits.detach().numpy()
        out[f'{name}_loss'] = np.float64(loss.item())
        out[f'{name}_grad_keys'] = np.array(keys)
        out[f'{name}_grad_norms'] = np.array([float(grads[k].double().norm()) for k in keys])
        for k in SLICES:
            out[f'{name}_grad::{k}'] = grads[k].flatten()[:64].numpy()
        print(name, 'pairs', len(labels), 'pos', int(labels.sum()), 'loss', loss.item(),
              'grad norm', float(torch.sqrt(sum(g.double().pow(2).sum() for g in grads.values()))))
    np.savez_compressed(os.path.join(HERE, 'train_step.npz'), **out)
