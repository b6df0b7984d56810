"""Shared helpers for the tests (golden loading, seeded models)."""
import ast
import os

import numpy as np
import torch

from tests.conftest import GOLDEN

MODEL_CASES = ['test_patch32_64', 'small_hd64', 'small_hd32', 'puzzle_patch8_64', 'hisfrag20_patch16_512']


def load_model_case(name):
    z = np.load(os.path.join(GOLDEN, f'model_{name}.npz'), allow_pickle=False)
    kw = ast.literal_eval(str(z['kwargs']))
    return z, kw


def case_inputs(z, kw):
    from vited_b200 import synthetic
    n = int(z['n_pairs'])
    imgs = synthetic.synthetic_images(2 * n, kw['img_size'], seed=int(z['image_seed']))
    return imgs[:n], imgs[n:]


def shapes_from_kwargs(kw, mlp_ratio=4.0, in_chans=3):
    """state_dict key -> shape for a ViT-ED of these constructor kwargs (independent of any model class)."""
    d, p = kw['embed_dim'], kw['patch_size']
    ne = (kw['img_size'] // p) ** 2
    hid = int(d * mlp_ratio)
    sh = {'cls_token': (1, 1, d), 'pos_embed': (1, ne + 1, d), 'patch_embed.proj.weight': (d, in_chans, p, p),
          'patch_embed.proj.bias': (d,), 'norm.weight': (d,), 'norm.bias': (d,),
          'head.weight': (kw['num_classes'], d), 'head.bias': (kw['num_classes'],)}

    def lin(prefix, o, i):
        sh[prefix + '.weight'] = (o, i)
        sh[prefix + '.bias'] = (o,)

    def ln(prefix):
        sh[prefix + '.weight'] = (d,)
        sh[prefix + '.bias'] = (d,)

    for l in range(kw['depth']):
        b = f'blocks.{l}'
        ln(b + '.norm1'); lin(b + '.attn.qkv', 3 * d, d); lin(b + '.attn.proj', d, d)
        ln(b + '.norm2'); lin(b + '.mlp.fc1', hid, d); lin(b + '.mlp.fc2', d, hid)
    for l in range(kw['c_depth']):
        b = f'cross_blocks.{l}'
        ln(b + '.norm1'); lin(b + '.attn.qkv', 3 * d, d); lin(b + '.attn.proj', d, d)
        ln(b + '.norm_cross'); ln(b + '.norm_context')
        lin(b + '.cross_attn.q', d, d); lin(b + '.cross_attn.kv', 2 * d, d); lin(b + '.cross_attn.proj', d, d)
        ln(b + '.norm2'); lin(b + '.mlp.fc1', hid, d); lin(b + '.mlp.fc2', d, hid)
    return sh


def make_gpu_model(kw, weight_seed):
    import vited_b200
    from vited_b200 import synthetic
    model = vited_b200.VisionTransformerCustom(mlp_ratio=4., qkv_bias=True, **kw)
    sd = synthetic.synthetic_state_dict(model, seed=weight_seed)
    model.load_state_dict(sd, strict=True)
    return model.cuda().eval(), sd
