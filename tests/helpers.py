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
    from vited_b200 import synthetic
    return synthetic.state_dict_shapes(mlp_ratio=mlp_ratio, in_chans=in_chans, **kw)


def make_gpu_model(kw, weight_seed):
    import vited_b200
    from vited_b200 import synthetic
    model = vited_b200.VisionTransformerCustom(mlp_ratio=4., qkv_bias=True, **kw)
    sd = synthetic.synthetic_state_dict(model, seed=weight_seed)
    model.load_state_dict(sd, strict=True)
    return model.cuda().eval(), sd
