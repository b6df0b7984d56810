"""Deterministic synthetic workloads (weights and inputs) shared by bench.py, the tests and the golden-vector
generator. No datasets or checkpoints exist offline, so BASELINE.json's configs are built from seeds."""
import numpy as np
import torch


def synthetic_state_dict(model_or_shapes, seed=0):
    """A state_dict for the ViT-ED key set, drawn key by key from a seeded CPU generator so it does not depend on
    module construction order. Distributions follow the reference's effective init: encoder Linear
    weights ~ N(0, .02) with zero bias; cross-block Linear weights and biases ~ U(+-1/sqrt(fan_in)); LayerNorm
    weights near 1; pos_embed ~ N(0, .02)."""
    if isinstance(model_or_shapes, dict):
        shapes = model_or_shapes
    else:
        shapes = {k: tuple(v.shape) for k, v in model_or_shapes.state_dict().items()}
    g = torch.Generator(device='cpu')
    out = {}
    for i, key in enumerate(sorted(shapes)):
        shape = shapes[key]
        g.manual_seed(seed * 1000003 + i)
        if '.norm' in key or key.startswith('norm'):
            if key.endswith('weight'):
                t = 1.0 + 0.1 * torch.randn(shape, generator=g)
            else:
                t = 0.05 * torch.randn(shape, generator=g)
        elif key == 'cls_token':
            t = 0.02 * torch.randn(shape, generator=g)
        elif key == 'pos_embed':
            t = 0.02 * torch.randn(shape, generator=g)
        elif key.startswith('cross_blocks.'):
            fan_in = shape[-1] if key.endswith('weight') else _fan_in_of_bias(shapes, key)
            bound = 1.0 / np.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif key.startswith('patch_embed.'):
            fan_in = int(np.prod(shapes['patch_embed.proj.weight'][1:]))
            bound = 1.0 / np.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif key.endswith('bias'):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        out[key] = t.float()
    return out


def _fan_in_of_bias(shapes, key):
    return shapes[key[:-len('bias')] + 'weight'][-1]


def synthetic_images(n, img_size, seed=0, in_chans=3, smooth=True):
    """[n, C, S, S] fp32 in [-1, 1] (what ToTensor + Normalize(.5,.5) produce). ``smooth`` mixes a low-frequency
    gradient with noise so logits spread (SURVEY 8d 'Weights')."""
    rng = np.random.default_rng(seed)
    noise = rng.integers(0, 256, size=(n, in_chans, img_size, img_size), dtype=np.uint8).astype(np.float32)
    if smooth:
        yy, xx = np.meshgrid(np.linspace(0, 1, img_size, dtype=np.float32), np.linspace(0, 1, img_size, dtype=np.float32),
                             indexing='ij')
        phase = rng.uniform(0, 2 * np.pi, size=(n, in_chans, 1, 1)).astype(np.float32)
        freq = rng.uniform(0.5, 3.0, size=(n, in_chans, 1, 1)).astype(np.float32)
        grad = 127.5 * (1 + np.sin(2 * np.pi * freq * (xx[None, None] + 0.7 * yy[None, None]) + phase))
        u8 = np.clip(np.round(0.6 * grad + 0.4 * noise), 0, 255)
    else:
        u8 = noise
    t = torch.from_numpy(u8.astype(np.float32) / 255.0)
    return (t - 0.5) / 0.5


def synthetic_puzzle_image(rows, cols, piece=64, seed=0):
    """BGR uint8 [rows*piece, cols*piece, 3] from default_rng(seed)."""
    rng = np.random.default_rng(seed)
    h, w = rows * piece, cols * piece
    noise = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8).astype(np.float32)
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float32) / h, np.arange(w, dtype=np.float32) / w, indexing='ij')
    base = np.stack([127.5 * (1 + np.sin(2 * np.pi * (f1 * xx + f2 * yy))) for f1, f2 in ((1.5, 0.5), (0.7, 2.1), (2.3, 1.1))],
                    axis=-1)
    return np.clip(np.round(0.6 * base + 0.4 * noise), 0, 255).astype(np.uint8)
