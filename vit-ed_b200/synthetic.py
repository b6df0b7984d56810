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


def state_dict_shapes(img_size, patch_size, num_classes, embed_dim, depth, c_depth, num_heads=None, mlp_ratio=4.0,
                      in_chans=3):
    """state_dict key -> shape of a ViT-ED with these constructor arguments (models/vision_transformer.py:282-376),
    without building a module: what ``synthetic_state_dict`` needs when only the weights are wanted."""
    d, p = embed_dim, patch_size
    ne = (img_size // p) ** 2
    hid = int(d * mlp_ratio)
    sh = {'cls_token': (1, 1, d), 'pos_embed': (1, ne + 1, d), 'patch_embed.proj.weight': (d, in_chans, p, p),
          'patch_embed.proj.bias': (d,), 'norm.weight': (d,), 'norm.bias': (d,),
          'head.weight': (num_classes, d), 'head.bias': (num_classes,)}

    def lin(prefix, o, i):
        sh[prefix + '.weight'] = (o, i)
        sh[prefix + '.bias'] = (o,)

    def ln(prefix):
        sh[prefix + '.weight'] = (d,)
        sh[prefix + '.bias'] = (d,)

    for l in range(depth):
        b = f'blocks.{l}'
        ln(b + '.norm1'); lin(b + '.attn.qkv', 3 * d, d); lin(b + '.attn.proj', d, d)
        ln(b + '.norm2'); lin(b + '.mlp.fc1', hid, d); lin(b + '.mlp.fc2', d, hid)
    for l in range(c_depth):
        b = f'cross_blocks.{l}'
        ln(b + '.norm1'); lin(b + '.attn.qkv', 3 * d, d); lin(b + '.attn.proj', d, d)
        ln(b + '.norm_cross'); ln(b + '.norm_context')
        lin(b + '.cross_attn.q', d, d); lin(b + '.cross_attn.kv', 2 * d, d); lin(b + '.cross_attn.proj', d, d)
        ln(b + '.norm2'); lin(b + '.mlp.fc1', hid, d); lin(b + '.mlp.fc2', d, hid)
    return sh


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


def synthetic_fragments(n_writers, per_writer, img_size, seed=0, in_chans=3, writer_seed=None):
    """Fragments with a writer signal, for the retrieval checks (there is no Hisfrag20 data offline): every writer has
    its own stroke direction, stroke frequency and ink / paper colours (drawn from ``writer_seed``, default ``seed``);
    a fragment is that pattern at a random phase and stretch plus per-fragment noise (drawn from ``seed``). Returns
    (images [n, C, S, S] fp32 in [-1, 1], writer labels [n] int64), writer-major."""
    wrng = np.random.default_rng(seed if writer_seed is None else writer_seed)
    rng = np.random.default_rng(seed + 7919)
    yy, xx = np.meshgrid(np.arange(img_size, dtype=np.float32), np.arange(img_size, dtype=np.float32), indexing='ij')
    imgs, labels = [], []
    for w in range(n_writers):
        theta = wrng.uniform(0, np.pi)
        freq = wrng.uniform(0.02, 0.12)
        ink = wrng.uniform(0.0, 0.45, size=in_chans).astype(np.float32)
        paper = wrng.uniform(0.55, 1.0, size=in_chans).astype(np.float32)
        for _ in range(per_writer):
            phase = rng.uniform(0, 2 * np.pi)
            wob = rng.uniform(0.8, 1.25)
            s = 0.5 * (1 + np.sin(2 * np.pi * freq * wob * (np.cos(theta) * xx + np.sin(theta) * yy) + phase))
            img = paper[:, None, None] * s[None] + ink[:, None, None] * (1 - s[None])
            img = img + 0.08 * rng.standard_normal(img.shape).astype(np.float32)
            u8 = np.clip(np.round(img * 255.0), 0, 255).astype(np.uint8)
            imgs.append(u8)
            labels.append(w)
    t = torch.from_numpy(np.stack(imgs).astype(np.float32) / 255.0)
    return (t - 0.5) / 0.5, torch.tensor(labels, dtype=torch.int64)
