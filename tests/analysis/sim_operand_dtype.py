"""CPU study (test infrastructure, uses the oracle): how much of the logit error against the fp32 reference comes from
rounding the GEMM / attention OPERANDS to a 16-bit type, for bf16 (8 significand bits) and fp16 (11 bits).

The engine keeps the residual stream, LayerNorm statistics, softmax and all accumulators in fp32 and rounds exactly
these tensors to 16 bits: weights, LayerNorm outputs, q/k/v, the attention probabilities and output, the GELU
activations, the im2col pixels. This script re-runs the oracle's arithmetic with a rounding hook at those points.

    python tests/analysis/sim_operand_dtype.py [n_items]
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import importlib.util

from oracle import vited_oracle as orc


def load_pkg():
    spec = importlib.util.spec_from_file_location('vited_b200', os.path.join(ROOT, 'vit-ed_b200', '__init__.py'),
                                                  submodule_search_locations=[os.path.join(ROOT, 'vit-ed_b200')])
    mod = importlib.util.module_from_spec(spec)
    sys.modules['vited_b200'] = mod
    spec.loader.exec_module(mod)
    return mod


class Sim:
    def __init__(self, sd, heads, dt):
        self.sd, self.H, self.dt = sd, heads, dt
        self.w = {k: self.r(v) if (k.endswith('.weight') and v.dim() >= 2) else v for k, v in sd.items()}

    def r(self, t):
        return t if self.dt is None else t.to(self.dt).float()

    def ln(self, x, p):
        return self.r(F.layer_norm(x, (x.shape[-1],), self.sd[p + '.weight'], self.sd[p + '.bias'], orc.EPS))

    def lin(self, x, p):
        return F.linear(x, self.w[p + '.weight'], self.sd.get(p + '.bias'))

    def sdpa(self, q, k, v):
        B, Nq, D = q.shape
        hd = D // self.H
        sp = lambda t: self.r(t).view(B, t.shape[1], self.H, hd).transpose(1, 2)
        q, k, v = sp(q), sp(k), sp(v)
        s = (q @ k.transpose(-1, -2)) * hd ** -0.5
        m = s.amax(-1, keepdim=True)
        p = torch.exp(s - m)
        l = p.sum(-1, keepdim=True)           # fp32 row sum of the unrounded probabilities, as in the kernels
        o = (self.r(p) @ v) / l
        return self.r(o.transpose(1, 2).reshape(B, Nq, D))

    def attn(self, h, p):
        q, k, v = self.lin(h, p + '.qkv').chunk(3, dim=-1)
        return self.lin(self.sdpa(q, k, v), p + '.proj')

    def mlp(self, h, p):
        return self.lin(self.r(F.gelu(self.lin(h, p + '.fc1'))), p + '.fc2')

    def embed(self, images):
        w = self.w['patch_embed.proj.weight']
        d, c, p, _ = w.shape
        B, _, S, _ = images.shape
        g = S // p
        cols = images.view(B, c, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, c * p * p)
        return self.r(cols) @ w.view(d, -1).t() + self.sd['patch_embed.proj.bias']

    def encode(self, images):
        x = self.embed(images) + self.sd['pos_embed'][:, 1:]
        depth, _ = orc._depths(self.sd)
        for l in range(depth):
            p = f'blocks.{l}'
            x = x + self.attn(self.ln(x, p + '.norm1'), p + '.attn')
            x = x + self.mlp(self.ln(x, p + '.norm2'), p + '.mlp')
        return x

    def decode(self, ctx, images):
        x = torch.cat([self.sd['cls_token'].expand(images.shape[0], -1, -1), self.embed(images)], 1) + self.sd['pos_embed']
        _, c_depth = orc._depths(self.sd)
        for l in range(c_depth):
            p = f'cross_blocks.{l}'
            x = x + self.attn(self.ln(x, p + '.norm1'), p + '.attn')
            q = self.lin(self.ln(x, p + '.norm_cross'), p + '.cross_attn.q')
            k, v = self.lin(self.ln(ctx, p + '.norm_context'), p + '.cross_attn.kv').chunk(2, dim=-1)
            x = x + self.lin(self.sdpa(q, k, v), p + '.cross_attn.proj')
            x = x + self.mlp(self.ln(x, p + '.norm2'), p + '.mlp')
        x = self.ln(x[:, :1], 'norm')
        return F.linear(x[:, 0], self.w['head.weight'], self.sd['head.bias'])

    @torch.no_grad()
    def grid(self, images, batch=120):
        n = images.shape[0]
        tok = self.encode(images)
        pairs = torch.as_tensor(orc.ordered_pairs(n))
        out = torch.zeros(n, n, self.sd['head.weight'].shape[0])
        for p0 in range(0, len(pairs), batch):
            sub = pairs[p0:p0 + batch]
            out[sub[:, 0], sub[:, 1]] = self.decode(tok[sub[:, 0]], images[sub[:, 1]])
        return out


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    load_pkg()
    from vited_b200 import synthetic
    from tests import helpers
    torch.set_num_threads(os.cpu_count())
    kw = dict(img_size=64, patch_size=8, num_classes=4, embed_dim=384, depth=8, c_depth=8, num_heads=12)
    shapes = helpers.shapes_from_kwargs(kw)
    off = ~torch.eye(n, dtype=torch.bool)
    for seed in (0, 5):
        sd = synthetic.synthetic_state_dict(shapes, seed=seed)
        images = synthetic.synthetic_images(n, 64, seed=33)
        want = Sim(sd, 12, None).grid(images)
        chk = (want - orc.score_puzzle_grid(sd, 12, images, batch=120)).abs().max().item()
        print(f'weights seed {seed}: fp32 simulation vs oracle: {chk:.2e}')
        for name, dt in (('bf16', torch.bfloat16), ('fp16', torch.float16)):
            got = Sim(sd, 12, dt).grid(images)
            err = (got - want).abs()
            agree = (got.argmax(-1) == want.argmax(-1))[off].float().mean().item()
            print(f'weights seed {seed}  operands {name}: max err {err[off].max().item():.5f}  mean err '
                  f'{err[off].mean().item():.5f}  argmax agreement {agree:.4f}', flush=True)
