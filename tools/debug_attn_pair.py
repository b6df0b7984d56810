"""Diagnostics for the two-units-per-tile puzzle attention (GPU box only): max error per row class against fp32 torch."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vited_b200 import _lib as L  # noqa: E402


def run(n_seq, H, has_cls, cross=False, n_ctx=3):
    hd, Np = 32, 64
    D = H * hd
    g = torch.Generator(device='cuda').manual_seed(n_seq * 7 + H)
    rows = n_seq * Np + (n_seq if (has_cls or cross) else 0)
    scale = hd ** -0.5
    o = torch.full((rows, D), float('nan'), dtype=L.act_dtype(), device='cuda')
    if not cross:
        qkv = torch.randn(rows, 3 * D, device='cuda', generator=g).to(L.act_dtype())
        st = L.lib.vited_op_attention(qkv.data_ptr(), 3 * D, qkv.data_ptr() + 2 * D, 3 * D, qkv.data_ptr() + 4 * D, 3 * D,
                                      o.data_ptr(), D, n_seq, H, hd, Np, has_cls, Np, has_cls, n_seq, None, scale, 0, None)
        patch = qkv[:n_seq * Np].view(n_seq, Np, 3 * D)
        seq = torch.cat([qkv[n_seq * Np:].view(n_seq, 1, 3 * D), patch], dim=1) if has_cls else patch
        q, k, v = [seq[..., i * D:(i + 1) * D].reshape(n_seq, -1, H, hd).permute(0, 2, 1, 3).float() for i in range(3)]
    else:
        qb = torch.randn(rows, D, device='cuda', generator=g).to(L.act_dtype())
        kv = torch.randn(n_ctx * Np, 2 * D, device='cuda', generator=g).to(L.act_dtype())
        idx = torch.randint(0, n_ctx, (n_seq,), device='cuda', generator=g, dtype=torch.int32)
        st = L.lib.vited_op_attention(qb.data_ptr(), D, kv.data_ptr(), 2 * D, kv.data_ptr() + 2 * D, 2 * D, o.data_ptr(), D,
                                      n_seq, H, hd, Np, 1, Np, 0, n_ctx, idx.data_ptr(), scale, 0, None)
        seq = torch.cat([qb[n_seq * Np:].view(n_seq, 1, D), qb[:n_seq * Np].view(n_seq, Np, D)], dim=1)
        q = seq.reshape(n_seq, Np + 1, H, hd).permute(0, 2, 1, 3).float()
        kvs = kv.view(n_ctx, Np, 2, H, hd)[idx.long()]
        k = kvs[:, :, 0].permute(0, 2, 1, 3).float()
        v = kvs[:, :, 1].permute(0, 2, 1, 3).float()
        has_cls = 1
    assert st == 0, L.last_error()
    torch.cuda.synchronize()
    ref = ((q @ k.transpose(-1, -2)) * scale).softmax(-1) @ v          # [B, H, Nq, hd]
    got_patch = o[:n_seq * Np].view(n_seq, Np, H, hd).permute(0, 2, 1, 3).float()
    if has_cls:
        got_cls = o[n_seq * Np:].view(n_seq, 1, H, hd).permute(0, 2, 1, 3).float()
        e_cls = (got_cls - ref[:, :, :1]).abs()
        e_patch = (got_patch - ref[:, :, 1:]).abs()
    else:
        e_cls = torch.zeros(1)
        e_patch = (got_patch - ref).abs()
    units = e_patch.reshape(n_seq * H, Np, hd)
    ua, ub = units[0::2], units[1::2]
    print(f'n_seq={n_seq} H={H} cls={has_cls} cross={cross}: patch rows max err unit A {ua.nan_to_num(99).max().item():.4g} '
          f'unit B {(ub.nan_to_num(99).max().item() if len(ub) else 0):.4g} | rows 0-31 {units[:, :32].nan_to_num(99).max().item():.4g} '
          f'rows 32-63 {units[:, 32:].nan_to_num(99).max().item():.4g} | cls rows {e_cls.nan_to_num(99).max().item():.4g}')


if __name__ == '__main__':
    for args in [(1, 2, 1), (1, 2, 0), (7, 12, 1), (3, 5, 1), (333, 12, 1)]:
        run(*args)
    run(9, 12, 1, cross=True, n_ctx=4)
    run(401, 12, 1, cross=True, n_ctx=7)
