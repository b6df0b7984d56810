"""Runs / times the Hisfrag-shape attention launches (GPU box only): python tools/profile_attn_l64.py [time]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vited_b200 import _lib as L  # noqa: E402

P, H, hd, Np, D = 255, 6, 64, 1024, 384
M = P * (Np + 1)
torch.manual_seed(0)
qkv = torch.randn(M, 3 * D, device='cuda').to(L.act_dtype())
o = torch.empty(M, D, dtype=L.act_dtype(), device='cuda')
kv = torch.randn(32 * Np, 2 * D, device='cuda').to(L.act_dtype())
q = torch.randn(M, D, device='cuda').to(L.act_dtype())
idx = (torch.arange(P, device='cuda') % 32).int()


def self_attn(impl):
    L.check(L.lib.vited_op_attention(qkv.data_ptr(), 3 * D, qkv.data_ptr() + 2 * D, 3 * D, qkv.data_ptr() + 4 * D, 3 * D,
                                     o.data_ptr(), D, P, H, hd, Np, 1, Np, 1, P, None, hd ** -0.5, impl, None), 'attn')


def cross_attn(impl):
    L.check(L.lib.vited_op_attention(q.data_ptr(), D, kv.data_ptr(), 2 * D, kv.data_ptr() + 2 * D, 2 * D, o.data_ptr(), D,
                                     P, H, hd, Np, 1, Np, 0, 32, idx.data_ptr(), hd ** -0.5, impl, None), 'attn')


def self_attn_nocls(impl):
    # the same launch without class tokens (1024 queries / keys per sequence): the regular work items alone
    L.check(L.lib.vited_op_attention(qkv.data_ptr(), 3 * D, qkv.data_ptr() + 2 * D, 3 * D, qkv.data_ptr() + 4 * D, 3 * D,
                                     o.data_ptr(), D, P, H, hd, Np, 0, Np, 0, P, None, hd ** -0.5, impl, None), 'attn')


if len(sys.argv) > 1 and sys.argv[1] == 'time':
    import json
    for name, fn, fl in (('self', self_attn, 4.0 * P * H * 1025 * 1025 * hd), ('cross', cross_attn, 4.0 * P * H * 1025 * 1024 * hd),
                         ('self_nocls', self_attn_nocls, 4.0 * P * H * 1024 * 1024 * hd)):
        for impl in (0, 2):
            for _ in range(3):
                fn(impl)
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); fn(impl); e.record(); torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            ts.sort()
            ms = ts[len(ts) // 2]
            print(json.dumps(dict(op=f'attn_l64_{name}_impl{impl}', pairs=P, ms=ms, tflops=fl / ms / 1e9)))
else:
    for _ in range(2):
        self_attn(0)
    for _ in range(2):
        cross_attn(0)
    torch.cuda.synchronize()
    print('ok')
