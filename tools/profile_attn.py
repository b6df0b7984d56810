"""Runs the puzzle-shape attention launches a few times (for ncu captures; GPU box only)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vited_b200 import _lib as L  # noqa: E402

P, H, hd, Np, D = 4032, 12, 32, 64, 384
M = P * 65
torch.manual_seed(0)
reps = int(os.environ.get('REPS', '2'))
qkv = torch.randn(M, 3 * D, device='cuda').to(L.act_dtype())
o = torch.empty(M, D, dtype=L.act_dtype(), device='cuda')
for _ in range(reps):
    L.check(L.lib.vited_op_attention(qkv.data_ptr(), 3 * D, qkv.data_ptr() + 2 * D, 3 * D, qkv.data_ptr() + 4 * D, 3 * D,
                                     o.data_ptr(), D, P, H, hd, Np, 1, Np, 1, P, None, hd ** -0.5, 0, None), 'attn')
kv = torch.randn(540 * Np, 2 * D, device='cuda').to(L.act_dtype())
q = torch.randn(M, D, device='cuda').to(L.act_dtype())
idx = (torch.arange(P, device='cuda') // 539).int()
for _ in range(reps):
    L.check(L.lib.vited_op_attention(q.data_ptr(), D, kv.data_ptr(), 2 * D, kv.data_ptr() + 2 * D, 2 * D, o.data_ptr(), D,
                                     P, H, hd, Np, 1, Np, 0, 540, idx.data_ptr(), hd ** -0.5, 0, None), 'attn')
torch.cuda.synchronize()
print('ok')
