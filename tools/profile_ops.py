"""Runs each hot kernel of the path once per REPS at the bench chunk shape (for ncu captures; GPU box only) and writes a
manifest of the launches in order (engine profile-class name, algorithmic bytes per launch with the engine's own
formulas) so that tools/ncu_traffic.py can pair the capture's rows with the bench line's kernel classes.
ROWS = token rows per launch (default 4032 * 65; the bench chunk is 7867 pairs * 65 = 511355)."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vited_b200 import _lib as L  # noqa: E402

P = int(os.environ.get('ROWS', 4032 * 65)) // 65
M = P * 65
D = 384
torch.manual_seed(0)
reps = int(os.environ.get('REPS', '2'))
manifest = []


def gemm(N, K, act, cls):
    A = torch.randn(M, K, device='cuda').to(L.act_dtype())
    W = (torch.randn(N, K, device='cuda') / math.sqrt(K)).to(L.act_dtype())
    b = torch.randn(N, device='cuda')
    C = torch.empty(M, N, dtype=L.act_dtype(), device='cuda')
    for _ in range(reps):
        L.check(L.lib.vited_op_gemm(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, act, 0, None), 'gemm')
        manifest.append(dict(cls=cls, kernel='gemm_tc_pair_kernel', algorithmic_bytes=2.0 * (M * K + N * K + M * N)))
    torch.cuda.synchronize()


gemm(1152, 384, 0, 'gemm_n1152_k384')
gemm(384, 384, 0, 'gemm_n384_k384')
if not os.environ.get('HOT'):
    gemm(1536, 384, 1, 'gemm_n1536_k384_gelu')
    gemm(384, 1536, 0, 'gemm_n384_k1536')


def gemm_ln(K):
    A = torch.randn(M, K, device='cuda').to(L.act_dtype())
    W = (torch.randn(384, K, device='cuda') / math.sqrt(K)).to(L.act_dtype())
    b = torch.randn(384, device='cuda')
    xx = torch.randn(M, 384, device='cuda')
    lw = torch.ones(384, device='cuda'); lb = torch.zeros(384, device='cuda')
    hh = torch.empty(M, 384, dtype=L.act_dtype(), device='cuda')
    for _ in range(reps):
        L.check(L.lib.vited_op_gemm_resid_ln(A.data_ptr(), W.data_ptr(), b.data_ptr(), xx.data_ptr(), lw.data_ptr(), lb.data_ptr(),
                                             hh.data_ptr(), M, 384, K, 1e-6, None), 'gemm_ln')
        manifest.append(dict(cls=f'gemm_ln_n384_k{K}', kernel='gemm_ln_pair_kernel',
                             algorithmic_bytes=2.0 * (M * K + 384 * K) + M * 384 * 10.0))
    torch.cuda.synchronize()


gemm_ln(384)
if not os.environ.get('HOT'):
    gemm_ln(1536)


def mlp_ln():
    hin = torch.randn(M, 384, device='cuda').to(L.act_dtype())
    W1 = (torch.randn(1536, 384, device='cuda') / math.sqrt(384)).to(L.act_dtype())
    W2 = (torch.randn(384, 1536, device='cuda') / math.sqrt(1536)).to(L.act_dtype())
    b1 = torch.randn(1536, device='cuda'); b2 = torch.randn(384, device='cuda')
    xx = torch.randn(M, 384, device='cuda')
    lw = torch.ones(384, device='cuda'); lb = torch.zeros(384, device='cuda')
    for _ in range(reps):
        L.check(L.lib.vited_op_mlp_resid_ln(hin.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), xx.data_ptr(),
                                            lw.data_ptr(), lb.data_ptr(), hin.data_ptr(), M, 384, 1536, 1e-6, None), 'mlp_ln')
        manifest.append(dict(cls='mlp_ln_d384_h1536', kernel='mlp_ln_pair_kernel',
                             algorithmic_bytes=4.0 * 1536 * 384 + M * 384 * 12.0))
    torch.cuda.synchronize()


mlp_ln()
H, hd, Np = 12, 32, 64
qkv = torch.randn(M, 3 * D, device='cuda').to(L.act_dtype())
o = torch.empty(M, D, dtype=L.act_dtype(), device='cuda')
for _ in range(reps):
    L.check(L.lib.vited_op_attention(qkv.data_ptr(), 3 * D, qkv.data_ptr() + 2 * D, 3 * D, qkv.data_ptr() + 4 * D, 3 * D,
                                     o.data_ptr(), D, P, H, hd, Np, 1, Np, 1, P, None, hd ** -0.5, 0, None), 'attn')
    manifest.append(dict(cls='attn_self', kernel='attn_p64_kernel', algorithmic_bytes=M * D * 8.0))
kv = torch.randn(540 * Np, 2 * D, device='cuda').to(L.act_dtype())
q = torch.randn(M, D, device='cuda').to(L.act_dtype())
idx = (torch.arange(P, device='cuda') // 539).int()
for _ in range(reps):
    L.check(L.lib.vited_op_attention(q.data_ptr(), D, kv.data_ptr(), 2 * D, kv.data_ptr() + 2 * D, 2 * D, o.data_ptr(), D,
                                     P, H, hd, Np, 1, Np, 0, 540, idx.data_ptr(), hd ** -0.5, 0, None), 'attn')
    manifest.append(dict(cls='attn_cross', kernel='attn_p64_kernel', algorithmic_bytes=M * D * 4.0))
if not os.environ.get('HOT'):
    x = torch.randn(M, D, device='cuda')
    delta = torch.randn(M, D, device='cuda').to(L.act_dtype())
    w = torch.ones(D, device='cuda'); bb = torch.zeros(D, device='cuda')
    h = torch.empty(M, D, dtype=L.act_dtype(), device='cuda')
    for _ in range(reps):
        L.check(L.lib.vited_op_resid_ln(x.data_ptr(), delta.data_ptr(), w.data_ptr(), bb.data_ptr(), h.data_ptr(), P, 64, 1, D,
                                        1e-6, None), 'ln')
        manifest.append(dict(cls='resid_ln', kernel='resid_ln_kernel', algorithmic_bytes=M * D * 12.0))
torch.cuda.synchronize()
if os.environ.get('MANIFEST'):
    json.dump(dict(rows=M, launches=manifest), open(os.environ['MANIFEST'], 'w'))
print('ok')
