"""Debug: per-tile clock trace of the long-sequence attention kernel (needs tools/bin/libvited_trace.so, built with
-DVITED_ATTN_TRACE). Prints, for block 0 and both groups, the cycle deltas between the phases of each key tile."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, 'tools', 'bin', 'libvited_trace.so'))
vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
lib.vited_op_attention.argtypes = [vp, ci, vp, ci, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp, cf, ci, vp]
lib.vited_op_attention.restype = ci
lib.vited_last_error.restype = ctypes.c_char_p
P, H, hd, Np, D = 255, 6, 64, 1024, 384
M = P * (Np + 1)
torch.manual_seed(0)
qkv = torch.randn(M, 3 * D, device='cuda').bfloat16()
o = torch.empty(M, D, dtype=torch.bfloat16, device='cuda')
for _ in range(2):
    st = lib.vited_op_attention(qkv.data_ptr(), 3 * D, qkv.data_ptr() + 2 * D, 3 * D, qkv.data_ptr() + 4 * D, 3 * D,
                                o.data_ptr(), D, P, H, hd, Np, 1, Np, 1, P, None, hd ** -0.5, 0, None)
    assert st == 0, lib.vited_last_error()
torch.cuda.synchronize()
buf = np.zeros(2 * 64 * 12, dtype=np.uint64)
assert lib.vited_debug_attn_trace(buf.ctypes.data_as(vp)) == 0
tr = buf.reshape(2, 64, 12).astype(np.int64)
t0 = tr[:, 0, 0].min()
names = ['S ready', 'ld done', 'max+vote', 'turn', 'exp done', 'st done', 'bar', 'mma issued']
for g in range(2):
    print(f'group {g}: tile  start   ' + '  '.join(f'{n:>9s}' for n in names[1:]) + '   -> next S')
    for t in range(1, 30):
        e = tr[g, t]
        if e[0] == 0:
            continue
        d = [e[i] - e[i - 1] if e[i] and e[i - 1] else -1 for i in range(1, 8)]
        nxt = tr[g, t + 1, 0] - e[7] if tr[g, t + 1, 0] else -1
        print(f'   {t:3d} {e[0] - t0:8d}   ' + '  '.join(f'{x:9d}' for x in d) + f'   {nxt:8d}   | pv_issue {e[8]-e[6]:5d} commits {e[9]-e[8]:5d} kvwait {e[10]-e[9]:5d} qk {e[7]-e[10]:5d}')
