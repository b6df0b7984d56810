"""Debug: clock trace of the puzzle attention kernel, block 0, first warp of every softmax group (needs
tools/bin/trace/libvited_b200.so = the library built with -DVITED_ATTN_TRACE)."""
import ctypes, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, 'tools', 'bin', 'trace', 'libvited_b200.so'))
vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
lib.vited_op_attention.argtypes = [vp, ci, vp, ci, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp, cf, ci, vp]
lib.vited_op_attention.restype = ci
P, H, hd, Np, D = 4032, 12, 32, 64, 384
M = P * 65
torch.manual_seed(0)
qkv = torch.randn(M, 3 * D, device='cuda').to(torch.float16)
o = torch.empty(M, D, dtype=torch.float16, device='cuda')
for _ in range(2):
    assert lib.vited_op_attention(qkv.data_ptr(), 3 * D, qkv.data_ptr() + 2 * D, 3 * D, qkv.data_ptr() + 4 * D, 3 * D,
                                  o.data_ptr(), D, P, H, hd, Np, 1, Np, 1, P, None, hd ** -0.5, 0, None) == 0
torch.cuda.synchronize()
buf = np.zeros(4 * 64 * 8, dtype=np.uint64)
assert lib.vited_debug_p64_trace(buf.ctypes.data_as(vp)) == 0
tr = buf.reshape(4, 64, 8).astype(np.int64)
t0 = tr[0, 10, 0]
names = ['waitS', 'softmax -> P', 'wait O (PV round trip)', 'epilogue']
for g in range(2):
    print(f'--- group {g}: unit | start | ' + ' | '.join(names) + ' | total')
    for k in range(10, 30):
        e = tr[g, k]
        d = [e[i + 1] - e[i] for i in range(4)]
        print(f'  {k:3d} | {e[0] - t0:7d} | ' + ' | '.join(f'{x:6d}' for x in d) + f' | {e[4] - e[0]:6d}')
