"""Debug: clock trace of the long-sequence attention kernel, block 0. Needs tools/bin/trace/libvited_b200.so = the library
built with -DVITED_ATTN_TRACE (the nvcc line of vit-ed_b200/csrc/build.sh plus that define, -o tools/bin/trace/libvited_b200.so);
the current kernel keeps the softmax-side trace points only: MMA warp per group (wait P start / end, PV issued, QK issued) and softmax warp 0 per group
(wait S start / end, P stored)."""
import ctypes, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, 'tools', 'bin', 'trace', 'libvited_b200.so'))
vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
lib.vited_op_attention.argtypes = [vp, ci, vp, ci, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp, cf, ci, vp]
lib.vited_op_attention.restype = ci
P, H, hd, Np, D = 255, 6, 64, 1024, 384
M = P * (Np + 1)
torch.manual_seed(0)
qkv = torch.randn(M, 3 * D, device='cuda').to(torch.float16)
o = torch.empty(M, D, dtype=torch.float16, device='cuda')
for _ in range(2):
    assert lib.vited_op_attention(qkv.data_ptr(), 3 * D, qkv.data_ptr() + 2 * D, 3 * D, qkv.data_ptr() + 4 * D, 3 * D,
                                  o.data_ptr(), D, P, H, hd, Np, 1, Np, 1, P, None, hd ** -0.5, 0, None) == 0
torch.cuda.synchronize()
buf = np.zeros(4 * 128 * 4, dtype=np.uint64)
assert lib.vited_debug_attn_trace(buf.ctypes.data_as(vp)) == 0
tr = buf.reshape(4, 128, 4).astype(np.int64)
t0 = tr[2, 0, 0]
for g in range(2):
    print(f'--- group {g}: half tile | softmax: waitS_start waitS_end(+wait) P_stored(+compute) | mma: waitP_start waitP_end(+wait) PV_issued QK_issued')
    for j in range(20, 44):
        s, m = tr[2 + g, j], tr[g, j]
        print(f'  {j:3d} | {s[0]-t0:7d} {s[1]-t0:7d} (+{s[1]-s[0]:5d}) {s[2]-t0:7d} (+{s[2]-s[1]:5d}) | {m[0]-t0:7d} {m[1]-t0:7d} (+{m[1]-m[0]:5d}) {m[2]-t0:7d} (+{m[2]-m[1]:4d}) {m[3]-t0:7d} (+{m[3]-m[2]:4d})')
