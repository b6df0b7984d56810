"""Debug: clock trace of the fused GEMM + residual + LayerNorm kernel, CTA 0 (needs tools/bin/trace/libvited_b200.so = the
library built with -DVITED_LN_TRACE: the nvcc line of vit-ed_b200/csrc/build.sh plus that define)."""
import ctypes, math, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, 'tools', 'bin', 'trace', 'libvited_b200.so'))
vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
lib.vited_op_gemm_resid_ln.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, cf, vp]
lib.vited_op_gemm_resid_ln.restype = ci
M = 4032 * 65
for K in (384, 1536):
    A = torch.randn(M, K, device='cuda').to(torch.float16)
    W = (torch.randn(384, K, device='cuda') / math.sqrt(K)).to(torch.float16)
    b = torch.randn(384, device='cuda'); x = torch.randn(M, 384, device='cuda')
    lw = torch.ones(384, device='cuda'); lb = torch.zeros(384, device='cuda')
    h = torch.empty(M, 384, dtype=torch.float16, device='cuda')
    for _ in range(2):
        assert lib.vited_op_gemm_resid_ln(A.data_ptr(), W.data_ptr(), b.data_ptr(), x.data_ptr(), lw.data_ptr(), lb.data_ptr(), h.data_ptr(), M, 384, K, 1e-6, None) == 0
    torch.cuda.synchronize()
    buf = np.zeros(3 * 32 * 8, dtype=np.uint64)
    assert lib.vited_debug_ln_trace(buf.ctypes.data_as(vp)) == 0
    tr = buf.reshape(3, 32, 8).astype(np.int64)
    t0 = tr[1, 0, 0]
    print(f'K = {K}: tile | MMA warp: wait-tempty  mainloop | epilogue warp(q0,c0): wait-tfull  pass1  stats  pass2 | warp(q3,c1): wait-tfull pass1 stats pass2 | tile period')
    for t in range(2, 12):
        m, e, f = tr[1, t], tr[0, t], tr[2, t]
        print(f'  {t:2d} | {m[1]-m[0]:7d} {m[2]-m[1]:7d} | {e[1]-e[0]:7d} {e[2]-e[1]:7d} {e[3]-e[2]:6d} {e[4]-e[3]:7d} | {f[1]-f[0]:7d} {f[2]-f[1]:7d} {f[3]-f[2]:6d} {f[4]-f[3]:7d} | {tr[1, t + 1, 0] - m[0]:7d}')
