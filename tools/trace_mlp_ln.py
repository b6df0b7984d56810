"""Debug: clock trace of the fused MLP + residual + LayerNorm kernel, CTA 0 (needs tools/bin/trace/libvited_b200.so = the
library built with -DVITED_MLP_TRACE, see tools/build_variants.sh)."""
import ctypes, math, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.environ.get('TRACE_LIB') or os.path.join(ROOT, 'tools', 'bin', 'trace', 'libvited_b200.so'))
vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
lib.vited_op_mlp_resid_ln.argtypes = [vp] * 9 + [ci, ci, ci, cf, vp]
lib.vited_op_mlp_resid_ln.restype = ci
M = int(os.environ.get('ROWS', 4032 * 65))
h = torch.randn(M, 384, device='cuda').to(torch.float16)
W1 = (torch.randn(1536, 384, device='cuda') / math.sqrt(384)).to(torch.float16)
W2 = (torch.randn(384, 1536, device='cuda') / math.sqrt(1536)).to(torch.float16)
b1 = torch.randn(1536, device='cuda'); b2 = torch.randn(384, device='cuda'); x = torch.randn(M, 384, device='cuda')
lw = torch.ones(384, device='cuda'); lb = torch.zeros(384, device='cuda')
for _ in range(2):
    assert lib.vited_op_mlp_resid_ln(h.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), x.data_ptr(),
                                     lw.data_ptr(), lb.data_ptr(), h.data_ptr(), M, 384, 1536, 1e-6, None) == 0
torch.cuda.synchronize()
buf = np.zeros(2 * 32 * 8, dtype=np.int64)
assert lib.vited_debug_mlp_trace(buf.ctypes.data_as(vp)) == 0
tr = buf.reshape(2, 32, 8)
print('tile | issuer: wait-A  G1(0..1)+wait-O  [wait-O]  main-loop  (of which wait-P  wait-W, per tile) | GELU/epilogue warp: gelu-phase (wait-S) wait-tfull pass1 pass2 | tile period')
for t in range(1, 12):
    m, e = tr[1, t], tr[0, t]
    wp = m[5] - tr[1, t - 1, 5]
    ww = m[6] - tr[1, t - 1, 6]
    print(f' {t:2d} | {m[1]-m[0]:6d} {m[3]-m[1]:7d} [{m[3]-m[2]:6d}] {m[4]-m[3]:7d} ({wp:6d} {ww:6d}) | {e[1]-e[0]:7d} ({e[6]:6d}) {e[2]-e[1]:6d} {e[3]-e[2]:6d} {e[4]-e[3]:6d} | {tr[1, t + 1, 0] - m[0]:7d}')
buf2 = np.zeros(32 * 8, dtype=np.int64)
if hasattr(lib, 'vited_debug_mlp_trace2') and lib.vited_debug_mlp_trace2(buf2.ctypes.data_as(vp)) == 0:
    t2 = buf2.reshape(32, 8)
    print('pass 1 of epilogue warp (q0, c0), cycles summed over its 6 chunks: TMEM load + wait for the residual box | wait for the '
          'previous store (out-box free) | add + st.shared | statistics | tcgen05.st + proxy fence + syncwarp | TMA issue (lane 0)')
    for t in range(1, 12):
        print(f' {t:2d} | ' + ' '.join(f'{int(v):6d}' for v in t2[t, :6]) + f' | sum {int(t2[t, :6].sum()):6d}')
