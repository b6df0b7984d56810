"""GPU: correctness of the 4-CTA-cluster (multicast weights) GEMM variant, selected with VITED_GEMM_QUAD=1."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vited_b200 import _lib as L
g = torch.Generator(device='cuda').manual_seed(1)
for (M, N, K, act) in [(80000, 1536, 384, 1), (80001, 384, 1536, 0), (76001, 1152, 384, 0), (262080, 384, 384, 0), (79990, 768, 384, 0), (262080, 1536, 384, 1)]:
    A = torch.randn(M, K, device='cuda', generator=g).bfloat16(); W = (torch.randn(N, K, device='cuda', generator=g) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device='cuda', generator=g); C = torch.full((M, N), float('nan'), dtype=torch.bfloat16, device='cuda')
    for rep in range(3):
        st = L.lib.vited_op_gemm(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, act, 0, None)
        torch.cuda.synchronize(); assert st == 0, L.last_error()
    ref = A.float() @ W.float().t() + b
    ref = torch.nn.functional.gelu(ref) if act else ref
    assert torch.isfinite(C.float()).all(), ('nan', M, N, K)
    err = (C.float() - ref).abs(); tol = 1e-2 * ref.abs() + 2e-2
    bad = err > tol
    print('shape', M, N, K, act, 'max err', float(err.max()), 'bad', int(bad.sum()))
    assert not bad.any()
print('quad ok')
