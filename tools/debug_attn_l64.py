"""Debug: where does the long-sequence attention differ from the reference? (GPU box)"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vited_b200 import _lib as L

def run(n_seq, H, hd, n_patch, has_cls):
    D = H * hd
    rows = n_seq * n_patch + (n_seq if has_cls else 0)
    g = torch.Generator(device='cuda').manual_seed(1)
    qkv = torch.randn(rows, 3 * D, device='cuda', generator=g).to(L.act_dtype())
    o = torch.full((rows, D), float('nan'), dtype=L.act_dtype(), device='cuda')
    st = L.lib.vited_op_attention(qkv.data_ptr(), 3 * D, qkv.data_ptr() + 2 * D, 3 * D, qkv.data_ptr() + 4 * D, 3 * D, o.data_ptr(), D,
                                  n_seq, H, hd, n_patch, has_cls, n_patch, has_cls, n_seq, None, hd ** -0.5, 0, None)
    L.check(st, 'attn'); torch.cuda.synchronize()
    patch = qkv[:n_seq * n_patch].view(n_seq, n_patch, 3 * D)
    seq = torch.cat([qkv[n_seq * n_patch:].view(n_seq, 1, 3 * D), patch], dim=1) if has_cls else patch
    q, k, v = [seq[..., i * D:(i + 1) * D].reshape(n_seq, -1, H, hd).permute(0, 2, 1, 3).float() for i in range(3)]
    ref = ((q @ k.transpose(-1, -2)) * hd ** -0.5).softmax(-1) @ v          # [B,H,N,hd]
    gp = o[:n_seq * n_patch].view(n_seq, n_patch, H, hd).permute(0, 2, 1, 3).float()
    got = torch.cat([o[n_seq * n_patch:].view(n_seq, 1, H, hd).permute(0, 2, 1, 3).float(), gp], dim=2) if has_cls else gp
    err = (got - ref).abs()
    err = torch.nan_to_num(err, nan=9.0)
    print(f'case {n_seq}x{H}x{hd}x{n_patch} cls={has_cls}: max err {err.max().item():.4f}')
    off = 1 if has_cls else 0
    e = err[:, :, off:].reshape(n_seq, H, n_patch // 128, 128, hd).amax(dim=(3, 4))   # per (b, h, 128-row block)
    bad = (e > 2e-2).nonzero()
    print('  bad (b,h,qblock) count', len(bad), 'of', e.numel(), 'first', bad[:12].tolist())
    if has_cls:
        ec = err[:, :, 0].amax(-1)
        print('  cls rows bad', (ec > 2e-2).sum().item(), 'of', ec.numel())
    # per column error pattern for the first bad block
    if len(bad):
        b, h, qb = bad[0].tolist()
        blk = err[b, h, off + qb * 128: off + qb * 128 + 128]
        print('  first bad block: rows bad', (blk.amax(1) > 2e-2).sum().item(), 'cols bad', (blk.amax(0) > 2e-2).sum().item())

for c in [(1, 1, 64, 256, 0), (1, 1, 64, 256, 1), (2, 1, 64, 512, 0), (9, 3, 64, 512, 0), (40, 6, 64, 1024, 1)]:
    run(*c)
