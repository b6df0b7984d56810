"""GPU box: piece preparation for one 1000-piece puzzle (configs[2] shape): host path (per-piece cv2 / PIL /
torchvision, what pieces_to_batch does) vs vited_prepare_pieces (BGR2LAB on the host + H2D + one kernel)."""
import ctypes, os, sys, time
import cv2
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vited_b200 import _lib, pieces, synthetic  # noqa: E402

img = synthetic.synthetic_puzzle_image(25, 40, 64, seed=0)
for _ in range(3):
    pieces.prepare_pieces_device(img, 64, 0.14, 64)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    out, _ = pieces.prepare_pieces_device(img, 64, 0.14, 64)
torch.cuda.synchronize()
dev_ms = (time.perf_counter() - t0) / 10 * 1e3
t0 = time.perf_counter()
for _ in range(10):
    lab = cv2.cvtColor(img, cv2.COLOR_BGR2LAB)
lab_ms = (time.perf_counter() - t0) / 10 * 1e3
lab_d = torch.from_numpy(lab).cuda()
side, off = pieces.erosion_crop(64, 0.14)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
e0.record()
for _ in range(10):
    _lib.check(_lib.lib.vited_prepare_pieces(ctypes.c_void_p(lab_d.data_ptr()), img.shape[0], img.shape[1], 64, side, off, 64,
                                             ctypes.c_void_p(out.data_ptr()), None, stream), 'prep')
e1.record(); torch.cuda.synchronize()
kern_ms = e0.elapsed_time(e1) / 10
t0 = time.perf_counter()
lab_pieces, _ = pieces.make_pieces_lab(img, 64, 0.14)
host = pieces.pieces_to_batch(lab_pieces, 64).cuda()
torch.cuda.synchronize()
host_ms = (time.perf_counter() - t0) * 1e3
gbs = (1000 * (side * side * 3 + 3 * 64 * 64 * 4)) / (kern_ms * 1e-3) / 1e9
print(f'1000 pieces: device path {dev_ms:.1f} ms end to end (cv2 BGR2LAB of the image on the host {lab_ms:.1f} ms, '
      f'ABI call incl. coefficient upload {kern_ms:.3f} ms = {gbs:.0f} GB/s algorithmic), host path {host_ms:.0f} ms; '
      f'identical: {torch.equal(out, host)}')
