"""Last 50 GPU-seconds of round 2: the resumable form of grid.score_fragments against the one-call form on the real
engine (small Hisfrag-style model), then smoke()."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
t0 = time.time()
import torch
import vited_b200
from vited_b200 import grid, synthetic
from tests import helpers
z, kw = helpers.load_model_case('small_hd64')
model, sd = helpers.make_gpu_model(kw, 5)
images = synthetic.synthetic_images(9, kw['img_size'], seed=3).cuda()
one = grid.score_fragments(model, images)
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, 'test_result_rank{rank}.pt')
    blocks = grid.score_fragments(model, images, resume_path=path, block_rows=2, save_every=2)
    again = grid.score_fragments(model, images, resume_path=path, block_rows=2)
torch.cuda.synchronize()
print('resumable == one call:', torch.equal(blocks, one), ' reloaded == one call:', torch.equal(again, one),
      ' max |diff|', (blocks - one).abs().max().item(), f' [{time.time() - t0:.1f}s]', flush=True)
import __graft_entry__ as g
g.smoke()
print(f'[{time.time() - t0:.1f}s]')
