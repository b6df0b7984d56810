"""GPU, 2 ranks under NCCL: the package's own multi-GPU entry points (grid.score_puzzle, grid.score_puzzles,
grid.score_fragments: row sharding + the all-gather of the score blocks) against the same grids scored on one GPU.
Skipped on boxes with fewer than two GPUs (the sharding / gather logic itself is covered on CPU with gloo in
tests/test_host_logic.py)."""
import os

import numpy as np
import pytest
import torch

from tests import helpers

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import vited_b200
    from vited_b200 import grid, synthetic
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        out = {}
        # puzzles: the real puzzle model on a batch of three small puzzles of different sizes + one single puzzle
        z, kw = helpers.load_model_case('puzzle_patch8_64')
        model, _ = helpers.make_gpu_model(kw, 3)
        sizes = [11, 6, 14]
        puzzles = [synthetic.synthetic_images(n, kw['img_size'], seed=40 + i).cuda() for i, n in enumerate(sizes)]
        many = grid.score_puzzles(model, puzzles)
        one = grid.score_puzzle(model, puzzles[2])
        single = [model.score_grid(p, vited_b200.GRID_ORDERED_OFFDIAG, 0, p.shape[0]) for p in puzzles]
        out['puzzles'] = max((a - b).abs().max().item() for a, b in zip(many, single))
        out['puzzle'] = (one - single[2]).abs().max().item()
        out['puzzles_equal'] = all(torch.equal(a, b) for a, b in zip(many, single))
        # fragments: upper-triangular grid, rows split as the reference sampler does
        z, kw = helpers.load_model_case('small_hd64')
        fmodel, _ = helpers.make_gpu_model(kw, 2)
        frags = synthetic.synthetic_images(23, kw['img_size'], seed=9).cuda()
        sim = grid.score_fragments(fmodel, frags)
        upper = fmodel.score_grid(frags, vited_b200.GRID_UPPER_TRI_DIAG, 0, 23)[..., 0]
        out['fragments'] = (sim - grid.mirror_upper(upper)).abs().max().item()
        out['symmetric'] = bool(torch.equal(sim, sim.t()))
        # training step: every rank its own batch, gradients averaged over the ranks (what DDP does)
        from vited_b200 import train
        z, kw = helpers.load_model_case('small_hd64')
        tmodel, _ = helpers.make_gpu_model(kw, 2)
        samples = synthetic.synthetic_images(6, kw['img_size'], seed=70 + rank).cuda()
        gen = torch.Generator().manual_seed(5)
        train.train_step(tmodel, samples, [0, 0, 0, 1, 1, 1], generator=gen, all_reduce=False)
        local = torch.cat([p.grad.reshape(-1) for p in tmodel.parameters()]).clone()
        gen = torch.Generator().manual_seed(5)
        train.train_step(tmodel, samples, [0, 0, 0, 1, 1, 1], generator=gen, all_reduce=True)
        averaged = torch.cat([p.grad.reshape(-1) for p in tmodel.parameters()])
        both = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(both, local)
        want = sum(both) / world
        out['ddp'] = ((averaged - want).abs().max() / want.abs().max()).item()
        out['ddp_differs_from_local'] = bool((averaged - local).abs().max() > 0)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_sharded_grids_under_nccl_equal_single_gpu_grids():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, out in results:
        # the same kernels on the same rows: a row shard launches other chunk shapes than the full grid, so allow the
        # rounding noise of the unfused / fused sub-block boundary (see test_grid_properties_puzzle_model)
        assert out['puzzles'] < 1e-2 and out['puzzle'] < 1e-2 and out['fragments'] < 1e-2, (rank, out)
        assert out['symmetric'], rank
        assert out['ddp'] < 1e-6 and out['ddp_differs_from_local'], (rank, out)
