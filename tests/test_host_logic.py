"""CPU: host-side mirror of the reference interface -- key set, integer bookkeeping, consumer layouts, sharding."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from tests import helpers
from tests.conftest import GOLDEN
import vited_b200
from vited_b200 import grid, pieces


@pytest.mark.parametrize('name', helpers.MODEL_CASES)
def test_state_dict_layout_matches_reference(name):
    z, kw = helpers.load_model_case(name)
    model = vited_b200.VisionTransformerCustom(mlp_ratio=4., qkv_bias=True, **kw)
    sd = model.state_dict()
    assert sorted(sd.keys()) == [str(k) for k in z['keys']]
    assert len(sd) == int(z['n_state_dict'])
    assert sum(p.numel() for p in model.parameters()) == int(z['n_params'])
    shapes = helpers.shapes_from_kwargs(kw)
    for k, v in sd.items():
        assert tuple(v.shape) == shapes[k], k


def test_build_model_and_unsupported_options():
    for name in ('puzzle', 'hisfrag', 'test'):
        cfg = vited_b200.get_config(name)
        m = vited_b200.build_model(cfg)
        assert m.embed_dim == cfg.MODEL.PJS.EMBED_DIM and m.num_classes == cfg.MODEL.NUM_CLASSES
    cfg = vited_b200.get_config('puzzle')
    cfg.MODEL.TYPE = 'resnet'
    with pytest.raises(NotImplementedError):
        vited_b200.build_model(cfg)
    with pytest.raises(NotImplementedError):
        vited_b200.VisionTransformerCustom(img_size=64, patch_size=8, embed_dim=384, qk_norm=True)
    with pytest.raises(AssertionError):
        vited_b200.VisionTransformerCustom(img_size=64, patch_size=8, embed_dim=100, num_heads=12)


def test_pair_enumeration_and_sampler_against_reference_vectors():
    z = np.load(os.path.join(GOLDEN, 'integer_paths.npz'))
    for n in (1, 2, 5, 9):
        got = grid.ordered_pairs(n)
        np.testing.assert_array_equal(got, z[f'entries_{n}'].reshape(-1, 2))
        for idx, (i, j) in enumerate(got):
            assert grid.ordered_pair_index(int(i), int(j), n) == idx
    for n in (1, 4, 7):
        np.testing.assert_array_equal(grid.upper_tri_pairs(n), z[f'combos_{n}'])
    assert len(grid.ordered_pairs(540)) == 291060 and len(grid.upper_tri_pairs(4096)) == 8390656
    for n, world in z['sampler_cases']:
        ref = z[f'sampler_{n}_{world}']
        for rank in range(int(world)):
            lo, hi = grid.hisfrag_row_range(int(n), int(world), rank)
            if ref[rank, 0] == -2:
                assert lo == hi
            else:
                assert [lo, hi] == ref[rank].tolist()


def test_equal_row_range_covers_everything():
    for n, world in ((540, 1), (540, 8), (1000, 8), (7, 8), (20000, 4)):
        rows = []
        for r in range(world):
            lo, hi = grid.equal_row_range(n, world, r)
            rows.extend(range(lo, hi))
        assert rows == list(range(n))


def test_piece_preparation_bit_exact_with_reference():
    z = np.load(os.path.join(GOLDEN, 'integer_paths.npz'))
    img = z['puzzle_image']
    for erosion, tag in ((0.0, '0p0'), (0.07, '0p07'), (0.14, '0p14')):
        lab, grid_size = pieces.make_pieces_lab(img, 64, erosion)
        assert list(grid_size) == z[f'grid_{tag}'].tolist()
        np.testing.assert_array_equal(np.stack(lab), z[f'pieces_lab_{tag}'])       # uint8, bit-exact
        i, j = z[f'pair_entry_{tag}']
        pair = torch.stack([pieces.piece_to_tensor(lab[int(i)], 64), pieces.piece_to_tensor(lab[int(j)], 64)])
        np.testing.assert_array_equal(pair.numpy(), z[f'pair_tensor_{tag}'])       # fp32, bit-exact
    assert pieces.erosion_crop(64, 0.07) == (60, 2) and pieces.erosion_crop(64, 0.14) == (56, 4)
    with pytest.raises(ValueError):
        pieces.grid_geometry(10, 10, 64)


def test_fragment_prep_matches_reference_transform():
    """hisfrag.py:89-93 test transform = CenterCrop(512) -> ToTensor -> Normalize(.5,.5) on a PIL image."""
    from PIL import Image
    rng = np.random.default_rng(3)
    img = Image.fromarray(rng.integers(0, 256, size=(600, 700, 3), dtype=np.uint8))
    t = pieces.fragment_to_tensor(img, 512)
    arr = np.asarray(img)[44:556, 94:606].astype(np.float32) / 255.0          # centre crop offsets (600-512)/2, (700-512)/2
    want = (torch.from_numpy(arr).permute(2, 0, 1) - 0.5) / 0.5
    assert t.shape == (3, 512, 512) and torch.equal(t, want)


def test_consumer_layouts():
    from oracle import vited_oracle as orc

    class Side:
        top, right, bottom, left = 0, 1, 2, 3

    class Piece:
        def __init__(self, i):
            self.origin_piece_id = i

    logits = torch.randn(3, 3, 4, generator=torch.Generator().manual_seed(0))
    dist = grid.puzzle_distance(logits)
    fn = grid.make_distance_function(dist, Side)
    for i in range(3):
        for j in range(3):
            for si in range(4):
                for sj in range(4):
                    want = orc.puzzle_distance_lookup(logits.numpy(), i, j, si, sj)
                    got = fn(Piece(i), si, Piece(j), sj)
                    assert got == pytest.approx(want, rel=1e-6) or (got == want == float('inf'))
    upper = torch.triu(torch.randn(5, 5, generator=torch.Generator().manual_seed(1)))
    sim = grid.mirror_upper(upper)
    assert torch.equal(sim, sim.t()) and torch.equal(torch.triu(sim), upper)
    d = grid.similarity_to_distance(sim)
    assert d.dtype == np.float16 and np.array_equal(d, (1 - sim.type(torch.float16)).numpy())


class _FakeModel:
    """Stands in for the engine so the sharding / gather logic can run on CPU with gloo: score(i, j) is a closed
    form of (i, j, c)."""
    num_classes = 4

    def score_grid(self, images, mode, row_begin, row_end, out=None):
        res = self._score(images, mode, row_begin, row_end)
        if out is not None:
            out.copy_(res)
            return out
        return res

    def _score(self, images, mode, row_begin, row_end):
        n = images.shape[0]
        i = torch.arange(row_begin, row_end).view(-1, 1, 1).float()
        j = torch.arange(n).view(1, -1, 1).float()
        c = torch.arange(self.num_classes).view(1, 1, -1).float()
        full = i * 1000 + j + c / 10
        if mode == vited_b200.GRID_ORDERED_OFFDIAG:
            mask = (i != j).expand_as(full)
        else:
            mask = (j >= i).expand_as(full)
        return torch.where(mask, full, torch.zeros_like(full))


def _gloo_worker(rank, world, n, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        model = _FakeModel()
        images = torch.zeros(n, 3, 8, 8)
        puzzle = grid.score_puzzle(model, images)
        # a batch of puzzles of different sizes, (puzzle, row) units sharded over the ranks; puzzle 1 is handed over as
        # a callable and must only be materialised on ranks that own rows of it
        sizes = [5, 3, 9]
        called = []
        batch = [torch.zeros(sizes[0], 3, 8, 8), lambda: called.append(1) or torch.zeros(sizes[1], 3, 8, 8),
                 torch.zeros(sizes[2], 3, 8, 8)]
        many = grid.score_puzzles(model, batch, n_pieces=sizes)
        owns1 = any(p == 1 for p, _, _ in grid.puzzle_unit_ranges(sizes, world, rank))
        assert len(called) == (1 if owns1 else 0)
        model.num_classes = 1
        frag = grid.score_fragments(model, images)
        # the resumable form (per-rank file, rows in saved blocks) gives the same matrix on every rank
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, 'test_result_rank{rank}.pt')
            frag_blocks = grid.score_fragments(model, images, resume_path=path, block_rows=2, save_every=2)
            assert torch.equal(frag_blocks, frag)
            saved = torch.load(path.format(rank=rank))
            assert saved['is_finished'] and tuple(saved['rows']) == grid.hisfrag_row_range(n, world, rank)
        q.put((rank, puzzle.numpy(), frag.numpy(), [m.numpy() for m in many]))
    finally:
        dist.destroy_process_group()


def test_row_sharding_and_gather_world2_gloo():
    n, world = 11, 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, n, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = _FakeModel()
    single_c4 = _FakeModel()
    want_puzzle = single.score_grid(torch.zeros(n, 3, 8, 8), vited_b200.GRID_ORDERED_OFFDIAG, 0, n)
    single.num_classes = 1
    want_frag = grid.mirror_upper(single.score_grid(torch.zeros(n, 3, 8, 8), vited_b200.GRID_UPPER_TRI_DIAG, 0, n)[..., 0])
    sizes = [5, 3, 9]
    want_many = [single_c4.score_grid(torch.zeros(k, 3, 8, 8), vited_b200.GRID_ORDERED_OFFDIAG, 0, k) for k in sizes]
    for rank, puzzle, frag, many in results:
        assert np.array_equal(puzzle, want_puzzle.numpy()), f'rank {rank} puzzle grid differs'
        assert np.array_equal(frag, want_frag.numpy()), f'rank {rank} fragment grid differs'
        for got, want in zip(many, want_many):
            assert np.array_equal(got, want.numpy()), f'rank {rank}: a puzzle of the batch differs'


class _CrashingModel(_FakeModel):
    """_FakeModel that records its score_grid calls and raises once `fail_after` of them have been served."""
    num_classes = 1

    def __init__(self, fail_after=None):
        self.calls, self.fail_after = [], fail_after

    def score_grid(self, images, mode, row_begin, row_end, out=None):
        if self.fail_after is not None and len(self.calls) >= self.fail_after:
            raise RuntimeError('simulated crash')
        self.calls.append((row_begin, row_end))
        return super().score_grid(images, mode, row_begin, row_end, out)


def test_fragment_grid_crash_resume(tmp_path):
    """The crash-resume of the reference's test loop (hisfrag.py:181-195: rows already in the per-rank result file are
    skipped; :243-246: the file is rewritten every SAVE_TMP_FREQ row blocks and after the last one)."""
    n = 23
    images = torch.zeros(n, 3, 8, 8)
    want = grid.score_fragments(_CrashingModel(), images)
    path = str(tmp_path / 'test_result_rank{rank}.pt')
    crashing = _CrashingModel(fail_after=4)
    with pytest.raises(RuntimeError, match='simulated crash'):
        grid.score_fragments(crashing, images, resume_path=path, block_rows=3, save_every=2)
    assert crashing.calls == [(0, 3), (3, 6), (6, 9), (9, 12)]
    saved = torch.load(path.format(rank=0))
    # saves happen after blocks 0, 2, 4, ... and after the last: rows [0, 9) are on disk, block [9, 12) was lost
    assert saved['done'] == 9 and not saved['is_finished'] and saved['upper'].shape == (9, n)
    resumed = _CrashingModel()
    got = grid.score_fragments(resumed, images, resume_path=path, block_rows=3, save_every=2)
    assert resumed.calls[0] == (9, 12) and resumed.calls[-1] == (21, 23)
    assert torch.equal(got, want)
    assert torch.load(path.format(rank=0))['is_finished']
    # a finished file is reused without a single scoring call
    idle = _CrashingModel(fail_after=0)
    assert torch.equal(grid.score_fragments(idle, images, resume_path=path, block_rows=3), want) and idle.calls == []
    # a file of another grid is an error, not silently reused; remove_cache_file starts over
    with pytest.raises(vited_b200.VitedError, match='holds rows'):
        grid.score_fragments(_CrashingModel(), images[:20], resume_path=path)
    fresh = _CrashingModel()
    again = grid.score_fragments(fresh, images, resume_path=path, block_rows=50, remove_cache_file=True)
    assert fresh.calls == [(0, n)] and torch.equal(again, want)
    with pytest.raises(vited_b200.VitedError, match='must be positive'):
        grid.score_fragments(_CrashingModel(), images, resume_path=path, block_rows=0, remove_cache_file=True)


def test_puzzle_unit_ranges_cover_configs2():
    """BASELINE configs[2]: 20 puzzles x 1000 rows over 8 ranks -> 2,500 units each, contiguous, puzzle-major."""
    sizes = [1000] * 20
    seen = []
    for r in range(8):
        segs = grid.puzzle_unit_ranges(sizes, 8, r)
        assert sum(hi - lo for _, lo, hi in segs) == 2500
        seen += [(p, row) for p, lo, hi in segs for row in range(lo, hi)]
    assert seen == [(p, row) for p in range(20) for row in range(1000)]
    assert grid.puzzle_unit_ranges([3, 2], 4, 1) == [(0, 2, 3), (1, 0, 1)] and grid.puzzle_unit_ranges([3, 2], 4, 2) == [(1, 1, 2)]
    assert grid.puzzle_unit_ranges([3, 2], 4, 3) == [] and grid.puzzle_unit_ranges([3], 5, 4) == []


@pytest.mark.parametrize('shape', [(600, 700), (512, 512), (513, 515), (400, 700), (300, 200), (511, 1024)])
def test_center_crop_u8_matches_torchvision(shape):
    """pieces.center_crop_u8 (the byte-level crop in front of the device normalisation) against torchvision's
    CenterCrop on the PIL image, including zero padding of images smaller than the crop and odd margins."""
    from PIL import Image
    from torchvision import transforms
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    arr = rng.integers(0, 256, size=(shape[0], shape[1], 3), dtype=np.uint8)
    want = np.asarray(transforms.CenterCrop(512)(Image.fromarray(arr)))
    assert np.array_equal(pieces.center_crop_u8(arr, 512), want)


@pytest.mark.parametrize('name', ['test_patch32_64', 'small_hd64'])
def test_training_pair_construction_matches_reference(name):
    """train.prepare_pairs against the pair list the reference's own prepare_data produced (fixture
    tests/golden/train_step.npz; same seeded randperm draw over the negatives)."""
    from vited_b200 import train
    z = np.load(os.path.join(GOLDEN, 'train_step.npz'))
    torch.manual_seed(int(z[f'{name}_perm_seed']))
    groups, labels = train.prepare_pairs(z[f'{name}_targets'])
    assert np.array_equal(labels.numpy(), z[f'{name}_labels'])
    assert np.array_equal(groups[:, 0].numpy(), z[f'{name}_first'])
    from oracle import vited_oracle as orc
    og, ol = orc.train_pairs(z[f'{name}_targets'], int(z[f'{name}_perm_seed']))
    assert torch.equal(groups, og) and torch.equal(labels, ol)
    with pytest.raises(vited_b200.VitedError):                       # no CPU path: CUDA tensors only
        train.train_step(None, torch.zeros(2, 3, 64, 64), [0, 0])


def test_row_split_and_crop_geometry_properties():
    """Property tests (hypothesis) of the integer bookkeeping against the oracle restatements: the sampler's row
    boundaries for random grid sizes / world sizes, and the crop geometry for random image sizes, piece widths and
    erosion ratios (Python's banker's rounding of the crop offset included)."""
    from hypothesis import given, settings, strategies as st
    from oracle import vited_oracle as orc

    @settings(max_examples=60, deadline=None)
    @given(n=st.integers(1, 160), world=st.integers(1, 9))
    def sampler(n, world):
        first_col = grid.upper_tri_pairs(n)[:, 0]
        sizes = grid.indicates_row_ranges(first_col, world)
        assert sizes == orc.sampler_sizes(first_col, world)
        assert sizes[0] == 0 and sizes[-1] == n
        ranges = [grid.hisfrag_row_range(n, world, r) for r in range(world)]
        if all(sizes[i] <= sizes[i + 1] for i in range(len(sizes) - 1)):     # the reference can produce a backward
            rows = [x for lo, hi in ranges for x in range(lo, hi)]           # boundary for tiny grids; kept as is
            assert sorted(set(rows)) == sorted(rows) and set(rows) <= set(range(n))

    @settings(max_examples=100, deadline=None)
    @given(h=st.integers(64, 700), w=st.integers(64, 700), pw=st.sampled_from([32, 48, 64]),
           erosion=st.sampled_from([0.0, 0.03, 0.07, 0.1, 0.14, 0.2, 0.25, 0.33]))
    def geometry(h, w, pw, erosion):
        rows, cols, top, left = pieces.grid_geometry(h, w, pw)
        side, off = pieces.erosion_crop(pw, erosion)
        assert (rows, cols, top, left, side, off) == orc.crop_geometry(h, w, pw, erosion)
        assert rows >= 1 and cols >= 1 and top >= 0 and left >= 0
        assert top + rows * pw <= h and left + cols * pw <= w and 0 <= off and off + side <= pw

    sampler()
    geometry()


def test_score_puzzles_sharding_and_gather_layout_properties(monkeypatch):
    """grid.score_puzzles for random batches of puzzles and world sizes (more ranks than units, one-piece puzzles,
    puzzles split across several ranks): every rank's share is scored into its flat buffer, the all-gather is emulated
    in-process from those buffers, and every rank must end up with the single-process result."""
    from hypothesis import given, settings, strategies as st
    import torch.distributed as dist

    class _Model(_FakeModel):
        def parameters(self):            # a rank that owns no unit asks the model for its device
            return iter([torch.zeros(1)])
    model = _Model()

    @settings(max_examples=40, deadline=None)
    @given(sizes=st.lists(st.integers(1, 9), min_size=1, max_size=5), world=st.integers(1, 7))
    def run(sizes, world):
        batch = [torch.zeros(k, 3, 8, 8) for k in sizes]
        want = [model.score_grid(b, vited_b200.GRID_ORDERED_OFFDIAG, 0, b.shape[0]) for b in batch]
        # the shares partition the (puzzle, row) units
        units = [(p, r) for rank in range(world) for p, lo, hi in grid.puzzle_unit_ranges(sizes, world, rank)
                 for r in range(lo, hi)]
        assert units == [(p, r) for p, k in enumerate(sizes) for r in range(k)]
        longest = max(sum((hi - lo) * sizes[p] * 4 for p, lo, hi in grid.puzzle_unit_ranges(sizes, world, r))
                      for r in range(world))
        flats = []
        for rank in range(world):       # phase 1: every rank scores its share (no collective)
            monkeypatch.setattr(grid, '_dist_info', lambda rank=rank: (rank, world))
            flat = torch.full((longest,), float('nan'))
            part = grid.score_puzzles(model, batch, n_pieces=sizes, gather=False, blocks_out=flat)
            flats.append(flat)
            for p, lo, hi in grid.puzzle_unit_ranges(sizes, world, rank):
                blk = part[p] if (lo, hi) == (0, sizes[p]) else part[p][2]
                assert torch.equal(blk, want[p][lo:hi])
            owned = {p for p, _, _ in grid.puzzle_unit_ranges(sizes, world, rank)}
            assert all((part[p] is None) == (p not in owned) for p in range(len(sizes)))
        if world == 1:
            return

        def fake_all_gather(out, inp):   # phase 2: the one collective, emulated from the ranks' buffers
            assert inp.numel() == longest and out.numel() == world * longest
            out.copy_(torch.cat([torch.nan_to_num(f, nan=0.0) for f in flats]))
        monkeypatch.setattr(dist, 'all_gather_into_tensor', fake_all_gather)
        for rank in range(world):
            monkeypatch.setattr(grid, '_dist_info', lambda rank=rank: (rank, world))
            got = grid.score_puzzles(model, batch, n_pieces=sizes)
            for g, w in zip(got, want):
                assert torch.equal(g, w)

    run()


def test_score_fragments_sharding_and_gather_properties(monkeypatch):
    """grid.score_fragments for random grid and world sizes (more ranks than chunks included): the ranks' row blocks,
    put through an emulated all-gather of padded blocks, give every rank the single-process symmetric matrix."""
    from hypothesis import given, settings, strategies as st
    import torch.distributed as dist
    model = _FakeModel()
    model.num_classes = 1

    @settings(max_examples=40, deadline=None)
    @given(n=st.integers(2, 40), world=st.integers(2, 9))
    def run(n, world):
        sizes = grid.indicates_row_ranges(grid.upper_tri_pairs(n)[:, 0], world)
        images = torch.zeros(n, 3, 8, 8)
        if any(a > b for a, b in zip(sizes, sizes[1:])):
            # tiny grids: the reference's boundaries can go backwards -> every rank refuses, before any collective
            for rank in range(world):
                monkeypatch.setattr(grid, '_dist_info', lambda rank=rank: (rank, world))
                with pytest.raises(vited_b200.VitedError, match='cannot split'):
                    grid.score_fragments(model, images)
            return
        monkeypatch.setattr(grid, '_dist_info', lambda: (0, 1))
        want = grid.score_fragments(model, images)
        assert torch.equal(want, want.t())
        blocks = []
        for rank in range(world):
            monkeypatch.setattr(grid, '_dist_info', lambda rank=rank: (rank, world))
            blocks.append(grid.score_fragments(model, images, gather=False))
        assert sum(b.shape[0] for b in blocks) == n               # the shares partition the rows
        tallest = max(b.shape[0] for b in blocks)

        def fake_all_gather(bufs, padded):
            assert padded.shape[0] == tallest and len(bufs) == world
            for buf, b in zip(bufs, blocks):
                buf.zero_()
                buf[:b.shape[0]] = b
        monkeypatch.setattr(dist, 'all_gather', fake_all_gather)
        monkeypatch.setattr(dist, 'get_world_size', lambda *a, **k: world)
        for rank in range(world):
            monkeypatch.setattr(grid, '_dist_info', lambda rank=rank: (rank, world))
            assert torch.equal(grid.score_fragments(model, images), want)

    run()
