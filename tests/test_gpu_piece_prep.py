"""GPU: on-device piece preparation (C-ABI vited_prepare_pieces) against the reference's own per-piece library calls
(cv2 LAB2RGB, PIL bilinear Resize, ToTensor, Normalize) -- every output float bit-identical."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip('cv2')


@pytest.mark.parametrize('erosion,piece_width,img_size,grid', [
    (0.07, 64, 64, (5, 7)), (0.14, 64, 64, (4, 4)), (0.0, 64, 64, (3, 5)), (0.07, 48, 64, (4, 6)), (0.25, 96, 64, (2, 3)),
    (0.0, 96, 64, (2, 2)), (0.1, 64, 32, (3, 3))])
def test_device_preparation_is_bit_exact(erosion, piece_width, img_size, grid):
    from oracle import vited_oracle as orc
    from vited_b200 import pieces
    rng = np.random.default_rng(int(erosion * 100) + piece_width + img_size)
    h, w = grid[0] * piece_width + 13, grid[1] * piece_width + 6
    noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    smooth = cv2.GaussianBlur(noise, (0, 0), 4)
    for img in (noise, smooth):
        got, shape = pieces.prepare_pieces_device(img, piece_width, erosion, img_size)
        want = orc.prepare_pieces(img, piece_width, erosion, img_size)
        assert shape == grid and got.shape == want.shape
        assert torch.equal(got.cpu(), want)


def test_every_lab_triple_converts_like_cv2():
    """All 2^24 (L, a, b) triples through the kernel (a 4096 x 4096 'image' cut into 64-pixel pieces, no erosion, no
    resize) against cv2.cvtColor on the same array."""
    import ctypes
    from vited_b200 import _lib
    L, a, b = np.meshgrid(*[np.arange(256, dtype=np.uint8)] * 3, indexing='ij')
    lab = np.ascontiguousarray(np.stack([L, a, b], -1).reshape(4096, 4096, 3))
    want = cv2.cvtColor(lab, cv2.COLOR_LAB2RGB)
    lab_d = torch.from_numpy(lab).cuda()
    out = torch.empty((4096, 3, 64, 64), dtype=torch.float32, device='cuda')
    st = _lib.lib.vited_prepare_pieces(ctypes.c_void_p(lab_d.data_ptr()), 4096, 4096, 64, 64, 0, 64,
                                       ctypes.c_void_p(out.data_ptr()), None,
                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(st, 'vited_prepare_pieces')
    rgb = torch.round((out * 0.5 + 0.5) * 255).to(torch.uint8)                       # exact inverse of the fp32 steps
    rgb = rgb.view(64, 64, 3, 64, 64).permute(0, 3, 1, 4, 2).reshape(4096, 4096, 3).cpu().numpy()
    assert np.array_equal(rgb, want)


def test_prepared_pieces_feed_the_grid_like_host_prepared_ones():
    from tests import helpers
    from vited_b200 import grid, pieces, synthetic
    z, kw = helpers.load_model_case('small_hd32')
    model, _ = helpers.make_gpu_model(kw, 1)
    img = synthetic.synthetic_puzzle_image(3, 4, 64, seed=2)
    dev_batch, shape = pieces.prepare_pieces_device(img, 64, 0.07, kw['img_size'])
    lab_pieces, shape2 = pieces.make_pieces_lab(img, 64, 0.07)
    host_batch = pieces.pieces_to_batch(lab_pieces, kw['img_size'])
    assert shape == shape2 and torch.equal(dev_batch.cpu(), host_batch)
    assert torch.equal(grid.score_puzzle(model, dev_batch), grid.score_puzzle(model, host_batch.cuda()))


def test_bad_arguments_fail_loudly():
    import vited_b200
    from vited_b200 import pieces
    with pytest.raises((vited_b200.VitedError, ValueError)):
        pieces.prepare_pieces_device(np.zeros((32, 32, 3), np.uint8), 64, 0.07, 64)
    with pytest.raises(vited_b200.VitedError):
        pieces.prepare_pieces_device(np.zeros((256, 256, 3), np.uint8), 128, 0.0, 64)    # side 128 > kernel limit


def test_fragment_batch_on_device_equals_reference_transform():
    """Hisfrag test transform (hisfrag.py:89-93): CenterCrop(512) -> ToTensor -> Normalize, images of different sizes
    (one smaller than the crop), against torchvision on the PIL images."""
    from PIL import Image
    from vited_b200 import pieces
    rng = np.random.default_rng(12)
    imgs = [Image.fromarray(rng.integers(0, 256, size=s + (3,), dtype=np.uint8))
            for s in [(600, 700), (512, 512), (777, 513), (300, 900)]]
    got = pieces.fragments_to_batch_device(imgs, 512)
    want = torch.stack([pieces.fragment_to_tensor(im, 512) for im in imgs])
    assert got.shape == (4, 3, 512, 512) and torch.equal(got.cpu(), want)
    small = pieces.fragments_to_batch_device([np.asarray(imgs[0])[:70, :90]], 64)
    assert torch.equal(small.cpu(), pieces.fragment_to_tensor(Image.fromarray(np.asarray(imgs[0])[:70, :90]), 64)[None])
