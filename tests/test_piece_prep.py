"""CPU: the numpy restatements of OpenCV's 8-bit Lab -> sRGB and Pillow's 8-bit bilinear resize (oracle) against the
libraries themselves -- cv2 / PIL / torchvision are the calls the reference makes per piece
(data/datasets/pieces_dataset.py:35-46, data/transforms.py:14-18) -- so the oracle of SURVEY 8f row 2 is pinned."""
import numpy as np
import pytest
import torch

cv2 = pytest.importorskip('cv2')


def test_lab2rgb_restatement_equals_cv2_on_cube_slices_and_noise():
    from oracle import vited_oracle as orc
    rng = np.random.default_rng(0)
    lab = rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)
    assert np.array_equal(orc.lab2rgb_u8(lab), cv2.cvtColor(lab, cv2.COLOR_LAB2RGB))
    # full a x b planes at several L, including the linear segment (L <= 20) and the extremes
    for L in (0, 1, 20, 21, 22, 127, 128, 254, 255):
        a, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing='ij')
        plane = np.stack([np.full_like(a, L), a, b], -1)
        assert np.array_equal(orc.lab2rgb_u8(plane), cv2.cvtColor(plane, cv2.COLOR_LAB2RGB)), L


@pytest.mark.parametrize('side,out', [(60, 64), (56, 64), (64, 64), (50, 64), (64, 32), (96, 64), (31, 64), (90, 40)])
def test_pil_resize_restatement_equals_pillow(side, out):
    from PIL import Image
    from oracle import vited_oracle as orc
    rng = np.random.default_rng(side * 100 + out)
    img = rng.integers(0, 256, (side, side, 3), dtype=np.uint8)
    want = np.array(Image.fromarray(img).resize((out, out), Image.BILINEAR))
    assert np.array_equal(orc.pil_resize_bilinear_u8(img, out), want)


@pytest.mark.parametrize('erosion,piece_width,img_size', [(0.07, 64, 64), (0.14, 64, 64), (0.0, 64, 64), (0.07, 48, 64)])
def test_restated_preparation_equals_reference_calls(erosion, piece_width, img_size):
    from oracle import vited_oracle as orc
    rng = np.random.default_rng(7)
    h, w = 3 * piece_width + 11, 4 * piece_width + 5          # grid does not divide the image: centred offsets
    smooth = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 3)
    for img in (rng.integers(0, 256, (h, w, 3), dtype=np.uint8), smooth):
        want = orc.prepare_pieces(img, piece_width, erosion, img_size)
        got = orc.prepare_pieces_restated(cv2.cvtColor(img, cv2.COLOR_BGR2LAB), piece_width, erosion, img_size)
        assert want.shape == (12, 3, img_size, img_size) and want.dtype == torch.float32
        assert torch.equal(got, want)
