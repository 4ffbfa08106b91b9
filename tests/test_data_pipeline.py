"""Entrance of the sampling path: the conditioning image (data/prepare_data.py:17-40, data/util.py:74-83).  CPU: the oracle's restatement of
Pillow's fixed-point bicubic resampler against PIL itself and of transform_augment against torchvision (bit-exact).  GPU: the device
version (sr3_b200.data.util.lr_to_sr_input) against the oracle, bit-exact."""
import numpy as np
import pytest
import torch

from oracle import sr3_oracle as orc


@pytest.mark.parametrize("h,w,size", [(16, 16, 128), (64, 64, 512), (16, 16, 64), (24, 16, 96), (128, 128, 16), (37, 23, 100)])
def test_oracle_bicubic_matches_pil(h, w, size):
    from PIL import Image
    rs = np.random.RandomState(h * 7 + w + size)
    for trial in range(3):
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        if trial == 1:
            img[:, : w // 2] = 255                       # saturated edge: exercises the clip to [0, 255]
            img[:, w // 2:] = 0
        ref = np.asarray(Image.fromarray(img, "RGB").resize((size, size), Image.BICUBIC))
        got = orc.pil_resize_bicubic(img, size, size)
        assert got.shape == ref.shape and np.array_equal(got, ref), (h, w, size, trial, int(np.abs(got.astype(int) - ref.astype(int)).max()))


def test_oracle_matches_torchvision_resize_and_transform():
    """What prepare_data.py / util.py literally call: trans_fn.resize(PIL image, size, BICUBIC) and ToTensor + range mapping."""
    from PIL import Image
    import torchvision
    from torchvision.transforms import functional as trans_fn
    rs = np.random.RandomState(5)
    lr = rs.randint(0, 256, (16, 16, 3)).astype(np.uint8)
    sr_pil = trans_fn.resize(Image.fromarray(lr, "RGB"), 128, Image.BICUBIC)
    assert np.array_equal(np.asarray(sr_pil), orc.pil_resize_bicubic(lr, 128, 128))
    ref = torchvision.transforms.ToTensor()(sr_pil) * (1 - (-1)) + (-1)
    got = orc.lr_to_sr_input(lr, 128, (-1, 1))
    assert torch.equal(got, ref)
    assert torch.equal(orc.lr_to_sr_input(lr, 128, (-1, 1), flip=True), ref.flip(-1))


@pytest.mark.parametrize("n_in,n_out", [(16, 128), (64, 512), (128, 16), (37, 100), (100, 37), (7, 7)])
def test_library_coefficient_tables_match_the_oracle(n_in, n_out):
    """The C++ side of libsr3_b200.so builds the same integer tables (host code: runs without a GPU)."""
    import ctypes
    from sr3_b200 import _native
    b, kk, ks = orc.pil_bicubic_tables(n_in, n_out)
    bounds = (ctypes.c_int * (2 * n_out))()
    coef = (ctypes.c_int * (n_out * 64))()
    ksize = ctypes.c_int()
    assert _native.lib().sr3_pil_bicubic_tables(n_in, n_out, bounds, coef, n_out * 64, ctypes.byref(ksize)) == 0
    assert ksize.value == ks
    assert np.array_equal(np.ctypeslib.as_array(bounds).reshape(n_out, 2), b)
    assert np.array_equal(np.ctypeslib.as_array(coef)[: n_out * ks].reshape(n_out, ks), kk)


@pytest.mark.gpu
@pytest.mark.parametrize("B,h,size,flip", [(3, 16, 128, False), (2, 64, 512, False), (4, 16, 128, True), (1, 24, 96, False)])
def test_device_lr_to_sr_input_is_bit_exact(B, h, size, flip):
    from sr3_b200.data import util
    rs = np.random.RandomState(B + h + size)
    lr = rs.randint(0, 256, (B, h, h, 3)).astype(np.uint8)
    lr[0, :, : h // 2] = 255
    got = util.lr_to_sr_input(torch.from_numpy(lr).cuda(), size, min_max=(-1, 1), hflip=flip).cpu()
    ref = torch.stack([orc.lr_to_sr_input(lr[b], size, (-1, 1), flip=flip) for b in range(B)])
    assert got.shape == ref.shape == (B, 3, size, size) and got.dtype == torch.float32
    assert torch.equal(got, ref)
    u8 = util.resize_bicubic_u8(torch.from_numpy(lr).cuda(), size).cpu().numpy()
    assert np.array_equal(u8, np.stack([orc.pil_resize_bicubic(lr[b], size, size) for b in range(B)]))
