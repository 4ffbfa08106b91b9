"""core/metrics.py exit of the sampling path: the oracle restatement against the unmodified reference (CPU, when cv2-free import is possible
it is compared live; here its make_grid restatement is checked against torchvision), and the device versions (sr3_b200.core.metrics)
against the oracle, bit for bit."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from oracle import sr3_oracle as orc


def test_oracle_tensor2img_matches_reference_formulas():
    """The reference needs cv2 just to import core/metrics.py; its tensor2img body is torch/numpy only -- re-evaluated here line by line."""
    from torchvision.utils import make_grid
    g = torch.Generator().manual_seed(0)
    for shape in [(3, 16, 12), (1, 3, 16, 12), (16, 12), (5, 3, 8, 6), (4, 1, 8, 6), (9, 3, 7, 7)]:
        t = torch.randn(*shape, generator=g) * 0.8
        ref = t.squeeze().float().cpu().clamp_(-1, 1)
        ref = (ref - (-1)) / (1 - (-1))
        if ref.dim() == 4:
            img = np.transpose(make_grid(ref, nrow=int(math.sqrt(len(ref))), normalize=False).numpy(), (1, 2, 0))
        elif ref.dim() == 3:
            img = np.transpose(ref.numpy(), (1, 2, 0))
        else:
            img = ref.numpy()
        img = (img * 255.0).round().astype(np.uint8)
        got = orc.tensor2img(t)
        assert got.dtype == np.uint8 and got.shape == img.shape and np.array_equal(got, img), shape
    a = np.random.RandomState(0).randint(0, 256, (16, 12, 3)).astype(np.uint8)
    b = np.random.RandomState(1).randint(0, 256, (16, 12, 3)).astype(np.uint8)
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    assert orc.calculate_psnr(a, b) == 20 * math.log10(255.0 / math.sqrt(mse)) and orc.calculate_psnr(a, a) == float("inf")


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 128, 128), (1, 3, 32, 32), (64, 48), (5, 3, 8, 6), (4, 1, 8, 6), (16, 3, 32, 32), (11, 3, 16, 16)])
def test_device_tensor2img_is_bit_exact(shape):
    from sr3_b200.core import metrics
    g = torch.Generator().manual_seed(sum(shape))
    t = torch.randn(*shape, generator=g) * 0.8
    t.view(-1)[:7] = torch.tensor([-1.0, 1.0, 0.0, -0.00392157, 0.00392157, 2.0, -3.0])      # edges, half-way cases, out of range
    got = metrics.tensor2img(t.cuda())
    ref = orc.tensor2img(t)
    assert got.dtype == np.uint8 and got.shape == ref.shape, (got.shape, ref.shape)
    assert np.array_equal(got, ref)


@pytest.mark.gpu
def test_device_psnr_matches_oracle():
    from sr3_b200.core import metrics
    rs = np.random.RandomState(3)
    a = rs.randint(0, 256, (128, 128, 3)).astype(np.uint8)
    b = np.clip(a.astype(np.int32) + rs.randint(-9, 10, a.shape), 0, 255).astype(np.uint8)
    assert metrics.calculate_psnr(a, b) == orc.calculate_psnr(a, b)
    assert metrics.calculate_psnr(a, a) == float("inf")
    assert metrics.calculate_psnr(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) == orc.calculate_psnr(a, b)
