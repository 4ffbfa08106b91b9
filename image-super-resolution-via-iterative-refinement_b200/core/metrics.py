"""Device-side mirror of the two pieces of the reference's core/metrics.py that sit on the sampling path's exit (sr.py / infer.py call them
on every snapshot): `tensor2img` (core/metrics.py:8-34) and `calculate_psnr` (:42-50).  Same signatures and results; the conversion to
uint8 happens on the GPU, so a quarter of the bytes cross PCIe and the host never touches fp32 images."""
import ctypes
import math

import numpy as np
import torch

from .. import _native


def tensor2img(tensor, out_type=np.uint8, min_max=(-1, 1)):
    """4D (B,C,H,W), 3D (C,H,W) or 2D (H,W) CUDA tensor -> numpy HWC (or HW) uint8 image, RGB order; 4-D input is tiled like
    torchvision.utils.make_grid(nrow=int(sqrt(B))), exactly as the reference does."""
    if out_type != np.uint8:
        raise NotImplementedError("sr3_b200.core.metrics.tensor2img produces uint8 images only")
    if not tensor.is_cuda:
        raise _native.NativeLibraryError("tensor2img expects a CUDA tensor (there is no CPU path)")
    t = tensor.squeeze().float().contiguous()
    n_dim = t.dim()
    if n_dim == 4:
        n, C, H, W = t.shape
        nrow = int(math.sqrt(n))
    elif n_dim == 3:
        n, (C, H, W), nrow = 1, t.shape, 1
    elif n_dim == 2:
        n, C, (H, W), nrow = 1, 1, t.shape, 1
    else:
        raise TypeError('Only support 4D, 3D and 2D tensor. But received with dimension: {:d}'.format(n_dim))
    if n > 1:
        ncol = min(nrow, n)
        rows = (n + ncol - 1) // ncol
        GH, GW = rows * (H + 2) + 2, ncol * (W + 2) + 2
        if C == 1:                                   # make_grid turns single-channel images into 3 channels
            t = t.expand(n, 3, H, W).contiguous()
            C = 3
    else:
        GH, GW = H, W
    out = torch.empty(GH, GW, C, dtype=torch.uint8, device=t.device)
    with torch.cuda.device(t.device):
        _native._check(_native.lib().sr3_tensor2img(_native._ptr(t), _native._ptr(out), n, C, H, W, max(nrow, 1), float(min_max[0]), float(min_max[1]),
                                                    _native._stream()))
    img = out.cpu().numpy()
    return img[:, :, 0] if n_dim == 2 else img


def calculate_psnr(img1, img2):
    """uint8 images (numpy arrays or CUDA uint8 tensors of the same shape) -> PSNR in dB, float64 arithmetic as the reference."""
    a = torch.as_tensor(img1).to("cuda", torch.uint8).contiguous() if not (torch.is_tensor(img1) and img1.is_cuda) else img1.contiguous()
    b = torch.as_tensor(img2).to(a.device, torch.uint8).contiguous() if not (torch.is_tensor(img2) and img2.is_cuda) else img2.contiguous()
    assert a.shape == b.shape and a.dtype == torch.uint8 and b.dtype == torch.uint8
    ssd = ctypes.c_uint64()
    with torch.cuda.device(a.device):
        _native._check(_native.lib().sr3_ssd_u8(_native._ptr(a), _native._ptr(b), a.numel(), ctypes.byref(ssd), _native._stream()))
    mse = ssd.value / float(a.numel())
    if mse == 0:
        return float('inf')
    return 20 * math.log10(255.0 / math.sqrt(mse))
