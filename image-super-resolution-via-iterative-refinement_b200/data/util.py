"""Device-side mirror of the entrance of the sampling path: how the reference builds the conditioning image `SR` that
`GaussianDiffusion.super_resolution(x_in)` consumes.  data/prepare_data.py:17-40 upsamples the low-resolution image with
`trans_fn.resize(img, size, Image.BICUBIC)` (Pillow's two-pass fixed-point bicubic resampler) and data/util.py:74-83
(`transform_augment`) applies ToTensor, an optional horizontal flip and the [min, max] range mapping.  Here both run on the GPU, integer-exact
against Pillow, so raw 16x16 uint8 inputs can be fed to the sampler directly."""
import torch

from .. import _native


def _resize(lr_u8, size, want_u8, want_f32, min_max, hflip):
    if not (torch.is_tensor(lr_u8) and lr_u8.is_cuda and lr_u8.dtype == torch.uint8 and lr_u8.dim() == 4):
        raise _native.NativeLibraryError("expected a CUDA uint8 tensor [B, h, w, C] (HWC, as PIL / numpy images are laid out)")
    x = lr_u8.contiguous()
    B, h, w, C = x.shape
    H, W = (size, size) if isinstance(size, int) else tuple(size)
    u8 = torch.empty(B, H, W, C, dtype=torch.uint8, device=x.device) if want_u8 else None
    f32 = torch.empty(B, C, H, W, dtype=torch.float32, device=x.device) if want_f32 else None
    with torch.cuda.device(x.device):
        _native._check(_native.lib().sr3_resize_bicubic_u8(_native._ptr(x), _native._ptr(u8), _native._ptr(f32), B, h, w, C, H, W, int(bool(hflip)),
                                                           float(min_max[0]), float(min_max[1]), _native._stream()))
    return u8, f32


def resize_bicubic_u8(img_u8, size):
    """`trans_fn.resize(img, size, Image.BICUBIC)` for a batch of uint8 HWC images on the GPU -> uint8 [B, H, W, C]."""
    return _resize(img_u8, size, True, False, (0, 1), False)[0]


def lr_to_sr_input(lr_u8, size, min_max=(-1, 1), hflip=False):
    """prepare_data.py:32-34 + util.py:74-83: uint8 low-resolution images [B, h, w, 3] -> the network's conditioning input
    fp32 [B, 3, size, size] in [min, max] (bicubic upsampling, /255, optional horizontal flip, range mapping), all on the device."""
    return _resize(lr_u8, size, False, True, min_max, hflip)[1]
