from . import diffusion, unet  # noqa: F401
