"""Mirror of the reference's `model` package for the hot path only (networks.define_G and sr3_modules)."""
from . import networks  # noqa: F401
