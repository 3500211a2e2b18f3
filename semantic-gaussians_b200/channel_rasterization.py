"""Drop-in for the reference's ``channel_rasterization`` package
(submodules/channel-rasterization/channel_rasterization/__init__.py): run-time C-channel forward
AND a C-channel backward (the reference ships the backward for C == 3 only)."""
from .rasterizer import make_module

(GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians, _RasterizeGaussians, _C) = make_module("chn")

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
