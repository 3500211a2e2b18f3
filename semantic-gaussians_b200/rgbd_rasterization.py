"""Drop-in for the reference's ``rgbd_rasterization`` package
(submodules/rgbd-rasterization/rgbd_rasterization/__init__.py): RGB + median depth."""
from .rasterizer import make_module

(GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians, _RasterizeGaussians, _C) = make_module("rgbd")

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
