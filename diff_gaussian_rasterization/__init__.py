"""Drop-in for the `diff_gaussian_rasterization` package that latentSplat imports at
/root/reference/src/model/decoder/cuda_splatting.py:6-9 (Chrixtar/latent-gaussian-rasterization,
requirements.txt:34).  Backed by latentsplat_b200's sm_100a rasterizer (libls_raster.so).
"""
from latentsplat_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer"]
