"""Import-compatible shim: ``from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer``
(lib/utils/camera_utils.py:13 of the reference) resolves to the B200-native implementation."""
from street_gaussians_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                              rasterize_gaussians)
