"""Import-compatible shim: ``from simple_knn._C import distCUDA2`` (lib/models/gaussian_model.py:5 of the reference)."""
from street_gaussians_b200.rasterizer import distCUDA2  # noqa: F401
