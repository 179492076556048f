"""Import shim: `import diff_gaussian_rasterization` resolves here when fateavatar_b200/dropin is on sys.path
(or after fateavatar_b200.install()).  Same public names as DGR diff_gaussian_rasterization/__init__.py."""
from fateavatar_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    _RasterizeGaussians,
    rasterize_gaussians,
)
