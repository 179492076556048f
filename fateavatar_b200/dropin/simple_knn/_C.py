"""Import shim for `from simple_knn._C import distCUDA2` (volume_rendering/gaussian_model.py:25)."""
from fateavatar_b200.knn import distCUDA2  # noqa: F401
