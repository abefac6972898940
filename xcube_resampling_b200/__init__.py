"""B200-native spatial resampling: drop-in for xcube-resampling's hot path.

Entry points (same names, arguments and error behaviour as the reference):
``resample_in_space``, ``rectify_dataset``, ``reproject_dataset``,
``affine_transform_dataset`` and the ``GridMapping`` type.  All computation
runs in ``libxrs.so`` (hand-written CUDA for sm_100a, C ABI in
``include/xrs.h``); torch tensors serve as device buffers only.  There is no
CPU fallback.
"""

from .constants import LOG, SCALE_LIMIT, UV_DELTA  # noqa: F401
from .crs import CRS, CRS_CRS84, CRS_WGS84  # noqa: F401
from .dataset import DataArray, Dataset  # noqa: F401
from .gridmapping import GridMapping  # noqa: F401
from .version import version as __version__  # noqa: F401


def __getattr__(name):
    # the entry points import torch; keep `import xcube_resampling_b200` light
    if name == "rectify_dataset":
        from .rectify import rectify_dataset
        return rectify_dataset
    if name == "affine_transform_dataset":
        from .affine import affine_transform_dataset
        return affine_transform_dataset
    if name == "reproject_dataset":
        from .reproject import reproject_dataset
        return reproject_dataset
    if name == "resample_in_space":
        from .spatial import resample_in_space
        return resample_in_space
    raise AttributeError(name)
