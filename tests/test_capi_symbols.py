"""libxrs.so loads and exports every symbol include/xrs.h declares (no GPU needed)."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "xrs.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xrs_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    from xcube_resampling_b200 import build

    return build.build()


def test_header_declares_symbols():
    syms = _declared_symbols()
    assert "xrs_rectify_ij" in syms and "xrs_gather_ij" in syms and "xrs_version" in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"libxrs.so lacks {missing}"


def test_binding_table_covers_header(lib_path):
    from xcube_resampling_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    lib = _lib.load()
    assert lib.xrs_version() == 100


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np

    import xcube_resampling_b200 as xrs
    from xcube_resampling_b200._lib import XrsError

    lon = np.array([[1.0, 6.0], [0.0, 2.0]])
    lat = np.array([[56.0, 53.0], [52.0, 50.0]])
    ds = xrs.Dataset(data_vars=dict(rad=(("y", "x"), np.ones((2, 2)))),
                     coords=dict(lon=(("y", "x"), lon), lat=(("y", "x"), lat)))
    gm = xrs.GridMapping.regular((4, 4), (-1, 49), 2, "EPSG:4326")
    with pytest.raises(XrsError):
        xrs.rectify_dataset(ds, target_gm=gm, interp_methods=0)
