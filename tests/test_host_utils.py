"""Host-side option resolution and dataset helpers: the expectations of the reference's
tests/test_utils.py, restated for the numpy-backed Dataset of this package (CPU only)."""

import logging

import numpy as np
import pytest

import xcube_resampling_b200 as xrs
from xcube_resampling_b200.constants import FILLVALUE_INT, FILLVALUE_UINT8, FILLVALUE_UINT16
from xcube_resampling_b200.utils import (
    _get_agg_method,
    _get_fill_value,
    _get_grid_mapping_name,
    _get_interp_method,
    _get_recover_nan,
    _prep_interp_methods_downscale,
    _select_variables,
    clip_dataset_by_bbox,
    get_spatial_dims,
)

DA, DS = xrs.DataArray, xrs.Dataset


def _vars():
    int_var = DA(np.array([1, 2, 3], dtype=np.int32), dims=["x"])
    float_var = DA(np.array([1.0, 2.0, 3.0], dtype=np.float32), dims=["x"])
    return int_var, float_var


def test_get_spatial_dims():
    """tests/test_utils.py:29-46."""
    assert get_spatial_dims(DS(coords={"lon": ("lon", [0, 1]), "lat": ("lat", [0, 1])})) == ("lon", "lat")
    assert get_spatial_dims(DS(coords={"x": ("x", [0, 1]), "y": ("y", [0, 1])})) == ("x", "y")
    with pytest.raises(KeyError, match="No standard spatial dimensions found"):
        get_spatial_dims(DS(coords={"time": ("time", [0, 1])}))


def test_clip_dataset_by_bbox(caplog):
    """tests/test_utils.py:48-68."""
    with pytest.raises(ValueError, match="Expected bbox of length 4"):
        clip_dataset_by_bbox(DS(), bbox=[0, 0, 1])
    ds = DS({"data": (("lat", "lon"), np.array([[1, 2], [3, 4]]))}, coords={"lon": ("lon", [0, 1]), "lat": ("lat", [0, 1])})
    clipped = clip_dataset_by_bbox(ds, bbox=[1, 1, 2, 2])
    assert clipped.sizes["lat"] == 1 and clipped.sizes["lon"] == 1
    with caplog.at_level(logging.WARNING, logger="xcube.resampling"):
        clip_dataset_by_bbox(ds, bbox=[10, 10, 20, 20])
    assert any("Clipped dataset contains at least one zero-sized dimension." in m for m in caplog.messages)


def test_select_variables():
    """tests/test_utils.py:70-96."""
    ds = DS({"var1": ("x", [1, 2, 3]), "var2": ("x", [4, 5, 6]), "var3": ("x", [7, 8, 9])}, coords={"x": ("x", [0, 1, 2])})
    assert set(_select_variables(ds, variables=None).data_vars) == {"var1", "var2", "var3"}
    assert list(_select_variables(ds, variables="var1").data_vars) == ["var1"]
    res = _select_variables(ds, variables=["var1", "var3"])
    assert set(res.data_vars) == {"var1", "var3"} and "var2" not in res
    with pytest.raises(KeyError):
        _select_variables(ds, variables="nonexistent_var")


def test_get_grid_mapping_name():
    """tests/test_utils.py:98-123."""
    assert _get_grid_mapping_name(DS({"var1": ("x", [1, 2, 3])}, coords={"x": ("x", [0, 1, 2])})) is None
    ds = DS({"var1": DA([1, 2, 3], dims="x", attrs={"grid_mapping": "crs_var"})})
    assert _get_grid_mapping_name(ds) == "crs_var"
    ds = DS({"var1": ("x", [1, 2, 3]), "crs": ((), 0)}, coords={"x": ("x", [0, 1, 2])})
    assert _get_grid_mapping_name(ds) == "crs"
    ds = DS({"var1": ("x", [1, 2, 3])}, coords={"x": ("x", [0, 1, 2]), "spatial_ref": ((), 0)})
    assert _get_grid_mapping_name(ds) == "spatial_ref"
    ds = DS({"var1": DA([1, 2, 3], dims="x", attrs={"grid_mapping": "gm1"}), "crs": ((), 0)})
    with pytest.raises(AssertionError):
        _get_grid_mapping_name(ds)


def test_get_interp_method(caplog):
    """tests/test_utils.py:125-165."""
    int_var, float_var = _vars()
    assert _get_interp_method(None, "var", int_var) == 0
    assert _get_interp_method(None, "var", float_var) == 1
    assert _get_interp_method(1, "var", float_var) == 1
    assert _get_interp_method("nearest", "var", int_var) == "nearest"
    assert _get_interp_method({"var": "bilinear"}, "var", float_var) == "bilinear"
    assert _get_interp_method({np.dtype("float32"): "bilinear"}, "other", float_var) == "bilinear"
    with caplog.at_level(logging.WARNING, logger="xcube.resampling"):
        assert _get_interp_method({"something": "bilinear"}, "var", int_var) == 0
    assert any("Defaults are assigned" in m for m in caplog.messages)


def test_prep_interp_methods_downscale():
    """tests/test_utils.py:167-180."""
    assert _prep_interp_methods_downscale(None) is None
    assert _prep_interp_methods_downscale("triangular") == "bilinear"
    assert _prep_interp_methods_downscale("nearest") == "nearest"
    assert _prep_interp_methods_downscale(1) == 1
    assert _prep_interp_methods_downscale({"a": "triangular", "b": "nearest"}) == {"a": "bilinear", "b": "nearest"}
    same = {"a": "nearest", "b": "bilinear"}
    assert _prep_interp_methods_downscale(same) == same


def test_get_agg_method(caplog):
    """tests/test_utils.py:182-218 (the B200 table maps names to kernel ids, so names are compared)."""
    int_var, float_var = _vars()
    assert _get_agg_method(None, "var", int_var) == "center"
    assert _get_agg_method(None, "var", float_var) == "mean"
    assert _get_agg_method("center", "var", float_var) == "center"
    assert _get_agg_method({"var": "mean"}, "var", int_var) == "mean"
    assert _get_agg_method({np.dtype("float32"): "mean"}, "other", float_var) == "mean"
    with caplog.at_level(logging.WARNING, logger="xcube.resampling"):
        assert _get_agg_method({"something": "mean"}, "var", int_var) == "center"
    assert any("Defaults are assigned" in m for m in caplog.messages)


def test_get_recover_nan(caplog):
    """tests/test_utils.py:220-255."""
    int_var, float_var = _vars()
    assert _get_recover_nan(True, "var", int_var) is True
    assert _get_recover_nan(False, "var", float_var) is False
    assert _get_recover_nan({"var": True}, "var", int_var) is True
    assert _get_recover_nan({np.dtype("float32"): True}, "other", float_var) is True
    with caplog.at_level(logging.WARNING, logger="xcube.resampling"):
        assert _get_recover_nan({"something": True}, "var", int_var) is False
    assert any("Defaults are assigned" in m for m in caplog.messages)
    assert _get_recover_nan(None, "var", float_var) is False


def test_get_fill_value(caplog):
    """tests/test_utils.py:257-291."""
    int_var, float_var = _vars()
    uint8_var = DA(np.array([1, 2, 3], dtype=np.uint8), dims=["x"])
    uint16_var = DA(np.array([1, 2, 3], dtype=np.uint16), dims=["x"])
    assert _get_fill_value(-99, "var", int_var) == -99
    assert _get_fill_value(-9.9, "var", float_var) == -9.9
    assert _get_fill_value({"var": 1234}, "var", int_var) == 1234
    assert _get_fill_value({np.dtype("float32"): 3.14}, "other", float_var) == 3.14
    with caplog.at_level(logging.WARNING, logger="xcube.resampling"):
        assert _get_fill_value({"something": 42}, "var", int_var) == FILLVALUE_INT
    assert any("Fill value could not be derived" in m for m in caplog.messages)
    assert _get_fill_value(None, "var", uint8_var) == FILLVALUE_UINT8
    assert _get_fill_value(None, "var", uint16_var) == FILLVALUE_UINT16
    assert _get_fill_value(None, "var", int_var) == FILLVALUE_INT
    assert np.isnan(_get_fill_value(None, "var", float_var))
