"""Per-variable option resolution and small dataset helpers.

Semantics of ``xcube_resampling/utils.py:77-332`` of the reference: defaults by
dtype, lookup in a mapping first by ``str(var_name)`` then by dtype, a warning
on the ``xcube.resampling`` logger plus the default when the mapping has no
entry.
"""

from __future__ import annotations

from collections.abc import Hashable, Iterable, Mapping, Sequence

import numpy as np

from .constants import (
    AGG_CODES,
    FILLVALUE_FLOAT,
    FILLVALUE_INT,
    FILLVALUE_UINT8,
    FILLVALUE_UINT16,
    INTERP_METHOD_MAPPING,
    LOG,
)
from .dataset import DataArray, Dataset
from .gridmapping import GridMapping


def _lookup(mapping: Mapping, key: Hashable, dtype) -> object:
    value = mapping.get(str(key))
    if value is None:
        value = mapping.get(dtype)
        if value is None:  # allow np.float32 / "float32" style keys as well
            for k, v in mapping.items():
                try:
                    if not isinstance(k, str) and np.dtype(k) == np.dtype(dtype):
                        return v
                except TypeError:
                    continue
    return value


def _get_interp_method(interp_methods, key: Hashable, var: DataArray):
    """utils.py:192-215."""
    def default(dt):
        return 0 if np.issubdtype(dt, np.integer) else 1

    if isinstance(interp_methods, Mapping):
        method = _lookup(interp_methods, key, var.dtype)
        if method is None:
            LOG.warning(
                f"Interpolation method could not be derived from the mapping "
                f"`interp_methods` for data variable {key!r} with data type "
                f"{var.dtype!r}. Defaults are assigned."
            )
            method = default(var.dtype)
    elif isinstance(interp_methods, (int, str)) and not isinstance(interp_methods, bool):
        method = interp_methods
    else:
        method = default(var.dtype)
    return method


def _get_interp_method_int(interp_methods, key, var) -> int:
    """utils.py:218-227."""
    method = _get_interp_method(interp_methods, key, var)
    if isinstance(method, str):
        method = INTERP_METHOD_MAPPING[method]
    return method


def _get_interp_method_str(interp_methods, key, var) -> str:
    """utils.py:230-239."""
    method = _get_interp_method(interp_methods, key, var)
    if isinstance(method, int):
        method = INTERP_METHOD_MAPPING[method]
    return method


def _prep_interp_methods_downscale(interp_methods):
    """utils.py:242-254: the pre-downscale cannot do triangular."""
    if interp_methods == "triangular":
        return "bilinear"
    if isinstance(interp_methods, Mapping) and "triangular" in interp_methods.values():
        return {k: ("bilinear" if v == "triangular" else v) for k, v in interp_methods.items()}
    return interp_methods


def _get_agg_method(agg_methods, key: Hashable, var: DataArray) -> str:
    """utils.py:257-280; returns the method *name* (validated)."""
    def default(dt):
        return "center" if np.issubdtype(dt, np.integer) else "mean"

    if isinstance(agg_methods, Mapping):
        method = _lookup(agg_methods, key, var.dtype)
        if method is None:
            LOG.warning(
                f"Aggregation method could not be derived from the mapping `agg_methods` "
                f"for data variable {key!r} with data type {var.dtype!r}. Defaults "
                f"are assigned."
            )
            method = default(var.dtype)
    elif isinstance(agg_methods, str):
        method = agg_methods
    else:
        method = default(var.dtype)
    if method not in AGG_CODES:
        raise KeyError(method)
    return method


def _get_recover_nan(recover_nans, key: Hashable, var: DataArray) -> bool:
    """utils.py:283-302."""
    if isinstance(recover_nans, Mapping):
        value = _lookup(recover_nans, key, var.dtype)
        if value is None:
            LOG.warning(
                f"The method to recover nan could not be derived from the mapping "
                f"`recover_nans`  for data variable {key!r} with data type "
                f"{var.dtype!r}. Defaults are assigned."
            )
            value = False
    elif isinstance(recover_nans, bool):
        value = recover_nans
    else:
        value = False
    return value


def _get_fill_value(fill_values, key: Hashable, var: DataArray):
    """utils.py:305-332."""
    def default(dt):
        if dt == np.uint8:
            return FILLVALUE_UINT8
        if dt == np.uint16:
            return FILLVALUE_UINT16
        if np.issubdtype(dt, np.integer):
            return FILLVALUE_INT
        return FILLVALUE_FLOAT

    if isinstance(fill_values, Mapping):
        value = _lookup(fill_values, key, var.dtype)
        if value is None:
            LOG.warning(
                f"Fill value could not be derived from the mapping `fill_values` "
                f"for data variable {key!r} with data type {var.dtype!r}. Defaults "
                f"are assigned."
            )
            value = default(var.dtype)
    elif fill_values is not None:
        value = fill_values
    else:
        value = default(var.dtype)
    return value


def _is_equal_crs(source_gm: GridMapping, target_gm: GridMapping) -> bool:
    """utils.py:186-189: both geographic, or equal."""
    geographic = source_gm.crs.is_geographic and target_gm.crs.is_geographic
    return geographic or source_gm.crs.equals(target_gm.crs)


def _can_apply_affine_transform(source_gm: GridMapping, target_gm: GridMapping) -> bool:
    """utils.py:180-183."""
    GridMapping.assert_regular(source_gm, name="source_gm")
    GridMapping.assert_regular(target_gm, name="target_gm")
    return _is_equal_crs(source_gm, target_gm)


def _select_variables(ds: Dataset, variables: str | Iterable[str] | None = None) -> Dataset:
    """utils.py:154-160."""
    if variables is not None:
        if isinstance(variables, str):
            variables = [variables]
        ds = ds[list(variables)]
    return ds


def _get_grid_mapping_name(ds: Dataset) -> str | None:
    """utils.py:163-177."""
    names = []
    for var in ds.data_vars:
        if "grid_mapping" in ds[var].attrs:
            names.append(ds[var].attrs["grid_mapping"])
    if "crs" in ds:
        names.append("crs")
    if "spatial_ref" in ds.coords:
        names.append("spatial_ref")
    names = sorted(set(map(str, names)))
    assert len(names) <= 1, "Multiple grid mapping names found."
    return names[0] if names else None


def normalize_grid_mapping(ds: Dataset, gm: GridMapping) -> Dataset:
    """utils.py:127-151: one ``spatial_ref`` coordinate carrying the CF attributes."""
    gm_name = _get_grid_mapping_name(ds)
    if gm_name is not None and gm_name in ds:
        ds = ds.drop_vars(gm_name)
    ds = ds.assign_coords(spatial_ref=DataArray(np.array(0), dims=(), attrs=gm.crs.to_cf()))
    out = ds.copy()
    for name, var in ds.items():
        attrs = dict(var.attrs)
        attrs["grid_mapping"] = "spatial_ref"
        if getattr(var, "source", None) is not None:  # io.LazyDataArray: stays in its store
            out[name] = type(var)(var.source, dims=var.dims, attrs=attrs, name=name)
        else:
            out[name] = DataArray(var.values, dims=var.dims, attrs=attrs, name=name)
    return out


def get_spatial_dims(ds: Dataset) -> tuple[str, str]:
    """utils.py:46-74."""
    if "lat" in ds and "lon" in ds:
        return "lon", "lat"
    if "y" in ds and "x" in ds:
        return "x", "y"
    raise KeyError(
        f"No standard spatial dimensions found in dataset. "
        f"Expected pairs ('lon', 'lat') or ('x', 'y'), but found: {list(ds.sizes)}."
    )


def clip_dataset_by_bbox(ds: Dataset, bbox: Sequence[float], spatial_dims: tuple[str, str] | None = None) -> Dataset:
    """utils.py:77-124: label-based clip of 1-D spatial coordinates to a bounding box."""
    if len(bbox) != 4:
        raise ValueError(f"Expected bbox of length 4, got: {bbox}")
    if spatial_dims is None:
        spatial_dims = get_spatial_dims(ds)
    x_dim, y_dim = spatial_dims
    xv, yv = ds[x_dim].values, ds[y_dim].values
    x_sel = np.nonzero((xv >= bbox[0]) & (xv <= bbox[2]))[0]
    y_sel = np.nonzero((yv >= bbox[1]) & (yv <= bbox[3]))[0]
    xs = slice(x_sel[0], x_sel[-1] + 1) if x_sel.size else slice(0, 0)
    ys = slice(y_sel[0], y_sel[-1] + 1) if y_sel.size else slice(0, 0)

    def clip(var: DataArray) -> DataArray:
        idx = tuple(xs if d == x_dim else ys if d == y_dim else slice(None) for d in var.dims)
        return DataArray(var.values[idx], dims=var.dims, attrs=var.attrs, name=var.name)

    out = Dataset(
        data_vars={n: clip(v) for n, v in ds.items()},
        coords={n: clip(v) for n, v in ds.coords.items()},
        attrs=ds.attrs,
    )
    if any(size == 0 for size in out.sizes.values()):
        LOG.warning(
            "Clipped dataset contains at least one zero-sized dimension. "
            f"Check if the bounding box {bbox} overlaps with the dataset extent."
        )
    return out
