"""Dataset container used at the boundary.

The reference's entry points take and return ``xarray.Dataset``
(``rectify.py:54-64``, ``reproject.py:51-60``, ``affine.py:52-61``).  xarray is
an optional dependency of this build: when it is importable the entry points
accept and return ``xarray.Dataset``; otherwise (and always internally) data
travel as the small numpy-backed :class:`Dataset` / :class:`DataArray` below,
which expose the subset of the xarray surface the hot path and its tests use
(``dims``, ``shape``, ``dtype``, ``values``, ``attrs``, ``coords``,
``data_vars``, item access).
"""

from __future__ import annotations

from collections.abc import Hashable, Iterable, Mapping
from typing import Any

import numpy as np

try:  # optional
    import xarray as _xr
except Exception:  # pragma: no cover - xarray absent in the build image
    _xr = None


class DataArray:
    """A named-dimension numpy array (values are always host numpy)."""

    __slots__ = ("_data", "dims", "attrs", "name")

    def __init__(self, data: Any, dims: Iterable[Hashable] | str | None = None, attrs: Mapping | None = None,
                 name: Hashable | None = None):
        if isinstance(data, DataArray):
            dims = data.dims if dims is None else dims
            attrs = data.attrs if attrs is None else attrs
            name = data.name if name is None else name
            data = data._data
        arr = np.asarray(data)
        if dims is None:
            dims = tuple(f"dim_{k}" for k in range(arr.ndim))
        elif isinstance(dims, str):
            dims = (dims,)
        dims = tuple(dims)
        if len(dims) != arr.ndim:
            raise ValueError(f"dims {dims} do not match array of shape {arr.shape}")
        self._data = arr
        self.dims = dims
        self.attrs = dict(attrs) if attrs else {}
        self.name = name

    # xarray-like surface
    @property
    def values(self) -> np.ndarray:
        return self._data

    @property
    def data(self) -> np.ndarray:
        return self._data

    @property
    def shape(self):
        return self._data.shape

    @property
    def dtype(self):
        return self._data.dtype

    @property
    def ndim(self):
        return self._data.ndim

    @property
    def size(self):
        return self._data.size

    @property
    def sizes(self):
        return dict(zip(self.dims, self._data.shape))

    @property
    def chunks(self):
        return None

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._data, dtype=dtype)

    def __getitem__(self, key):
        out = self._data[key]
        if not isinstance(key, tuple):
            key = (key,)
        dims = []
        k = 0
        for d in self.dims:
            if k < len(key):
                kk = key[k]
                k += 1
                if isinstance(kk, (int, np.integer)):
                    continue
                if kk is Ellipsis:
                    raise NotImplementedError("Ellipsis indexing is not supported by this minimal DataArray")
            dims.append(d)
        return DataArray(out, dims=dims, attrs=self.attrs, name=self.name)

    def __repr__(self):
        return f"<b200 DataArray {self.name!r} {dict(zip(self.dims, self.shape))} {self.dtype}>"


class _VarView(Mapping):
    def __init__(self, ds: "Dataset", names):
        self._ds, self._names = ds, names

    def __getitem__(self, key):
        if key not in self._names:
            raise KeyError(key)
        return self._ds._vars[key]

    def __iter__(self):
        return iter([n for n in self._ds._vars if n in self._names])

    def __len__(self):
        return len(self._names)

    def to_dataset(self) -> "Dataset":
        return Dataset(coords={n: self._ds._vars[n] for n in self}, attrs={})


class Dataset:
    """Ordered mapping of named :class:`DataArray` split into data variables and coordinates."""

    def __init__(self, data_vars: Mapping | None = None, coords: Mapping | None = None, attrs: Mapping | None = None):
        self._vars: dict[Hashable, DataArray] = {}
        self._coord_names: set = set()
        self.attrs = dict(attrs) if attrs else {}
        for name, v in (coords or {}).items():
            self._set(name, v, is_coord=True)
        for name, v in (data_vars or {}).items():
            self._set(name, v, is_coord=False)

    def _set(self, name, v, is_coord):
        if isinstance(v, tuple):
            dims, data, *rest = v
            v = DataArray(data, dims=dims, attrs=rest[0] if rest else None, name=name)
        elif not isinstance(v, DataArray):
            arr = np.asarray(v)
            if arr.ndim == 1 and is_coord:
                v = DataArray(arr, dims=(name,), name=name)
            elif arr.ndim == 0:
                v = DataArray(arr, dims=(), name=name)
            else:
                raise ValueError(f"variable {name!r}: give a DataArray or a (dims, data) tuple")
        elif type(v) is DataArray:
            v = DataArray(v, name=name)
        else:  # a DataArray subclass (e.g. a device-resident coordinate placeholder) is kept as it is
            v.name = name
        self._vars[name] = v
        if is_coord:
            self._coord_names.add(name)
        else:
            self._coord_names.discard(name)

    # mapping surface
    def __getitem__(self, key):
        if isinstance(key, (list, tuple)):
            keep = list(key)
            return Dataset(
                data_vars={k: self._vars[k] for k in keep if k not in self._coord_names},
                coords={k: v for k, v in self._vars.items() if k in self._coord_names},
                attrs=self.attrs,
            )
        return self._vars[key]

    def __setitem__(self, key, value):
        self._set(key, value, is_coord=key in self._coord_names)

    def __contains__(self, key):
        return key in self._vars

    def __iter__(self):
        return iter(self.data_vars)

    def __getattr__(self, name):
        vars_ = self.__dict__.get("_vars")
        if vars_ is not None and name in vars_:
            return vars_[name]
        raise AttributeError(name)

    @property
    def data_vars(self) -> _VarView:
        return _VarView(self, {n for n in self._vars if n not in self._coord_names})

    @property
    def coords(self) -> _VarView:
        return _VarView(self, set(self._coord_names))

    @property
    def variables(self) -> Mapping:
        return dict(self._vars)

    def items(self):
        return [(n, self._vars[n]) for n in self.data_vars]

    @property
    def sizes(self) -> dict:
        out: dict = {}
        for v in self._vars.values():
            out.update(v.sizes)
        return out

    dims = sizes

    def drop_vars(self, names) -> "Dataset":
        if isinstance(names, (str, bytes)) or not isinstance(names, Iterable):
            names = [names]
        names = set(names)
        return Dataset(
            data_vars={k: v for k, v in self._vars.items() if k not in names and k not in self._coord_names},
            coords={k: v for k, v in self._vars.items() if k not in names and k in self._coord_names},
            attrs=self.attrs,
        )

    def assign_coords(self, coords: Mapping | None = None, **kw) -> "Dataset":
        new = self.copy()
        for k, v in {**(coords or {}), **kw}.items():
            new._set(k, v, is_coord=True)
        return new

    def copy(self) -> "Dataset":
        new = Dataset(attrs=self.attrs)
        new._vars = dict(self._vars)
        new._coord_names = set(self._coord_names)
        return new

    def __repr__(self):
        dv = ", ".join(f"{n}{tuple(v.shape)}" for n, v in self.items())
        return f"<b200 Dataset sizes={self.sizes} data_vars=[{dv}] coords={sorted(map(str, self._coord_names))}>"


# ---------------------------------------------------------------------------
# xarray bridge
# ---------------------------------------------------------------------------
def is_xarray(obj: Any) -> bool:
    return _xr is not None and isinstance(obj, _xr.Dataset)


def from_any(obj: Any) -> Dataset:
    """Normalise an input dataset (ours or xarray's) to :class:`Dataset`."""
    if isinstance(obj, Dataset):
        return obj
    if is_xarray(obj):
        coords = {n: DataArray(np.asarray(v.values), dims=v.dims, attrs=v.attrs, name=n) for n, v in obj.coords.items()}
        data_vars = {n: DataArray(np.asarray(v.values), dims=v.dims, attrs=v.attrs, name=n)
                     for n, v in obj.data_vars.items()}
        return Dataset(data_vars=data_vars, coords=coords, attrs=obj.attrs)
    raise TypeError(f"source_ds must be a Dataset, was {type(obj)}")


def to_like(result: Dataset, like: Any):
    """Return *result* as the same kind of object as the user's input."""
    if is_xarray(like):
        return _xr.Dataset(
            data_vars={n: _xr.DataArray(v.values, dims=v.dims, attrs=v.attrs) for n, v in result.items()},
            coords={n: _xr.DataArray(v.values, dims=v.dims, attrs=v.attrs) for n, v in result.coords.items()},
            attrs=result.attrs,
        )
    return result
