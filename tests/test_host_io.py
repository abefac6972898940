"""The I/O edge on the host: uncompressed Zarr-v2 directory stores and .npy files as lazy band sources,
the dataset opener, and that the option / grid-mapping plumbing keeps lazy variables lazy."""

import numpy as np
import pytest

from xcube_resampling_b200 import GridMapping
from xcube_resampling_b200.io import (
    LazyDataArray,
    NpySource,
    ZarrV2Source,
    open_zarr_dataset,
    write_zarr_array,
)
from xcube_resampling_b200.utils import normalize_grid_mapping


def test_zarr_v2_source_reads_ragged_chunks(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.random((5, 37, 53)).astype(np.float32)
    write_zarr_array(str(tmp_path / "a"), a, (2, 16, 20), ("band", "y", "x"))
    src = ZarrV2Source(str(tmp_path / "a"))
    assert src.shape == a.shape and src.dtype == np.float32 and src.ndim == 3
    out = np.empty((3, 37, 53), dtype=np.float32)
    src.read_bands(1, 3, out)
    assert np.array_equal(out, a[1:4])
    assert np.array_equal(src.read_all(), a)
    b = rng.integers(0, 200, (20, 31)).astype(np.uint8)
    write_zarr_array(str(tmp_path / "b"), b, (7, 31), ("y", "x"))
    assert np.array_equal(ZarrV2Source(str(tmp_path / "b")).read_all(), b)
    np.save(tmp_path / "c.npy", a)
    out = np.empty((2, 37, 53), dtype=np.float32)
    NpySource(str(tmp_path / "c.npy")).read_bands(3, 2, out)
    assert np.array_equal(out, a[3:5])


def test_compressed_store_is_refused(tmp_path):
    import json

    write_zarr_array(str(tmp_path / "a"), np.zeros((4, 4), dtype=np.float32), (2, 2), ("y", "x"))
    meta = json.load(open(tmp_path / "a" / ".zarray"))
    meta["compressor"] = {"id": "blosc"}
    json.dump(meta, open(tmp_path / "a" / ".zarray", "w"))
    with pytest.raises(NotImplementedError, match="zarr package"):
        ZarrV2Source(str(tmp_path / "a"))


def test_open_zarr_dataset_keeps_variables_lazy(tmp_path):
    from xcube_resampling_b200.synthetic import swath

    lon, lat = swath(40, 30, seed=1)
    data = np.random.default_rng(1).random((3, 30, 40)).astype(np.float32)
    store = tmp_path / "scene.zarr"
    write_zarr_array(str(store / "lon"), lon, (16, 16), ("y", "x"))
    write_zarr_array(str(store / "lat"), lat, (16, 16), ("y", "x"))
    write_zarr_array(str(store / "rad"), data, (1, 16, 40), ("band", "y", "x"))
    write_zarr_array(str(store / "band"), np.arange(3), (3,), ("band",))
    ds = open_zarr_dataset(str(store))
    assert isinstance(ds["rad"], LazyDataArray) and ds["rad"].shape == (3, 30, 40) and ds["rad"].dims == ("band", "y", "x")
    assert "lon" in ds.coords and np.array_equal(ds["lon"].values, lon) and np.array_equal(ds["band"].values, np.arange(3))
    gm = GridMapping.from_dataset(ds)
    assert gm.size == (40, 30) and not gm.is_regular
    ds2 = normalize_grid_mapping(ds, gm)
    assert isinstance(ds2["rad"], LazyDataArray) and ds2["rad"].attrs["grid_mapping"] == "spatial_ref"
    assert np.array_equal(ds2["rad"].values, data)  # materialises on demand
