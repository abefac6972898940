"""The I/O edge on the host: Zarr-v2 directory stores (uncompressed or stdlib / pyarrow codecs) and .npy files as lazy band sources,
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


def _compress_store(path, compress, compressor_meta):
    """Re-encode every chunk file of an uncompressed array written by write_zarr_array."""
    import json
    import os

    for name in os.listdir(path):
        if not name.startswith("."):
            raw = open(os.path.join(path, name), "rb").read()
            open(os.path.join(path, name), "wb").write(compress(raw))
    meta = json.load(open(os.path.join(path, ".zarray")))
    meta["compressor"] = compressor_meta
    json.dump(meta, open(os.path.join(path, ".zarray"), "w"))


def _codecs():
    import bz2
    import gzip
    import lzma
    import zlib

    import pyarrow as pa

    def lz4(raw):  # numcodecs LZ4: int32 size + one raw block
        return len(raw).to_bytes(4, "little") + pa.Codec("lz4_raw").compress(raw).to_pybytes()

    return {"zlib": (lambda r: zlib.compress(r, 1), {"id": "zlib", "level": 1}),
            "gzip": (lambda r: gzip.compress(r, 1), {"id": "gzip", "level": 1}),
            "bz2": (lambda r: bz2.compress(r, 1), {"id": "bz2", "level": 1}),
            "lzma": (lzma.compress, {"id": "lzma", "format": 1, "check": -1, "preset": None, "filters": None}),
            "zstd": (lambda r: pa.Codec("zstd").compress(r).to_pybytes(), {"id": "zstd", "level": 1}),
            "lz4": (lz4, {"id": "lz4", "acceleration": 1})}


@pytest.mark.parametrize("codec", ["zlib", "gzip", "bz2", "lzma", "zstd", "lz4"])
def test_compressed_chunks_are_decoded(tmp_path, codec):
    compress, meta = _codecs()[codec]
    rng = np.random.default_rng(3)
    a = (rng.integers(0, 50, (3, 21, 34)) / 7).astype(np.float32)
    write_zarr_array(str(tmp_path / "a"), a, (2, 8, 16), ("band", "y", "x"))
    _compress_store(str(tmp_path / "a"), compress, meta)
    src = ZarrV2Source(str(tmp_path / "a"))
    assert np.array_equal(src.read_all(), a)
    out = np.empty((2, 21, 34), dtype=np.float32)
    src.read_bands(1, 2, out)
    assert np.array_equal(out, a[1:3])


def test_codec_known_answers():
    """Hand-assembled streams (RFC 8878 raw block; LZ4 block format, literals only) -- not produced by the
    libraries that decode them."""
    from xcube_resampling_b200.io import _zstd_content_size, chunk_decoder

    zstd_frame = bytes.fromhex("28b52ffd" "20" "03" "190000") + b"abc"  # single segment, FCS = 3, last raw block of 3
    assert _zstd_content_size(zstd_frame) == 3 and chunk_decoder({"id": "zstd"})(zstd_frame) == b"abc"
    # FCS field sizes: 2 bytes carry value - 256; no single-segment flag -> a window descriptor byte comes first
    assert _zstd_content_size(bytes.fromhex("28b52ffd" "40" "58" "0001")) == 256 + 256
    assert _zstd_content_size(bytes.fromhex("28b52ffd" "40" "58" "e803")) == 1000 + 256
    assert _zstd_content_size(bytes.fromhex("28b52ffd" "80" "58" "40420f00")) == 1_000_000
    assert _zstd_content_size(bytes.fromhex("28b52ffd" "00" "58" "00")) is None
    with pytest.raises(ValueError):
        _zstd_content_size(b"\x00" * 8)
    lz4_chunk = (3).to_bytes(4, "little") + bytes([0x30]) + b"abc"  # token: 3 literals, no match
    assert chunk_decoder({"id": "lz4"})(lz4_chunk) == b"abc"
    assert chunk_decoder(None) is None
    with pytest.raises(NotImplementedError, match="zarr package"):
        chunk_decoder({"id": "blosc", "cname": "lz4"})
