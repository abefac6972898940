"""The I/O edge on the host: Zarr-v2 directory stores (uncompressed or stdlib / pyarrow codecs) and .npy files as lazy band sources,
the dataset opener, and that the option / grid-mapping plumbing keeps lazy variables lazy."""

import numpy as np
import pytest

from xcube_resampling_b200 import GridMapping
from xcube_resampling_b200.io import (
    NetCDF3Source,
    open_netcdf_dataset,
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
    meta["compressor"] = {"id": "jpeg2k"}
    json.dump(meta, open(tmp_path / "a" / ".zarray", "w"))
    with pytest.raises(NotImplementedError, match="zarr package"):
        ZarrV2Source(str(tmp_path / "a"))
    meta["compressor"], meta["filters"] = None, [{"id": "delta", "dtype": "<f4"}]
    json.dump(meta, open(tmp_path / "a" / ".zarray", "w"))
    with pytest.raises(NotImplementedError, match="filtered"):
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
        chunk_decoder({"id": "jpeg2k"})


def _blosc_frame(raw: bytes, typesize: int, blocksize: int, shuffle: bool, cname: str, dont_split: bool = False,
                 store_streams: bool = False) -> bytes:
    """A Blosc-1 frame assembled from the container description in ``io.blosc_decompress`` (16-byte header,
    block offsets, per-block streams each prefixed by its int32 size); inner streams from zlib / pyarrow."""
    import zlib

    import pyarrow as pa

    fmt, enc = {"lz4": (1, lambda r: pa.Codec("lz4_raw").compress(r).to_pybytes()),
                "snappy": (2, lambda r: pa.Codec("snappy").compress(r).to_pybytes()),
                "zlib": (3, lambda r: zlib.compress(r, 1)),
                "zstd": (4, lambda r: pa.Codec("zstd").compress(r).to_pybytes())}[cname]
    flags = (fmt << 5) | (0x10 if dont_split else 0) | (0x1 if shuffle else 0)
    nblocks = -(-len(raw) // blocksize)
    body, starts = b"", []
    for b in range(nblocks):
        block = np.frombuffer(raw[b * blocksize:(b + 1) * blocksize], dtype=np.uint8)
        if shuffle and typesize > 1:
            n = len(block) // typesize
            block = np.concatenate([block[:n * typesize].reshape(n, typesize).T.reshape(-1), block[n * typesize:]])
        split = (not dont_split and typesize <= 16 and blocksize // typesize >= 128 and len(block) == blocksize)
        nstreams = typesize if split else 1
        ne = len(block) // nstreams
        starts.append(16 + 4 * nblocks + len(body))
        for k in range(nstreams):
            part = block[k * ne:(k + 1) * ne].tobytes()
            comp = enc(part)
            if store_streams or len(comp) >= ne:  # stored verbatim: size field == decoded size
                comp = part
            assert len(comp) != ne or comp == part
            body += len(comp).to_bytes(4, "little") + comp
    head = bytes([2, 1, flags, typesize]) + len(raw).to_bytes(4, "little") + blocksize.to_bytes(4, "little")
    total = 16 + 4 * nblocks + len(body)
    return head + total.to_bytes(4, "little") + b"".join(s.to_bytes(4, "little") for s in starts) + body


def test_blosc_known_answers():
    """Frames written out byte by byte from c-blosc's published container layout."""
    from xcube_resampling_b200.io import blosc_decompress, chunk_decoder

    assert chunk_decoder({"id": "blosc", "cname": "lz4", "clevel": 5, "shuffle": 1, "blocksize": 0}) is blosc_decompress
    # flag bit 1: the payload is a plain copy, no block offsets
    stored = bytes.fromhex("02" "01" "02" "01" "03000000" "03000000" "13000000") + b"abc"
    assert blosc_decompress(stored) == b"abc"
    # LZ4 (format 1 in bits 5-7), one block, one stream of 4 bytes: token 0x30 = 3 literals, no match
    lz4 = bytes.fromhex("02" "01" "20" "01" "03000000" "03000000" "1c000000" "14000000" "04000000" "30") + b"abc"
    assert blosc_decompress(lz4) == b"abc"
    # byte shuffle (bit 0), typesize 2, two elements 0x0102, 0x0304: the block holds the low bytes, then the high
    # bytes; the stream's size field equals its decoded size, so it is stored verbatim
    shuf = bytes.fromhex("02" "01" "21" "02" "04000000" "04000000" "1c000000" "14000000" "04000000" "02040103")
    assert np.array_equal(np.frombuffer(blosc_decompress(shuf), dtype="<u2"), [0x0102, 0x0304])
    # a trailing byte beyond the last whole element is not shuffled
    shuf5 = bytes.fromhex("02" "01" "21" "02" "05000000" "05000000" "1d000000" "14000000" "05000000" "02040103" "ff")
    assert blosc_decompress(shuf5) == bytes.fromhex("02010403ff")
    assert blosc_decompress(bytes.fromhex("02" "01" "21" "04" "00000000" "00000000" "10000000")) == b""
    with pytest.raises(NotImplementedError, match="bit-shuffled"):
        blosc_decompress(bytes.fromhex("02" "01" "24" "01" "03000000" "03000000" "1c000000" "14000000" "04000000" "30")
                         + b"abc")
    with pytest.raises(NotImplementedError, match="BloscLZ"):
        blosc_decompress(bytes.fromhex("02" "01" "00" "01" "03000000" "03000000" "1c000000" "14000000" "04000000" "30")
                         + b"abc")
    with pytest.raises(ValueError):
        blosc_decompress(b"\x02\x01")
    with pytest.raises(ValueError, match="outside"):
        blosc_decompress(bytes.fromhex("02" "01" "20" "01" "03000000" "03000000" "1c000000" "14000000" "40000000" "30")
                         + b"abc")


@pytest.mark.parametrize("cname", ["lz4", "snappy", "zlib", "zstd"])
@pytest.mark.parametrize("layout", ["split", "dont_split", "no_shuffle", "stored_streams"])
def test_blosc_frames_round_trip(cname, layout):
    """Several blocks, a partial last block, typesize streams per block when splitting applies."""
    from xcube_resampling_b200.io import blosc_decompress

    rng = np.random.default_rng(11)
    raw = (rng.integers(0, 40, 1500) / 8).astype(np.float32).tobytes() + b"\x07\x09\x0b"  # 6003 bytes
    frame = _blosc_frame(raw, 4, 2048, shuffle=layout != "no_shuffle", cname=cname, dont_split=layout == "dont_split",
                         store_streams=layout == "stored_streams")
    assert blosc_decompress(frame) == raw
    assert blosc_decompress(frame + b"padding") == raw  # cbytes, not the buffer length, bounds the frame


def test_blosc_compressed_store(tmp_path):
    """A store as xarray's ``to_zarr`` writes it by default (Blosc, LZ4, byte shuffle), ragged chunks included."""
    rng = np.random.default_rng(5)
    a = (rng.integers(0, 50, (3, 40, 70)) / 7).astype(np.float32)
    write_zarr_array(str(tmp_path / "a"), a, (2, 16, 32), ("band", "y", "x"))
    _compress_store(str(tmp_path / "a"), lambda r: _blosc_frame(r, 4, 1024, True, "lz4"),
                    {"id": "blosc", "cname": "lz4", "clevel": 5, "shuffle": 1, "blocksize": 0})
    src = ZarrV2Source(str(tmp_path / "a"))
    assert np.array_equal(src.read_all(), a)
    out = np.empty((2, 40, 70), dtype=np.float32)
    src.read_bands(1, 2, out)
    assert np.array_equal(out, a[1:3])


def test_open_zarr_dataset_on_a_blosc_store(tmp_path):
    """Every array of the store is compressed, the 1-D axis coordinates and the 0-D ``spatial_ref`` included."""
    import json

    from xcube_resampling_b200.synthetic import swath

    lon, lat = swath(40, 30, seed=2)
    data = np.random.default_rng(2).random((3, 30, 40)).astype(np.float32)
    store = tmp_path / "scene.zarr"
    write_zarr_array(str(store / "lon"), lon, (16, 16), ("y", "x"))
    write_zarr_array(str(store / "lat"), lat, (16, 16), ("y", "x"))
    write_zarr_array(str(store / "rad"), data, (1, 16, 40), ("band", "y", "x"))
    write_zarr_array(str(store / "band"), np.arange(7, dtype=np.int64), (4,), ("band7",))
    write_zarr_array(str(store / "spatial_ref"), np.asarray(5, dtype=np.int64), (), ())
    meta = {"id": "blosc", "cname": "zstd", "clevel": 3, "shuffle": 1, "blocksize": 0}
    for name, typesize in (("lon", 8), ("lat", 8), ("rad", 4), ("band", 8), ("spatial_ref", 8)):
        _compress_store(str(store / name), lambda r, t=typesize: _blosc_frame(r, t, 4096, True, "zstd"), meta)
    ds = open_zarr_dataset(str(store))
    assert np.array_equal(ds["band"].values, np.arange(7)) and ds["spatial_ref"].values == 5
    assert np.array_equal(ds["lon"].values, lon) and np.array_equal(ds["lat"].values, lat)
    assert isinstance(ds["rad"], LazyDataArray) and np.array_equal(ds["rad"].values, data)
    # a missing chunk of a 1-D array is fill value; "NaN" is JSON's spelling of the float special
    write_zarr_array(str(store / "t"), np.arange(6, dtype=np.float64), (4,), ("t",))
    (store / "t" / "1").unlink()
    m = json.load(open(store / "t" / ".zarray"))
    m["fill_value"] = "NaN"
    json.dump(m, open(store / "t" / ".zarray", "w"))
    t = open_zarr_dataset(str(store))["t"].values
    assert np.array_equal(t[:4], np.arange(4)) and np.isnan(t[4:]).all() and t.shape == (6,)


def test_open_netcdf_classic_dataset(tmp_path):
    """A NetCDF classic (64-bit offset) file written by scipy: big-endian on disk, lazy 3-D variable over the
    memory map, coordinates / scalar grid-mapping variable / attributes as a Dataset the CF discovery accepts."""
    from scipy.io import netcdf_file

    from xcube_resampling_b200.synthetic import swath

    lon, lat = swath(40, 30, seed=3)
    rad = np.random.default_rng(3).random((3, 30, 40)).astype(np.float32)
    cls = np.random.default_rng(4).integers(0, 200, (30, 40)).astype(np.int16)
    path = str(tmp_path / "scene.nc")
    f = netcdf_file(path, "w", version=2)
    for name, n in (("band", 3), ("y", 30), ("x", 40)):
        f.createDimension(name, n)
    f.title = "scene"
    for name, typ, dims, values in (("rad", "f4", ("band", "y", "x"), rad), ("cls", "i2", ("y", "x"), cls),
                                    ("lon", "f8", ("y", "x"), lon), ("lat", "f8", ("y", "x"), lat),
                                    ("band", "i4", ("band",), np.arange(3))):
        v = f.createVariable(name, typ, dims)
        v[:] = values
    f.variables["rad"].units = "mW.m-2.sr-1.nm-1"
    f.close()
    ds = open_netcdf_dataset(path)
    assert ds.attrs == {"title": "scene"} and ds["rad"].attrs == {"units": "mW.m-2.sr-1.nm-1"}
    assert isinstance(ds["rad"], LazyDataArray) and isinstance(ds["rad"].source, NetCDF3Source)
    assert ds["rad"].dims == ("band", "y", "x") and ds["rad"].dtype == np.float32 and ds["rad"].dtype.isnative
    assert np.array_equal(ds["rad"].values, rad) and np.array_equal(ds["cls"].values, cls)
    out = np.empty((2, 30, 40), dtype=np.float32)
    ds["rad"].source.read_bands(1, 2, out)
    assert np.array_equal(out, rad[1:])
    assert "lon" in ds.coords and ds["lon"].dtype.isnative and np.array_equal(ds["lon"].values, lon)
    assert np.array_equal(ds["band"].values, np.arange(3))
    gm = GridMapping.from_dataset(ds)
    assert gm.size == (40, 30) and not gm.is_regular
    (tmp_path / "x.nc").write_bytes(b"\x89HDF\r\n\x1a\n" + bytes(64))
    with pytest.raises(NotImplementedError, match="NetCDF-4"):
        open_netcdf_dataset(str(tmp_path / "x.nc"))


def test_mask_and_scale_decoding_of_lazy_variables(tmp_path):
    """CF packing (Sentinel-3 radiances: uint16 + scale_factor + _FillValue) decoded on the reader thread:
    fill -> NaN, then raw * scale_factor + add_offset in float32; int32 data decode to float64; variables
    without coding attributes and the default (mask_and_scale=False) are untouched."""
    import json

    from xcube_resampling_b200.io import DecodedSource, cf_decoded_dtype

    rng = np.random.default_rng(9)
    rad = rng.integers(0, 65535, (3, 20, 30)).astype(np.uint16)
    rad[0, 0, :5] = 65535
    cnt = rng.integers(-1000, 1000, (20, 30)).astype(np.int32)
    cnt[3, 4] = -999
    plain = rng.random((20, 30)).astype(np.float32)
    store = tmp_path / "s.zarr"
    write_zarr_array(str(store / "rad"), rad, (1, 8, 16), ("band", "y", "x"))
    write_zarr_array(str(store / "cnt"), cnt, (8, 16), ("y", "x"))
    write_zarr_array(str(store / "plain"), plain, (8, 16), ("y", "x"))
    for name, extra in (("rad", dict(scale_factor=0.01, add_offset=1.5, _FillValue=65535, units="W")),
                        ("cnt", dict(scale_factor=0.5, missing_value=-999))):
        attrs = json.load(open(store / name / ".zattrs"))
        json.dump({**attrs, **extra}, open(store / name / ".zattrs", "w"))
    ds = open_zarr_dataset(str(store), mask_and_scale=True)
    got = ds["rad"].values
    assert ds["rad"].dtype == np.float32 and ds["rad"].attrs == {"units": "W"} and isinstance(ds["rad"].source, DecodedSource)
    assert np.isnan(got[0, 0, :5]).all() and np.isnan(got).sum() == (rad == 65535).sum()
    ok = rad != 65535
    assert np.array_equal(got[ok], rad[ok].astype(np.float32) * np.float32(0.01) + np.float32(1.5))
    assert np.allclose(got[ok], rad[ok] * 0.01 + 1.5, rtol=1e-6)
    part = np.empty((2, 20, 30), dtype=np.float32)
    ds["rad"].source.read_bands(1, 2, part)
    assert np.array_equal(part, got[1:], equal_nan=True)
    c = ds["cnt"].values
    assert c.dtype == np.float64 and np.isnan(c[3, 4]) and np.array_equal(np.delete(c.ravel(), 3 * 30 + 4),
                                                                          np.delete(cnt.ravel(), 3 * 30 + 4) * 0.5)
    assert ds["plain"].dtype == np.float32 and not isinstance(ds["plain"].source, DecodedSource)
    raw = open_zarr_dataset(str(store))
    assert raw["rad"].dtype == np.uint16 and raw["rad"].attrs["scale_factor"] == 0.01 and np.array_equal(raw["rad"].values, rad)
    assert cf_decoded_dtype(np.dtype("f4"), {"_FillValue": -1.0}) == np.float32
    assert cf_decoded_dtype(np.dtype("i8"), {"add_offset": 1}) == np.float64 and cf_decoded_dtype(np.dtype("u1"), {}) is None


def test_mask_and_scale_fill_value_spellings():
    """Fill values as lists, numpy scalars, out-of-range numbers and JSON's "NaN" string."""
    from xcube_resampling_b200.io import DecodedSource, LazySource

    class Mem(LazySource):
        def __init__(self, a):
            self.a, self.shape, self.dtype = a, a.shape, a.dtype

        def read_bands(self, b0, nb, out):
            out[:nb] = self.a[b0:b0 + nb]

    a = np.array([[[1, 2], [65535, 4]]], dtype=np.uint16)
    nan = np.nan
    for attrs, want in (({"_FillValue": 65535}, [1, 2, nan, 4]), ({"_FillValue": np.uint16(65535)}, [1, 2, nan, 4]),
                        ({"_FillValue": [65535, 2]}, [1, nan, nan, 4]),
                        ({"_FillValue": "NaN", "scale_factor": 2.0}, [2, 4, 131070, 8]),
                        ({"missing_value": 70000.0, "_FillValue": -1}, [1, 2, 65535, 4])):
        got = DecodedSource(Mem(a), attrs).read_all()
        assert got.dtype == np.float32 and np.array_equal(got.ravel(), np.float32(want), equal_nan=True), attrs
    f = np.array([[[1.5, nan], [-999.0, 4]]], dtype=np.float64)
    got = DecodedSource(Mem(f), {"_FillValue": nan, "missing_value": -999.0}).read_all()
    assert got.dtype == np.float64 and np.array_equal(got.ravel(), [1.5, nan, nan, 4], equal_nan=True)
