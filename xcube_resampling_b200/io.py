"""The I/O edge: chunked arrays on disk -> page-locked host staging -> H2D, overlapped.

The reference opens Zarr / NetCDF through xarray + dask and lets the dask chunks flow into its tile
tasks.  Here a data variable may be a :class:`LazyDataArray` whose bands live in a chunked store;
the band-chunk pipeline (``_pipeline.GatherPipeline``) then reads band chunk k+1 from the store into
one of two page-locked staging buffers on a reader thread while chunk k is being copied to the
device and chunk k-1 is being gathered -- decode / read, PCIe upload, kernels and download all
overlap, and the host never holds more than two chunks of a variable.

Stores understood without third-party packages:

* :class:`ZarrV2Source` -- a Zarr-v2 array in a directory store (C order, any chunking): what
  ``add_spatial_ref`` (``cfconv.py``) and the reference's Zarr helpers operate on.  Chunks may be
  uncompressed (``"compressor": null``: read straight into the staging buffer) or compressed with a
  codec whose stream format the standard library or pyarrow decodes (numcodecs ids ``zlib``, ``gzip``,
  ``bz2``, ``lzma``, ``zstd``, ``lz4``, and ``blosc`` -- zarr's default -- with an LZ4 / LZ4HC / Snappy /
  Zlib / Zstd inner codec, byte shuffle or none: :func:`blosc_decompress`), decoded on the reader thread.
  BloscLZ or bit-shuffled Blosc frames and Zarr filters need the ``zarr`` / ``numcodecs`` packages, which
  this build does not have; open such stores with xarray and pass the arrays instead.
* :class:`NpySource` -- a ``.npy`` file, memory-mapped.
* :class:`NetCDF3Source` -- a variable of a NetCDF classic / 64-bit-offset file (CDF-1 / CDF-2), memory-mapped
  through ``scipy.io.netcdf_file``; its big-endian values are byte-swapped by the copy into the staging
  buffer.  NetCDF-4 files are HDF5 containers and need ``netCDF4`` / ``h5py``, which this build does not
  have; open them with xarray and pass the arrays instead.

:func:`open_zarr_dataset` / :func:`open_netcdf_dataset` assemble a :class:`~xcube_resampling_b200.dataset.Dataset` from a
directory store: coordinate arrays (1-D, or 2-D named like coordinates) are read eagerly, every
other array becomes a lazy variable.
"""

from __future__ import annotations

import json
import os
from collections.abc import Mapping

import numpy as np

from .dataset import DataArray, Dataset


def _pyarrow_codec(name: str):
    try:
        import pyarrow as pa
    except ImportError as e:  # pragma: no cover - pyarrow is part of the image
        raise NotImplementedError(f"{name}-compressed Zarr chunks need pyarrow (or the zarr package)") from e
    return pa.Codec(name)


def chunk_decoder(compressor: dict | None):
    """``bytes -> bytes`` for a Zarr-v2 ``compressor`` entry (numcodecs configuration), ``None`` for
    uncompressed chunks; ``NotImplementedError`` for codecs this build cannot decode."""
    if compressor is None:
        return None
    cid = compressor.get("id")
    if cid == "zlib":
        import zlib

        return zlib.decompress
    if cid == "gzip":
        import gzip

        return gzip.decompress
    if cid == "bz2":
        import bz2

        return bz2.decompress
    if cid == "lzma":
        import lzma

        fmt, filters = compressor.get("format", lzma.FORMAT_XZ), compressor.get("filters")
        return lambda raw: lzma.decompress(raw, format=fmt, filters=filters)
    if cid == "zstd":
        codec = _pyarrow_codec("zstd")

        def zstd(raw):  # a standard frame; its header carries the decoded size
            size = _zstd_content_size(raw)
            if size is None:
                raise NotImplementedError("zstd frame without a content size")
            return codec.decompress(raw, decompressed_size=size).to_pybytes()

        return zstd
    if cid == "lz4":
        codec = _pyarrow_codec("lz4_raw")

        def lz4(raw):  # numcodecs: little-endian int32 decoded size, then one raw LZ4 block
            size = int.from_bytes(raw[:4], "little", signed=True)
            return codec.decompress(raw[4:], decompressed_size=size).to_pybytes()

        return lz4
    if cid == "blosc":
        return blosc_decompress
    raise NotImplementedError(f"Zarr chunks compressed with {cid!r} need the zarr package, which this build does not "
                              "have; open the store with xarray and pass the arrays instead")


_BLOSC_MAX_SPLITS, _BLOSC_MIN_BUFFERSIZE = 16, 128


def _blosc_stream_decoder(fmt: int):
    """``(bytes, decoded size) -> bytes`` of the inner codec a Blosc-1 frame names in bits 5-7 of its flags."""
    if fmt == 1:  # LZ4 / LZ4HC: one raw LZ4 block per stream
        codec = _pyarrow_codec("lz4_raw")
        return lambda raw, n: codec.decompress(raw, decompressed_size=n).to_pybytes()
    if fmt == 2:
        codec = _pyarrow_codec("snappy")
        return lambda raw, n: codec.decompress(raw, decompressed_size=n).to_pybytes()
    if fmt == 3:
        import zlib

        return lambda raw, n: zlib.decompress(raw)
    if fmt == 4:
        codec = _pyarrow_codec("zstd")
        return lambda raw, n: codec.decompress(raw, decompressed_size=n).to_pybytes()
    raise NotImplementedError("Blosc frames compressed with BloscLZ need the zarr / blosc packages, which this build "
                              "does not have (lz4, lz4hc, snappy, zlib and zstd frames are decoded)")


def blosc_decompress(frame: bytes) -> bytes:
    """Decode one Blosc-1 frame (what numcodecs' ``Blosc`` -- zarr-v2's default compressor -- writes per chunk),
    following the container layout c-blosc publishes (README_HEADER.rst, ``blosc_d`` in blosc.c):

    * 16-byte header: version, versionlz, flags, typesize, then little-endian uint32 ``nbytes`` (decoded size),
      ``blocksize``, ``cbytes`` (size of the whole frame).  Flags: bit 0 byte shuffle, bit 1 the payload is a
      plain copy, bit 2 bit shuffle, bit 4 blocks are not split, bits 5-7 the inner codec
      (0 BloscLZ, 1 LZ4 / LZ4HC, 2 Snappy, 3 Zlib, 4 Zstd).
    * ``ceil(nbytes / blocksize)`` int32 block offsets (from the start of the frame), then the blocks.  A block
      is ``typesize`` streams (one per byte position of the shuffled block) when splitting applies -- not the
      last, partial block; ``typesize <= 16``; at least 128 bytes per stream -- else one stream.  A stream is an
      int32 compressed size followed by the data, stored verbatim when that size equals the decoded size.
    * byte shuffle within a block of ``n`` elements: byte ``j`` of element ``i`` sits at ``j * n + i``; the
      ``blocksize % typesize`` trailing bytes are not shuffled.

    The inner codecs come from the standard library / pyarrow; there is no Blosc library in this image, so
    the tests assemble their frames by hand from this description.  Bit-shuffled frames and BloscLZ raise
    ``NotImplementedError``."""
    if len(frame) < 16:
        raise ValueError("not a Blosc frame: shorter than its 16-byte header")
    version, _, flags, typesize = frame[0], frame[1], frame[2], frame[3]
    nbytes, blocksize, cbytes = (int.from_bytes(frame[k:k + 4], "little") for k in (4, 8, 12))
    if version != 2 or cbytes > len(frame) or (nbytes and blocksize == 0):
        raise ValueError(f"not a Blosc-1 frame (version {version}, cbytes {cbytes} of {len(frame)} bytes)")
    if nbytes == 0:
        return b""
    if flags & 0x2:  # stored
        if 16 + nbytes > len(frame):
            raise ValueError("truncated Blosc frame")
        return bytes(frame[16:16 + nbytes])
    if flags & 0x4:
        raise NotImplementedError("bit-shuffled Blosc frames need the zarr / blosc packages, which this build does "
                                  "not have")
    decode = _blosc_stream_decoder(flags >> 5)
    typesize = max(typesize, 1)
    shuffled = bool(flags & 0x1) and typesize > 1
    nblocks = -(-nbytes // blocksize)
    starts = np.frombuffer(frame, dtype="<i4", count=nblocks, offset=16)
    out = np.empty(nbytes, dtype=np.uint8)
    for b in range(nblocks):
        bsize = min(blocksize, nbytes - b * blocksize)
        leftover = bsize != blocksize
        split = (not flags & 0x10 and typesize <= _BLOSC_MAX_SPLITS
                 and blocksize // typesize >= _BLOSC_MIN_BUFFERSIZE and not leftover)
        nstreams = typesize if split else 1
        neblock = bsize // nstreams
        pos, parts = int(starts[b]), []
        for _ in range(nstreams):
            if pos < 16 or pos + 4 > len(frame):
                raise ValueError("corrupt Blosc frame: stream outside the frame")
            csize = int.from_bytes(frame[pos:pos + 4], "little", signed=True)
            pos += 4
            if csize < 0 or pos + csize > len(frame):
                raise ValueError("corrupt Blosc frame: stream outside the frame")
            raw = frame[pos:pos + csize]
            pos += csize
            part = bytes(raw) if csize == neblock else decode(raw, neblock)
            if len(part) != neblock:
                raise ValueError("corrupt Blosc frame: stream decodes to the wrong size")
            parts.append(part)
        block = np.frombuffer(b"".join(parts), dtype=np.uint8)
        dst = out[b * blocksize:b * blocksize + bsize]
        if shuffled:
            n = bsize // typesize
            dst[:n * typesize] = block[:n * typesize].reshape(typesize, n).T.reshape(-1)
            dst[n * typesize:] = block[n * typesize:]
        else:
            dst[:] = block
    return out.tobytes()


def _zstd_content_size(raw: bytes) -> int | None:
    """Frame_Content_Size of a zstd frame header (RFC 8878, 3.1.1.1), ``None`` if the frame omits it."""
    if len(raw) < 6 or raw[:4] != b"\x28\xb5\x2f\xfd":
        raise ValueError("not a zstd frame")
    desc = raw[4]
    fcs_flag, single_segment, dict_flag = desc >> 6, (desc >> 5) & 1, desc & 3
    pos = 5 + (0 if single_segment else 1) + (0, 1, 2, 4)[dict_flag]
    n = (1 if single_segment else 0, 2, 4, 8)[fcs_flag]
    if n == 0:
        return None
    value = int.from_bytes(raw[pos:pos + n], "little")
    return value + 256 if n == 2 else value


class LazySource:
    """A (bands, h, w) or (h, w) array that can deliver whole bands into a caller's buffer."""

    shape: tuple
    dtype: np.dtype

    def read_bands(self, b0: int, nb: int, out: np.ndarray) -> None:  # pragma: no cover - interface
        """Fill ``out[:nb]`` (shape (>= nb, h, w), C-contiguous rows) with bands ``b0 : b0 + nb``."""
        raise NotImplementedError

    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def nbytes(self) -> int:
        return int(np.prod(self.shape)) * np.dtype(self.dtype).itemsize

    def read_all(self) -> np.ndarray:
        bands = 1 if len(self.shape) == 2 else self.shape[0]
        out = np.empty((bands,) + tuple(self.shape[-2:]), dtype=self.dtype)
        self.read_bands(0, bands, out)
        return out[0] if len(self.shape) == 2 else out


class NpySource(LazySource):
    """A ``.npy`` file read through a memory map."""

    def __init__(self, path: str):
        self._mm = np.load(path, mmap_mode="r")
        if self._mm.ndim not in (2, 3):
            raise ValueError(f"{path}: expected a 2-D or 3-D array, got {self._mm.ndim} dimensions")
        self.shape, self.dtype = tuple(self._mm.shape), self._mm.dtype

    def read_bands(self, b0, nb, out):
        src = self._mm[None] if self._mm.ndim == 2 else self._mm
        out[:nb] = src[b0:b0 + nb]


class NetCDF3Source(LazySource):
    """One 2-D or 3-D variable of an open ``scipy.io.netcdf_file`` (``mmap=True``).  The file object is kept
    alive by the source; values are delivered as stored, in native byte order."""

    def __init__(self, nc_file, name: str):
        self._file = nc_file  # owns the memory map the variable's data points into
        self._var = nc_file.variables[name]
        if len(self._var.shape) not in (2, 3):
            raise ValueError(f"{name}: expected a 2-D or 3-D variable, got shape {self._var.shape}")
        self.shape = tuple(int(n) for n in self._var.shape)
        self.dtype = self._var.data.dtype.newbyteorder("=")

    def read_bands(self, b0, nb, out):
        data = self._var.data
        out[:nb] = (data[None] if data.ndim == 2 else data)[b0:b0 + nb]


class ZarrV2Source(LazySource):
    """One Zarr-v2 array of a directory store (2-D or 3-D, C order; uncompressed or with a codec of
    :func:`chunk_decoder`)."""

    def __init__(self, path: str):
        meta = json.load(open(os.path.join(path, ".zarray")))
        if meta.get("zarr_format") != 2:
            raise ValueError(f"{path}: not a Zarr v2 array")
        if meta.get("filters"):
            raise NotImplementedError(f"{path}: filtered Zarr arrays need the zarr package, which this "
                                      "build does not have; open the store with xarray and pass the arrays instead")
        try:
            self._decode = chunk_decoder(meta.get("compressor"))
        except NotImplementedError as e:
            raise NotImplementedError(f"{path}: {e}") from None
        if meta.get("order", "C") != "C":
            raise NotImplementedError(f"{path}: only C-order chunks are supported")
        self.path = path
        self.shape = tuple(int(n) for n in meta["shape"])
        if len(self.shape) not in (2, 3):
            raise ValueError(f"{path}: expected a 2-D or 3-D array, got shape {self.shape}")
        self.chunks = tuple(int(n) for n in meta["chunks"])
        self.dtype = np.dtype(meta["dtype"])
        self.fill = meta.get("fill_value")
        self.sep = meta.get("dimension_separator", ".")
        attrs_path = os.path.join(path, ".zattrs")
        self.attrs = json.load(open(attrs_path)) if os.path.isfile(attrs_path) else {}

    def _chunk(self, idx):
        p = os.path.join(self.path, self.sep.join(str(i) for i in idx))
        if not os.path.isfile(p):  # a missing chunk is all fill value
            return np.full(self.chunks, _fill_value(self.fill, self.dtype), dtype=self.dtype)
        if self._decode is None:
            return np.fromfile(p, dtype=self.dtype).reshape(self.chunks)
        with open(p, "rb") as fh:
            return np.frombuffer(self._decode(fh.read()), dtype=self.dtype).reshape(self.chunks)

    def read_bands(self, b0, nb, out):
        three_d = len(self.shape) == 3
        cb = self.chunks[0] if three_d else 1
        ch, cw = self.chunks[-2:]
        h, w = self.shape[-2:]
        for kb in range(b0 // cb, -(-(b0 + nb) // cb)):
            lo, hi = max(b0, kb * cb), min(b0 + nb, (kb + 1) * cb)
            for kj in range(-(-h // ch)):
                for ki in range(-(-w // cw)):
                    block = self._chunk((kb, kj, ki) if three_d else (kj, ki))
                    if not three_d:
                        block = block[None]
                    j0, i0 = kj * ch, ki * cw
                    j1, i1 = min(h, j0 + ch), min(w, i0 + cw)
                    out[lo - b0:hi - b0, j0:j1, i0:i1] = block[lo - kb * cb:hi - kb * cb, :j1 - j0, :i1 - i0]


def cf_decoded_dtype(dtype: np.dtype, attrs: Mapping) -> np.dtype | None:
    """The float type CF mask-and-scale decoding gives a variable, ``None`` when its attributes ask for none:
    float32 for integers of up to 16 bits and floats of up to 32 bits (float32 holds them exactly), float64
    otherwise -- xarray's ``open_dataset(mask_and_scale=True)`` rule for variables without ``add_offset``
    (with one, xarray's choice has varied between releases; this build keeps the same rule)."""
    if not any(k in attrs for k in ("scale_factor", "add_offset", "_FillValue", "missing_value")):
        return None
    dtype = np.dtype(dtype)
    small = (dtype.kind in "iu" and dtype.itemsize <= 2) or (dtype.kind == "f" and dtype.itemsize <= 4)
    return np.dtype(np.float32 if small else np.float64)


class DecodedSource(LazySource):
    """CF mask-and-scale decoding of another source on the reader thread: ``_FillValue`` / ``missing_value``
    become NaN, then ``raw * scale_factor + add_offset`` in the decoded float type (:func:`cf_decoded_dtype`)."""

    def __init__(self, source: LazySource, attrs: Mapping):
        self.source, self.shape = source, tuple(source.shape)
        self.dtype = cf_decoded_dtype(source.dtype, attrs)
        if self.dtype is None:
            raise ValueError("the variable's attributes ask for no decoding")
        self._fills = []
        for key in ("_FillValue", "missing_value"):
            for f in np.asarray(attrs.get(key, [])).ravel().tolist():
                try:  # JSON attributes may spell the float specials as strings ("NaN"); those mask nothing new
                    f = float(f) if not isinstance(f, (int, np.integer)) else int(f)
                except (TypeError, ValueError):
                    continue
                if f == f and (source.dtype.kind == "f" or float(f).is_integer()):
                    self._fills.append(f)
        self._scale, self._offset = attrs.get("scale_factor"), attrs.get("add_offset")

    def read_bands(self, b0, nb, out):
        raw = np.empty((nb,) + self.shape[-2:], dtype=self.source.dtype)
        self.source.read_bands(b0, nb, raw)
        res = out[:nb]
        np.copyto(res, raw, casting="unsafe")
        for f in self._fills:  # a NaN fill value of float data is NaN already and is not in the list
            if raw.dtype.kind == "f" or np.iinfo(raw.dtype).min <= f <= np.iinfo(raw.dtype).max:
                res[raw == raw.dtype.type(f)] = np.nan
        if self._scale is not None:
            np.multiply(res, self.dtype.type(self._scale), out=res)
        if self._offset is not None:
            np.add(res, self.dtype.type(self._offset), out=res)


_CF_CODING_ATTRS = ("scale_factor", "add_offset", "_FillValue", "missing_value")


def _lazy_variable(source: LazySource, dims, attrs, name, mask_and_scale: bool) -> "LazyDataArray":
    if mask_and_scale and cf_decoded_dtype(source.dtype, attrs) is not None:
        source = DecodedSource(source, attrs)
        attrs = {k: v for k, v in attrs.items() if k not in _CF_CODING_ATTRS}  # moved to the encoding, as in xarray
    return LazyDataArray(source, dims=dims, attrs=attrs, name=name)


class LazyDataArray(DataArray):
    """A dataset variable whose values stay in a chunked store until a pipeline streams them (``source``);
    ``values`` materialises the whole array for anything that is not a streaming consumer."""

    __slots__ = ("source",)

    def __init__(self, source: LazySource, dims, attrs=None, name=None):
        self.source = source
        stand_in = np.broadcast_to(np.zeros((), dtype=source.dtype), source.shape)  # shape / dtype without memory
        DataArray.__init__(self, stand_in, dims=dims, attrs=attrs, name=name)

    @property
    def values(self) -> np.ndarray:
        return self.source.read_all()

    data = values


def _fill_value(fill, dtype: np.dtype):
    """The value of a missing chunk: ``fill_value`` of ``.zarray`` (``"NaN"`` / ``"Infinity"`` /
    ``"-Infinity"`` are the JSON spellings of the float specials; ``null`` leaves it undefined -- zero here)."""
    if isinstance(fill, str) or fill is None:
        return {"NaN": np.nan, "Infinity": np.inf, "-Infinity": -np.inf}.get(fill, 0) if dtype.kind == "f" else 0
    return fill


def _read_small_array(arr_dir: str, meta: dict) -> np.ndarray:
    """A 0-D or 1-D Zarr-v2 array (``spatial_ref``, axis coordinates), chunks decoded like those of
    :class:`ZarrV2Source`; missing chunks hold the fill value."""
    if meta.get("filters"):
        raise NotImplementedError(f"{arr_dir}: filtered Zarr arrays need the zarr package, which this build does not "
                                  "have; open the store with xarray and pass the arrays instead")
    try:
        decode = chunk_decoder(meta.get("compressor"))
    except NotImplementedError as e:
        raise NotImplementedError(f"{arr_dir}: {e}") from None
    dt = np.dtype(meta["dtype"])
    fill = _fill_value(meta.get("fill_value"), dt)

    def chunk(name, count):
        p = os.path.join(arr_dir, name)
        if not os.path.isfile(p):
            return np.full(count, fill, dtype=dt)
        raw = open(p, "rb").read()
        return np.frombuffer(raw if decode is None else decode(raw), dtype=dt)[:count]

    shape = tuple(int(n) for n in meta["shape"])
    if not shape:
        return chunk("0", 1)[0].copy()
    n, c = shape[0], int(meta["chunks"][0])
    return np.concatenate([chunk(str(k), c)[:min(c, n - k * c)] for k in range(-(-n // c))]) if n else np.empty(0, dt)


def open_zarr_dataset(path: str, coord_names=("lon", "lat", "x", "y", "longitude", "latitude",
                                              "transformed_x", "transformed_y"), mask_and_scale: bool = False) -> Dataset:
    """A Zarr-v2 directory store as a :class:`Dataset`: scalar, 1-D and coordinate-named arrays are read
    now, every other 2-D / 3-D array becomes a :class:`LazyDataArray`.  By default values are delivered
    as stored (what ``xarray.open_zarr(..., mask_and_scale=False)`` gives): ``scale_factor`` /
    ``add_offset`` / ``_FillValue`` attributes are passed through, not applied.  ``mask_and_scale=True``
    decodes the lazy variables that carry such attributes on the reader thread (:class:`DecodedSource`)."""
    data_vars, coords = {}, {}
    for item in sorted(os.listdir(path)):
        arr_dir = os.path.join(path, item)
        if not os.path.isfile(os.path.join(arr_dir, ".zarray")):
            continue
        meta = json.load(open(os.path.join(arr_dir, ".zarray")))
        attrs_path = os.path.join(arr_dir, ".zattrs")
        attrs = json.load(open(attrs_path)) if os.path.isfile(attrs_path) else {}
        dims = attrs.pop("_ARRAY_DIMENSIONS", None) or [f"dim_{k}" for k in range(len(meta["shape"]))]
        shape = tuple(meta["shape"])
        if len(shape) <= 1:
            coords[item] = DataArray(_read_small_array(arr_dir, meta), dims=dims if shape else (), attrs=attrs, name=item)
        else:
            src = ZarrV2Source(arr_dir)
            if item in coord_names:
                coords[item] = DataArray(src.read_all(), dims=dims, attrs=attrs, name=item)
            else:
                data_vars[item] = _lazy_variable(src, dims, attrs, item, mask_and_scale)
    group_attrs = os.path.join(path, ".zattrs")
    return Dataset(data_vars=data_vars, coords=coords, attrs=json.load(open(group_attrs)) if os.path.isfile(group_attrs) else {})


def open_netcdf_dataset(path: str, coord_names=("lon", "lat", "x", "y", "longitude", "latitude",
                                                "transformed_x", "transformed_y"), mask_and_scale: bool = False) -> Dataset:
    """A NetCDF classic file as a :class:`Dataset`, with the rules of :func:`open_zarr_dataset`: scalar, 1-D
    and coordinate-named variables are read now, every other 2-D / 3-D variable becomes a
    :class:`LazyDataArray` over the file's memory map.  Values as stored unless ``mask_and_scale=True``
    (then as in :func:`open_zarr_dataset`); ``bytes`` attributes are decoded to ``str`` as xarray does."""
    from scipy.io import netcdf_file

    with open(path, "rb") as fh:
        magic = fh.read(4)
    if magic[:3] != b"CDF":
        kind = "a NetCDF-4 / HDF5 file, which needs the netCDF4 or h5py package (not in this build); open it with " \
               "xarray and pass the arrays instead" if magic == b"\x89HDF" else "not a NetCDF classic file"
        raise NotImplementedError(f"{path}: {kind}")

    class MappedFile(netcdf_file):
        """Lazy variables keep views of the memory map for as long as they live, so the map is released by
        the garbage collector, not by close(): scipy's warning about exactly that is not news here."""

        def close(self):
            import warnings

            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                super().close()

        __del__ = close

    nc = MappedFile(path, "r", mmap=True, maskandscale=False)

    def text(v):
        return v.decode("utf-8", "replace") if isinstance(v, bytes) else v

    data_vars, coords = {}, {}
    for name, var in nc.variables.items():
        attrs = {k: text(v) for k, v in var._attributes.items()}
        dims = tuple(var.dimensions)
        if len(var.shape) in (2, 3) and name not in coord_names:
            data_vars[name] = _lazy_variable(NetCDF3Source(nc, name), dims, attrs, name, mask_and_scale)
            continue
        values = np.array(var.data, dtype=var.data.dtype.newbyteorder("="))  # a copy: independent of the map
        (coords if len(var.shape) <= 2 else data_vars)[name] = DataArray(values, dims=dims, attrs=attrs, name=name)
    return Dataset(data_vars=data_vars, coords=coords, attrs={k: text(v) for k, v in nc._attributes.items()})


def write_zarr_array(path: str, array: np.ndarray, chunks, dims) -> None:
    """Write ``array`` as an uncompressed Zarr-v2 array (test fixtures, synthetic benchmark stores)."""
    os.makedirs(path, exist_ok=True)
    array = np.asarray(array)
    chunks = tuple(int(c) for c in chunks)
    json.dump({"chunks": list(chunks), "compressor": None, "dtype": array.dtype.str, "fill_value": None, "filters": None,
               "order": "C", "shape": list(array.shape), "zarr_format": 2}, open(os.path.join(path, ".zarray"), "w"))
    json.dump({"_ARRAY_DIMENSIONS": list(dims)}, open(os.path.join(path, ".zattrs"), "w"))
    if array.ndim == 0:
        array.reshape(1).tofile(os.path.join(path, "0"))
        return
    grid = [range(-(-n // c)) for n, c in zip(array.shape, chunks)]
    for idx in np.ndindex(*[len(g) for g in grid]):
        block = np.zeros(chunks, dtype=array.dtype)
        sl = tuple(slice(k * c, min(n, (k + 1) * c)) for k, c, n in zip(idx, chunks, array.shape))
        part = array[sl]
        block[tuple(slice(0, s) for s in part.shape)] = part
        block.tofile(os.path.join(path, ".".join(str(k) for k in idx)))
