"""Minimal coordinate-reference-system descriptor.

The reference uses ``pyproj.CRS`` everywhere a CRS is needed
(``gridmapping/helpers.py:59-63``, ``utils.py:187-189``).  pyproj / PROJ are not
part of this build, so the hot path carries its own small descriptor: the
projection family and the parameters the fp64 device formulas need
(``include/xrs.h: struct xrs_proj``; formulas in SURVEY.md 7.5).  ``pyproj.CRS``
objects are accepted wherever a CRS is expected when pyproj is importable.
"""

from __future__ import annotations

import dataclasses
import re
from typing import Any

WGS84_A = 6378137.0
WGS84_INV_F = 298.257223563
GRS80_INV_F = 298.257222101

KIND_GEOGRAPHIC, KIND_TMERC, KIND_WEBMERC, KIND_LAEA = 0, 1, 2, 3
# A CRS this build holds no device formulas for (Lambert conformal conic, polar stereographic, national
# grids, rotated poles ...).  It can name the grid of a dataset and be compared -- everything the same-CRS
# paths need (affine_transform_dataset, rectify within one CRS, utils.py:186-189) -- but asking for its
# projection parameters, i.e. for a transform from or to it, is an error.
KIND_OPAQUE = -1
_KIND_NAMES = {0: "latitude_longitude", 1: "transverse_mercator", 2: "mercator", 3: "lambert_azimuthal_equal_area"}
# EPSG codes of geographic CRSs that are met in Earth-observation products; any other code without device
# formulas is taken as projected (there is no EPSG database in this build)
_GEOGRAPHIC_EPSG = frozenset({4019, 4030, 4035, 4047, 4148, 4167, 4171, 4173, 4230, 4267, 4269, 4272, 4283, 4289, 4312,
                              4313, 4314, 4322, 4490, 4612, 4617, 4619, 4674, 4937, 4979, 6318, 6668, 7844, 9057})
_GEOGRAPHIC_CF_NAMES = ("latitude_longitude", "rotated_latitude_longitude")


@dataclasses.dataclass(frozen=True)
class CRS:
    """Projection family + parameters; compare with :meth:`equals` / ``==``."""

    kind: int
    name: str
    a: float = WGS84_A
    inv_f: float = WGS84_INV_F
    lon0: float = 0.0
    lat0: float = 0.0
    k0: float = 1.0
    fe: float = 0.0
    fn: float = 0.0
    epsg: int | None = None
    # EPSG:4326 has (lat, lon) axis order, OGC:CRS84 (lon, lat).  Transforms always
    # run always_xy=True (reproject.py:124-126) so this only affects equality.
    lat_first: bool = False
    # KIND_OPAQUE only: is it geographic, and the CF attributes it was read from (sorted items)
    opaque_geographic: bool = False
    cf: tuple = ()

    # -- construction ------------------------------------------------------
    @staticmethod
    def from_epsg(code: int) -> "CRS":
        code = int(code)
        if code == 4326:
            return CRS(KIND_GEOGRAPHIC, "WGS 84", epsg=4326, lat_first=True)
        if code == 4258:
            return CRS(KIND_GEOGRAPHIC, "ETRS89", inv_f=GRS80_INV_F, epsg=4258, lat_first=True)
        if 32601 <= code <= 32660 or 32701 <= code <= 32760:
            zone = code % 100
            south = code >= 32700
            return CRS(KIND_TMERC, f"WGS 84 / UTM zone {zone}{'S' if south else 'N'}", lon0=6.0 * zone - 183.0,
                       k0=0.9996, fe=500000.0, fn=10000000.0 if south else 0.0, epsg=code)
        if 25828 <= code <= 25838:
            zone = code % 100
            return CRS(KIND_TMERC, f"ETRS89 / UTM zone {zone}N", inv_f=GRS80_INV_F, lon0=6.0 * zone - 183.0,
                       k0=0.9996, fe=500000.0, epsg=code)
        if code == 3857:
            return CRS(KIND_WEBMERC, "WGS 84 / Pseudo-Mercator", epsg=3857)
        if code == 3035:
            return CRS(KIND_LAEA, "ETRS89-extended / LAEA Europe", inv_f=GRS80_INV_F, lon0=10.0, lat0=52.0,
                       fe=4321000.0, fn=3210000.0, epsg=3035)
        if not 1024 <= code <= 32767:
            raise ValueError(f"EPSG:{code} is not a CRS code")
        return CRS(KIND_OPAQUE, f"EPSG:{code}", epsg=code, opaque_geographic=code in _GEOGRAPHIC_EPSG)

    @staticmethod
    def from_string(text: str) -> "CRS":
        t = text.strip()
        if t.upper() in ("OGC:CRS84", "CRS84", "CRS:84", "URN:OGC:DEF:CRS:OGC:1.3:CRS84"):
            return CRS(KIND_GEOGRAPHIC, "WGS 84 (CRS84)", epsg=None, lat_first=False)
        m = re.fullmatch(r"(?i)(?:epsg:|urn:ogc:def:crs:epsg::)(\d+)", t)
        if m:
            return CRS.from_epsg(int(m.group(1)))
        raise ValueError(f"cannot interpret CRS string {text!r}")

    @staticmethod
    def from_cf(attrs: dict) -> "CRS":
        """Inverse of :meth:`to_cf` for the projections this build knows."""
        if "epsg_code" in attrs:
            return CRS.from_string(str(attrs["epsg_code"]))
        name = attrs.get("grid_mapping_name")
        if name not in _KIND_NAMES.values() or (name == "mercator" and "crs_wkt" in attrs):
            code = _epsg_of_wkt(attrs.get("crs_wkt") or attrs.get("spatial_ref"))
            if code is not None:
                return CRS.from_epsg(code)
            if isinstance(name, str) and name and name != "mercator":
                items = tuple(sorted((str(k), v if isinstance(v, (str, int, float, bool)) else repr(v))
                                     for k, v in attrs.items()))
                return CRS(KIND_OPAQUE, str(attrs.get("crs_name") or attrs.get("projected_crs_name") or name),
                           opaque_geographic=name in _GEOGRAPHIC_CF_NAMES, cf=items)
        a = float(attrs.get("semi_major_axis", WGS84_A))
        inv_f = float(attrs.get("inverse_flattening", WGS84_INV_F))
        if name == "latitude_longitude":
            return CRS(KIND_GEOGRAPHIC, "unknown geographic", a=a, inv_f=inv_f, lat_first=True)
        if name == "transverse_mercator":
            return CRS(KIND_TMERC, "unknown tmerc", a=a, inv_f=inv_f,
                       lon0=float(attrs.get("longitude_of_central_meridian", 0.0)),
                       lat0=float(attrs.get("latitude_of_projection_origin", 0.0)),
                       k0=float(attrs.get("scale_factor_at_central_meridian", 1.0)),
                       fe=float(attrs.get("false_easting", 0.0)), fn=float(attrs.get("false_northing", 0.0)))
        if name == "lambert_azimuthal_equal_area":
            return CRS(KIND_LAEA, "unknown laea", a=a, inv_f=inv_f,
                       lon0=float(attrs.get("longitude_of_projection_origin", 0.0)),
                       lat0=float(attrs.get("latitude_of_projection_origin", 0.0)),
                       fe=float(attrs.get("false_easting", 0.0)), fn=float(attrs.get("false_northing", 0.0)))
        raise ValueError(f"cannot interpret CF grid mapping attributes {attrs!r}")

    # -- pyproj-like surface used by the hot path --------------------------
    @property
    def is_geographic(self) -> bool:
        return self.kind == KIND_GEOGRAPHIC or (self.kind == KIND_OPAQUE and self.opaque_geographic)

    @property
    def is_projected(self) -> bool:
        return not self.is_geographic

    @property
    def has_device_formulas(self) -> bool:
        """False for a CRS that can only take part in same-CRS operations."""
        return self.kind != KIND_OPAQUE

    @property
    def unit_name(self) -> str:
        """Unit of the first axis, as ``pyproj.CRS.axis_info[0].unit_name`` names it (gridmapping/base.py:402-404)."""
        return "degree" if self.is_geographic else "metre"

    def __str__(self) -> str:
        """Authority string where there is one (what ``str(pyproj.CRS)`` prints for these), else the name."""
        if self.epsg is not None:
            return f"EPSG:{self.epsg}"
        if self.kind == KIND_GEOGRAPHIC and not self.lat_first and self.a == WGS84_A and self.inv_f == WGS84_INV_F:
            return "OGC:CRS84"
        return self.name

    def _key(self):
        if self.kind == KIND_OPAQUE:
            return (KIND_OPAQUE, self.epsg, self.cf if self.epsg is None else ())
        return (self.kind, self.a, self.inv_f, self.lon0, self.lat0, self.k0, self.fe, self.fn, self.lat_first)

    def equals(self, other: Any) -> bool:
        other = normalize_crs(other)
        return self._key() == other._key()

    def __eq__(self, other):
        try:
            return self.equals(other)
        except (ValueError, TypeError):
            return NotImplemented

    def __hash__(self):
        return hash(self._key())

    def to_cf(self) -> dict:
        """CF grid-mapping attributes (subset of what ``pyproj.CRS.to_cf`` emits)."""
        if self.kind == KIND_OPAQUE:
            return dict(self.cf) if self.epsg is None else {"crs_name": self.name, "epsg_code": f"EPSG:{self.epsg}"}
        cf = {
            "grid_mapping_name": _KIND_NAMES[self.kind],
            "semi_major_axis": self.a,
            "inverse_flattening": self.inv_f,
            "crs_name": self.name,
        }
        if self.epsg is not None:
            cf["epsg_code"] = f"EPSG:{self.epsg}"
        elif self.kind == KIND_GEOGRAPHIC and not self.lat_first:
            cf["epsg_code"] = "OGC:CRS84"
        if self.kind == KIND_TMERC:
            cf.update(longitude_of_central_meridian=self.lon0, latitude_of_projection_origin=self.lat0,
                      scale_factor_at_central_meridian=self.k0, false_easting=self.fe, false_northing=self.fn)
        elif self.kind == KIND_LAEA:
            cf.update(longitude_of_projection_origin=self.lon0, latitude_of_projection_origin=self.lat0,
                      false_easting=self.fe, false_northing=self.fn)
        return cf

    def proj_params(self) -> tuple:
        """(kind, a, inv_f, lon0, lat0, k0, fe, fn) for ``struct xrs_proj``."""
        if self.kind == KIND_OPAQUE:
            raise ValueError(f"{self} has no projection formulas in the B200 resampling path: it can be resampled "
                             "within itself, but not transformed (supported: EPSG 4326, 4258, 326xx/327xx, 258xx, "
                             "3857, 3035 and CF transverse_mercator / lambert_azimuthal_equal_area parameters)")
        return (self.kind, self.a, self.inv_f, self.lon0, self.lat0, self.k0, self.fe, self.fn)


CRS_WGS84 = CRS.from_epsg(4326)
CRS_CRS84 = CRS.from_string("OGC:CRS84")


def _epsg_of_wkt(wkt) -> int | None:
    """The EPSG code a WKT string closes with (WKT2 ``ID["EPSG",n]]`` / WKT1 ``AUTHORITY["EPSG","n"]]``): the
    identifier of the CRS itself, not of one of its components."""
    if not isinstance(wkt, str):
        return None
    m = re.search(r'(?:ID|AUTHORITY)\[\s*"EPSG"\s*,\s*"?(\d+)"?\s*\]\s*\]\s*$', wkt.strip())
    return int(m.group(1)) if m else None


def normalize_crs(crs: Any) -> CRS:
    """gridmapping/helpers.py:59-63 (_normalize_crs) for this build's CRS type."""
    if isinstance(crs, CRS):
        return crs
    if isinstance(crs, str):
        return CRS.from_string(crs)
    if isinstance(crs, int):
        return CRS.from_epsg(crs)
    # duck-typed pyproj.CRS
    to_epsg = getattr(crs, "to_epsg", None)
    if callable(to_epsg):
        code = to_epsg()
        to_cf = getattr(crs, "to_cf", None)
        out = CRS.from_epsg(code) if code is not None else CRS.from_cf(to_cf()) if callable(to_cf) else None
        if out is not None:
            if out.kind == KIND_OPAQUE and isinstance(getattr(crs, "is_geographic", None), bool):
                out = dataclasses.replace(out, opaque_geographic=crs.is_geographic)  # the object knows better
            return out
    raise TypeError(f"crs must be a CRS, an EPSG code or a CRS string, was {type(crs)}")
