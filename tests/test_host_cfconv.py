"""CF grid-mapping discovery (xcube_resampling_b200/cfconv.py) against the expectations of the
reference's tests/gridmapping/test_cfconv.py: which CRS, under which key, with which coordinate
variables -- and ``add_spatial_ref`` on an uncompressed Zarr-v2 directory store."""

import json
import os

import numpy as np
import pytest

from xcube_resampling_b200 import CRS, CRS_CRS84, CRS_WGS84, DataArray, Dataset, GridMapping
from xcube_resampling_b200.cfconv import (
    GridCoords,
    GridMappingProxy,
    add_spatial_ref,
    find_potential_coord_vars,
    get_dataset_grid_mapping_proxies,
)

UTM_33N = CRS.from_epsg(32633)


def _check(gmp, crs, name, x_name, y_name):
    assert isinstance(gmp, GridMappingProxy) and isinstance(gmp.coords, GridCoords)
    assert gmp.crs == crs and gmp.name == name
    assert gmp.coords.x.name == x_name and gmp.coords.y.name == y_name


def test_no_crs_lon_lat_common_names():
    # test_cfconv.py:58-75
    ds = Dataset(coords=dict(lon=DataArray(np.linspace(10, 12, 11), dims="lon"),
                             lat=DataArray(np.linspace(50, 52, 11), dims="lat")))
    gms = get_dataset_grid_mapping_proxies(ds)
    assert list(gms) == [None]
    _check(gms[None], CRS_WGS84, "latitude_longitude", "lon", "lat")


def test_no_crs_lon_lat_standard_names():
    # test_cfconv.py:77-103
    ds = Dataset(coords=dict(
        weird_x=DataArray(np.linspace(10, 12, 11), dims="i", attrs=dict(standard_name="longitude")),
        weird_y=DataArray(np.linspace(50, 52, 11), dims="j", attrs=dict(standard_name="latitude"))))
    gms = get_dataset_grid_mapping_proxies(ds)
    assert list(gms) == [None]
    _check(gms[None], CRS_WGS84, "latitude_longitude", "weird_x", "weird_y")


def test_crs_x_y_with_common_and_standard_names():
    # test_cfconv.py:105-153
    ds = Dataset(dict(crs=DataArray(np.array(0), dims=(), attrs=UTM_33N.to_cf())),
                 coords=dict(x=DataArray(np.linspace(1000, 12000, 11), dims="x"),
                             y=DataArray(np.linspace(5000, 52000, 11), dims="y")))
    gms = get_dataset_grid_mapping_proxies(ds)
    assert list(gms) == ["crs"]
    _check(gms["crs"], UTM_33N, "transverse_mercator", "x", "y")
    ds = Dataset(dict(crs=DataArray(np.array(0), dims=(), attrs=UTM_33N.to_cf())),
                 coords=dict(myx=DataArray(np.linspace(1000, 12000, 11), dims="x",
                                           attrs=dict(standard_name="projection_x_coordinate")),
                             myy=DataArray(np.linspace(5000, 52000, 11), dims="y",
                                           attrs=dict(standard_name="projection_y_coordinate"))))
    gms = get_dataset_grid_mapping_proxies(ds)
    _check(gms["crs"], UTM_33N, "transverse_mercator", "myx", "myy")


def test_latitude_longitude_with_x_y():
    # test_cfconv.py:155-179: a CRS-84 GeoTIFF opened with rioxarray
    ds = Dataset(dict(band_1=DataArray(np.zeros((11, 11)), dims=["y", "x"]),
                      spatial_ref=DataArray(np.array(0), dims=(), attrs=CRS_CRS84.to_cf())),
                 coords=dict(x=DataArray(np.linspace(10, 20, 11), dims="x"),
                             y=DataArray(np.linspace(50, 40, 11), dims="y")))
    gms = get_dataset_grid_mapping_proxies(ds)
    assert list(gms) == ["spatial_ref"]
    gmp = gms["spatial_ref"]
    assert gmp.crs.is_geographic and gmp.name == "latitude_longitude"
    assert gmp.coords.x.name == "x" and gmp.coords.y.name == "y"
    gm = GridMapping.from_dataset(ds)
    assert gm.crs.is_geographic and gm.size == (11, 11) and gm.is_regular and not gm.is_j_axis_up


def test_grid_mapping_attribute_and_dataset_attrs():
    attrs = UTM_33N.to_cf()
    ds = Dataset(dict(band=DataArray(np.zeros((5, 6)), dims=["y", "x"], attrs=dict(grid_mapping="my_crs")),
                      my_crs=DataArray(np.array(0), dims=(), attrs=attrs)),
                 coords=dict(x=DataArray(np.linspace(0, 50, 6), dims="x"), y=DataArray(np.linspace(40, 0, 5), dims="y")))
    assert list(get_dataset_grid_mapping_proxies(ds)) == ["my_crs"]
    ds = Dataset(coords=dict(x=DataArray(np.linspace(0, 50, 6), dims="x"), y=DataArray(np.linspace(40, 0, 5), dims="y")),
                 attrs=attrs)
    gms = get_dataset_grid_mapping_proxies(ds)
    assert list(gms) == [None] and gms[None].crs == UTM_33N
    # projected coordinates without any CRS: only usable with the caller's default
    ds = Dataset(coords=dict(x=DataArray(np.linspace(0, 50, 6), dims="x"), y=DataArray(np.linspace(40, 0, 5), dims="y")))
    assert get_dataset_grid_mapping_proxies(ds) == {}
    assert get_dataset_grid_mapping_proxies(ds, missing_projected_crs=UTM_33N)[None].crs == UTM_33N
    with pytest.raises(ValueError, match="cannot find any grid mapping in dataset"):
        GridMapping.from_dataset(ds)
    assert GridMapping.from_dataset(ds, crs=UTM_33N).crs == UTM_33N


def test_2d_lon_lat_and_bounds_are_not_coordinates():
    # test_cfconv.py (_find_potential_coord_vars): bounds variables are excluded, 2-D coordinates found
    lon = np.add.outer(np.zeros(4), np.linspace(10, 12, 5))
    lat = np.add.outer(np.linspace(52, 50, 4), np.zeros(5))
    ds = Dataset(dict(rad=DataArray(np.zeros((4, 5)), dims=["y", "x"]),
                      lon_bnds=DataArray(np.zeros((4, 5)), dims=["y", "x"]),
                      lat_bounds=DataArray(np.zeros((4, 5)), dims=["y", "x"]),
                      cube=DataArray(np.zeros((2, 4, 5)), dims=["t", "y", "x"])),
                 coords=dict(lon=DataArray(lon, dims=["y", "x"], attrs=dict(bounds="lon_bnds")),
                             lat=DataArray(lat, dims=["y", "x"])))
    names = find_potential_coord_vars(ds)
    assert "lon" in names and "lat" in names and "rad" in names
    assert "lon_bnds" not in names and "lat_bounds" not in names and "cube" not in names
    gms = get_dataset_grid_mapping_proxies(ds)
    _check(gms[None], CRS_WGS84, "latitude_longitude", "lon", "lat")
    gm = GridMapping.from_dataset(ds)
    assert gm.size == (5, 4) and gm.xy_dim_names == ("x", "y")


def test_incomplete_coordinates_warn():
    ds = Dataset(dict(crs=DataArray(np.array(0), dims=(), attrs=UTM_33N.to_cf())),
                 coords=dict(x=DataArray(np.linspace(1000, 12000, 11), dims="x")))
    with pytest.warns(UserWarning, match="missing x- and/or y-coordinates"):
        assert get_dataset_grid_mapping_proxies(ds, emit_warnings=True) == {}


def test_add_spatial_ref_to_zarr_directory(tmp_path):
    # test_cfconv.py (AddSpatialRefTest): a (y, x) array gains grid_mapping, a scalar CRS array appears
    store = tmp_path / "cube.zarr"
    os.makedirs(store / "band")
    json.dump({"zarr_format": 2}, open(store / ".zgroup", "w"))
    json.dump({"chunks": [2, 3], "compressor": None, "dtype": "<f4", "fill_value": None, "filters": None, "order": "C",
               "shape": [4, 6], "zarr_format": 2}, open(store / "band" / ".zarray", "w"))
    json.dump({"_ARRAY_DIMENSIONS": ["y", "x"]}, open(store / "band" / ".zattrs", "w"))
    os.makedirs(store / "t")
    json.dump({"chunks": [3], "compressor": None, "dtype": "<i8", "fill_value": None, "filters": None, "order": "C",
               "shape": [3], "zarr_format": 2}, open(store / "t" / ".zarray", "w"))
    json.dump({"_ARRAY_DIMENSIONS": ["t"]}, open(store / "t" / ".zattrs", "w"))
    json.dump({"metadata": {}, "zarr_consolidated_format": 1}, open(store / ".zmetadata", "w"))
    add_spatial_ref(str(store), UTM_33N)
    assert json.load(open(store / "band" / ".zattrs"))["grid_mapping"] == "spatial_ref"
    assert "grid_mapping" not in json.load(open(store / "t" / ".zattrs"))
    sr = json.load(open(store / "spatial_ref" / ".zattrs"))
    assert sr["grid_mapping_name"] == "transverse_mercator" and sr["_ARRAY_DIMENSIONS"] == []
    assert CRS.from_cf(sr) == UTM_33N
    meta = json.load(open(store / ".zmetadata"))["metadata"]
    assert "spatial_ref/.zarray" in meta and meta["band/.zattrs"]["grid_mapping"] == "spatial_ref"
    with pytest.raises(TypeError):
        add_spatial_ref(42, UTM_33N)


def test_crs_in_dataset_attrs_with_wkt():
    # test_cfconv.py:181-224: the CF attributes of EPSG:4326 (crs_wkt included) as DATASET attributes
    wkt = ('GEOGCRS["WGS 84",ENSEMBLE["World Geodetic System 1984 ensemble",MEMBER["World Geodetic System 1984 '
           '(Transit)"],ELLIPSOID["WGS 84",6378137,298.257223563,LENGTHUNIT["metre",1]],ENSEMBLEACCURACY[2.0]],'
           'PRIMEM["Greenwich",0,ANGLEUNIT["degree",0.0174532925199433]],CS[ellipsoidal,2],AXIS["geodetic latitude '
           '(Lat)",north,ORDER[1],ANGLEUNIT["degree",0.0174532925199433]],AXIS["geodetic longitude (Lon)",east,'
           'ORDER[2],ANGLEUNIT["degree",0.0174532925199433]],ID["EPSG",4326]]')
    ds = Dataset(coords=dict(lon=DataArray(np.linspace(10, 12, 11), dims="lon"),
                             lat=DataArray(np.linspace(50, 52, 11), dims="lat")),
                 attrs={"crs_wkt": wkt, "semi_major_axis": 6378137.0, "semi_minor_axis": 6356752.314245179,
                        "inverse_flattening": 298.257223563, "reference_ellipsoid_name": "WGS 84",
                        "longitude_of_prime_meridian": 0.0, "prime_meridian_name": "Greenwich",
                        "geographic_crs_name": "WGS 84",
                        "horizontal_datum_name": "World Geodetic System 1984 ensemble",
                        "grid_mapping_name": "latitude_longitude"})
    gms = get_dataset_grid_mapping_proxies(ds)
    assert list(gms) == [None]
    _check(gms[None], CRS_WGS84, "latitude_longitude", "lon", "lat")


def test_single_point_coordinates_warn_once():
    # test_cfconv.py:226-237: one-element coordinates are no grid
    import warnings

    ds = Dataset(coords=dict(lon=DataArray(np.array([10]), dims="lon"), lat=DataArray(np.array([50]), dims="lat")))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        get_dataset_grid_mapping_proxies(ds, emit_warnings=True)
    assert len(w) == 1 and "missing x- and/or y-coordinates" in str(w[0].message)


def test_bounds_detection_and_coordinates_attribute():
    # test_cfconv.py:287-320
    ds = Dataset(coords={"lon": DataArray(np.linspace(0, 10, 5), dims="lon"),
                         "lat": DataArray(np.linspace(0, 5, 5), dims="lat"),
                         "lon_bnds": DataArray(np.linspace(0, 10, 10), dims="bnds"),
                         "lat_bounds": DataArray(np.linspace(0, 5, 10), dims="bnds"),
                         "alt": DataArray(np.linspace(0, 100, 5), dims="alt")})
    ds["lat"].attrs["bounds"] = "lat_bounds"
    names = find_potential_coord_vars(ds)
    assert {"lon", "lat", "alt"} <= set(names) and "lon_bnds" not in names and "lat_bounds" not in names
    ds = Dataset({"x": DataArray(np.array([0, 1]), dims="dim_0"), "y": DataArray(np.array([0, 1]), dims="dim_0")},
                 attrs={"coordinates": "x y"})
    names = find_potential_coord_vars(ds)
    assert "x" in names and "y" in names


def test_add_spatial_ref_custom_variable_name(tmp_path):
    # test_cfconv.py:430-470 (TestAddSpatialRef) on a directory store
    store = tmp_path / "c.zarr"
    os.makedirs(store / "data")
    json.dump({"zarr_format": 2}, open(store / ".zgroup", "w"))
    json.dump({"chunks": [3, 3], "compressor": None, "dtype": "<f4", "fill_value": 0.0, "filters": None, "order": "C",
               "shape": [3, 3], "zarr_format": 2}, open(store / "data" / ".zarray", "w"))
    json.dump({"_ARRAY_DIMENSIONS": ["y", "x"]}, open(store / "data" / ".zattrs", "w"))
    add_spatial_ref(str(store), CRS_WGS84, crs_var_name="spatial_ref_test", xy_dim_names=("x", "y"))
    sr = json.load(open(store / "spatial_ref_test" / ".zattrs"))
    assert json.load(open(store / "spatial_ref_test" / ".zarray"))["shape"] == []
    assert sr["_ARRAY_DIMENSIONS"] == [] and len(sr) > 1
    assert json.load(open(store / "data" / ".zattrs"))["grid_mapping"] == "spatial_ref_test"


@pytest.mark.parametrize("coords, x_name, y_name", [
    (dict(rlon=("rlon", {}), rlat=("rlat", {})), "rlon", "rlat"),
    (dict(u=("u", dict(standard_name="grid_longitude")), v=("v", dict(standard_name="grid_latitude"))), "u", "v")])
def test_rotated_pole_is_discovered(coords, x_name, y_name):
    # test_cfconv.py:239-285: a grid mapping this build cannot transform is still found with its coordinates
    pole = dict(grid_mapping_name="rotated_latitude_longitude", grid_north_pole_latitude=32.5,
                grid_north_pole_longitude=170.0)
    values = {x_name: np.linspace(-180, 180, 11), y_name: np.linspace(0, 90, 11)}
    ds = Dataset(dict(rotated_pole=DataArray(np.array(0), dims=(), attrs=pole)),
                 coords={n: DataArray(values[n], dims=d, attrs=a) for n, (d, a) in coords.items()})
    gms = get_dataset_grid_mapping_proxies(ds)
    assert list(gms) == ["rotated_pole"]
    gmp = gms["rotated_pole"]
    assert gmp.crs.is_geographic and not gmp.crs.has_device_formulas and gmp.name == "rotated_latitude_longitude"
    assert gmp.coords.x.name == x_name and gmp.coords.y.name == y_name
