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


C_CLIENT = r"""
#include <stdio.h>
#include <string.h>
#include "xrs.h"

/* A plain C99 client of the boundary: argument checks run on the host, so these calls need no GPU. */
int main(void) {
    if (xrs_version() != 100) return 1;
    if (xrs_quad_row_group() < 1) return 2;
    if (xrs_rectify_ij_workspace_bytes(100, 200, 50, 60) <= 0) return 3;
    if (xrs_rectify_ij_workspace_bytes(1, 1, 0, 0) != 0) return 4;
    /* null plane tables are rejected with an error message instead of a crash */
    if (xrs_gather_ij(NULL, NULL, 1, XRS_F32, 4, 4, 4, 0, 0, 4, 4, NULL, 4, 4, XRS_NEAREST, 0.0, NULL) == 0) return 5;
    if (strstr(xrs_last_error(), "null") == NULL) return 6;
    if (xrs_gather_ij2(NULL, NULL, NULL, 1, XRS_F32, 4, 4, 4, 0, 0, 4, 4, NULL, 4, 4, XRS_BILINEAR, 0.0, 0.0, NULL) == 0)
        return 7;
    printf("ok %d\n", xrs_version());
    return 0;
}
"""


def test_header_is_c99_and_a_c_client_links(lib_path, tmp_path):
    """include/xrs.h is a C header (no C++ in the signatures): a C99 translation unit compiles with
    -pedantic, links against libxrs.so and gets host-side argument errors through xrs_last_error()."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "client.c"
    src.write_text(C_CLIENT)
    exe = tmp_path / "client"
    lib_dir = os.path.dirname(lib_path)
    cmd = [gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", f"-I{os.path.join(ROOT, 'include')}", str(src), "-o", str(exe),
           f"-L{lib_dir}", f"-l:{os.path.basename(lib_path)}", f"-Wl,-rpath,{lib_dir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0 and run.stdout.strip() == "ok 100", (run.returncode, run.stdout, run.stderr)


def test_integration_md_stub_binds_and_matches_the_binding_table(lib_path, monkeypatch):
    """The ctypes stub INTEGRATION.md shows a reference maintainer is real code: it is executed here
    against libxrs.so, and every function it binds has exactly the argument and result types of the
    package's own binding table (``_lib.SIGNATURES``), i.e. of include/xrs.h."""
    from xcube_resampling_b200 import _lib

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "_bind(" in b)
    monkeypatch.setenv("XRS_LIBRARY", lib_path)
    ns: dict = {}
    exec(compile(stub, "INTEGRATION.md", "exec"), ns)
    bound = {v.__name__: v for v in ns.values() if isinstance(v, ctypes._CFuncPtr)}
    assert {"xrs_rectify_ij", "xrs_gather_ij", "xrs_gather_ij2", "xrs_reproject", "xrs_affine", "xrs_coarsen",
            "xrs_tile_src_bboxes", "xrs_transform_points"} <= set(bound)
    for name, fn in bound.items():
        restype, argtypes = _lib.SIGNATURES[name]
        assert list(fn.argtypes) == list(argtypes), f"{name}: INTEGRATION.md stub and the binding table disagree"
        assert fn.restype == restype, name
    with pytest.raises(RuntimeError, match="null"):
        ns["check"](bound["xrs_gather_ij"](None, None, 1, 0, 4, 4, 4, 0, 0, 4, 4, None, 4, 4, 0, 0.0, None))
