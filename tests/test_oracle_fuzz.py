"""Short fuzz campaign of the oracle against the reference's own kernels.  Needs /root/reference
(build container); skipped elsewhere.  Runs in a subprocess because importing the reference
installs stub modules for dask / xarray / pyproj."""

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference sources not present")
def test_oracle_equals_reference_kernels_on_random_swaths():
    pytest.importorskip("numba")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "fuzz_oracle_vs_reference.py"), "11", "120"],
                       capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    for line in ("120 cases, 0 mismatches", "reproject: 40 cases, 0 mismatches", "coarsen: 40 cases, 0 mismatches"):
        assert line in r.stdout, r.stdout[-2000:]
