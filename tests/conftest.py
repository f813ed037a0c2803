import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(autouse=True, scope="session")
def _oracle_default_scalar_rules():
    """The oracle's functions default to scalar_rules="numpy1" (the NumPy the reference documents); the product
    defaults to the regime of the installed NumPy.  Tests compare the two under one stated regime: numpy1 unless a
    test selects the other itself with monkeypatch (the reference-run fixtures, the numpy2 tests).  Session scope:
    module-scoped fixtures build their plans before any function-scoped fixture runs."""
    from xcontour_b200 import utils
    old = utils.NUMPY_SCALAR_RULES
    utils.NUMPY_SCALAR_RULES = "numpy1"
    yield
    utils.NUMPY_SCALAR_RULES = old


@pytest.fixture(scope="session")
def vort():
    """lat[256], lon[512], absolute_vorticity[256,512] of the reference's
    Data/barotropic_vorticity.nc (committed fixture, see make_golden.py)."""
    d = np.load(os.path.join(GOLDEN, "barotropic_vorticity.npz"))
    return d["latitude"], d["longitude"], d["absolute_vorticity"]


def synth_c4(S, ny=721, nx=1440, seed=1234, dtype=np.float32):
    """Synthetic ERA5-like tracer of SURVEY.md §8(d): sin(phi) + wave + noise."""
    lat = np.linspace(-90.0, 90.0, ny)
    lon = np.arange(nx) * (360.0 / nx)
    phi, lam = np.deg2rad(lat)[:, None], np.deg2rad(lon)[None, :]
    out = np.empty((S, ny, nx), dtype=dtype)
    for s in range(S):
        rng = np.random.default_rng(seed + s)
        phase = 2 * np.pi * rng.random()
        out[s] = (np.sin(phi) + 0.3 * np.cos(phi) ** 2 * np.sin(6 * lam + 3 * phi + phase)
                  + 0.02 * rng.standard_normal((ny, nx))).astype(dtype)
    return lat.astype(np.float32), lon.astype(np.float32), out
