"""
CPU suite (-m "not gpu"): the oracle against the reference's golden vector and
its own cross-checks, the host-side logic, and the C-ABI surface (library loads,
every declared symbol is exported -- no compute calls without a GPU).
"""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import xcontour_oracle as O
from conftest import GOLDEN, ROOT, synth_c4


# ------------------------------------------------------------------ golden vector
def test_contours_golden_notebook():
    """notebooks/1.Keff_atmos.ipynb:102-119 -- 36 printed fp32 levels, bit-exact."""
    g = json.load(open(os.path.join(GOLDEN, "contours_pv.json")))
    N = g["levels_N"]
    for row in g["printed"]:
        vals = np.array([np.float32(x) for x in row])
        # a 1-slice tracer whose min/max are the printed end levels
        q = np.array([[[vals[0], vals[-1]]]], dtype=np.float32)
        ctr = O.cal_contours(q, N, increase=g["increase"], dtype=np.float32)[0]
        got = ctr[g["columns"]]
        assert got.dtype == np.float32
        assert np.array_equal(got, vals), (got, vals)
        # and the printed decimal strings themselves
        for a, b in zip(got, row):
            assert np.format_float_scientific(a, precision=8, unique=False) == \
                np.format_float_scientific(np.float32(b), precision=8, unique=False)


def test_contours_promotion_rule_is_discriminated():
    """An all-fp32 evaluation of the steps does NOT reproduce the golden vector: the
    levels are computed in fp64 (np.vectorize hands `levels` over as np.int64, so
    `1.0/divisor` is a float64) and rounded to fp32 once, which is what the oracle
    restates and what the reference's own code returns when run here
    (tests/test_reference_golden.py)."""
    g = json.load(open(os.path.join(GOLDEN, "contours_pv.json")))
    N, mism = g["levels_N"], 0
    for row in g["printed"]:
        vals = np.array([np.float32(x) for x in row])
        start, stop = vals[0], vals[-1]
        steps = np.float32(1.0 / (N - 1)) * (stop - start)              # all-fp32 steps
        ctr = (np.float64(steps) * np.arange(N) + np.float64(start)).astype(np.float32)
        mism += int(np.sum(ctr[g["columns"]] != vals))
    assert mism > 0


def test_contour_coord_and_decreasing():
    q = np.random.default_rng(0).standard_normal((3, 8, 9)).astype(np.float32)
    c = O.cal_contours(q, 11, increase=False)
    assert np.all(np.diff(c, axis=1) < 0)
    assert np.array_equal(c[:, 0], q.reshape(3, -1).max(1))
    assert np.array_equal(O.contour_coord(5), np.arange(5, dtype=np.float32))


# ------------------------------------------------------------------ histogram path
@pytest.mark.parametrize("increase,lt", [(True, True), (True, False), (False, True), (False, False)])
def test_hist_matches_strict_except_ties(vort, increase, lt):
    """tests/test_hist.py:132-167 compares the two paths by eye; here: they agree
    everywhere except for cells sitting exactly on a level / the extreme cell."""
    lat, lon, q = vort
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    q3 = q[None]
    ctr = O.cal_contours(q3, 121, increase)
    a_h = O.cal_integral_within_contours_hist(q3, ctr[0], dA, lt)[0]
    a_s = O.cal_integral_within_contours(q3, ctr[0], dA, lt)[0].astype(np.float64)
    tot = dA.astype(np.float64).sum()
    assert np.abs(a_h - a_s).max() / tot < 2e-5
    # invariants
    d = np.diff(a_h)
    assert np.all(d >= 0) or np.all(d <= 0)
    assert abs(max(a_h[0], a_h[-1]) - tot) / tot < 1e-5


def test_digitize_rule_edges():
    e = np.array([0.0, 1.0, 2.0, 3.0])
    x = np.array([-0.1, 0.0, 0.5, 1.0, 2.999, 3.0, 3.1, np.nan])
    assert O.digitize_bins(x, e).tolist() == [-1, 0, 0, 1, 2, 2, -1, -1]   # last edge closed by +1e-8
    e32 = np.array([280.0, 290.0, 300.0], dtype=np.float32)                # +1e-8 is a no-op in fp32
    assert O.digitize_bins(np.array([300.0], dtype=np.float32), e32).tolist() == [-1]


def test_hist_edges_branches():
    ctr = np.linspace(1.0, 2.0, 6).astype(np.float32)
    e_t, inc = O.hist_edges(ctr, time_branch=True)
    e_s, _ = O.hist_edges(ctr, time_branch=False)
    assert inc and e_t.dtype == np.float64 and e_s.dtype == np.float32
    assert np.array_equal(e_t[1:], ctr.astype(np.float64)) and np.array_equal(e_s[1:], ctr)
    e_d, inc = O.hist_edges(ctr[::-1].copy(), time_branch=True)
    assert (not inc) and np.array_equal(e_d, e_t)


# ------------------------------------------------------------------ LWA
@pytest.mark.parametrize("increase", [True, False])
@pytest.mark.parametrize("part", ["all", "upper", "lower"])
def test_lwa_fast_equals_bruteforce(vort, increase, part):
    lat, lon, q = vort
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    q3 = q[None, ::2, ::4].copy()
    lat2, dA2 = lat[::2], dA[::2, ::4].copy()
    ctr = O.cal_contours(q3, 61, increase)
    area = O.cal_integral_within_contours_hist(q3, ctr[0], dA2, True)
    tbl, c = O.cal_area_eqCoord_table_hist(lat2, np.ones_like(q3[0]), dA2, 0, increase, True)
    latEq = O.table_lookup_coordinates(area, tbl, c)
    Q = O.interp_to_coords(lat2, latEq, ctr)
    L = O.cal_local_wave_activity(q3, Q, dA2, lat2, increase, part)
    Lf = O.cal_local_wave_activity_fast(q3, Q, dA2, lat2, increase, part)
    assert np.abs(L - Lf).max() <= 1e-12 * np.abs(L).max()
    assert (L >= 0).all() if increase else (L <= 0).all()


def test_lwa_notebook_range_old_semantics(vort):
    """notebooks/2.LWA_atmos.ipynb cell 5 plots LWA in 0..28.  Those figures were
    produced with the pre-refactor weights (area/area.max * dy, still visible in
    the comment at core.py:787-788); with them the restated algorithm lands in
    that range (SURVEY.md hazard H2)."""
    lat, lon, q = vort
    q3 = q[None]
    dA = O.latlon_cell_area(lat, lon)
    ctr = O.cal_contours(q3, 121, True)
    area = O.cal_integral_within_contours_hist(q3, ctr[0], dA.astype(np.float32), True)
    tbl, c = O.cal_area_eqCoord_table_hist(lat, np.ones_like(q), dA.astype(np.float32), 0, True, True)
    Q = O.interp_to_coords(lat, O.table_lookup_coordinates(area, tbl, c), ctr)
    # old semantics: qe*mask*wei*dy  ==  fast path with ww := wei*dy
    dy = np.gradient(np.deg2rad(lat.astype(np.float64))) * O.Rearth
    wei = dA / dA.max()
    ww_old = wei * dy[:, None]
    fake_dA = np.sqrt(ww_old * ww_old.max())           # (fake/max(fake))*fake == ww_old
    L = O.cal_local_wave_activity_fast(q3, Q, fake_dA, lat, True)[0]
    assert 27.0 < L.max() < 31.0 and L.min() > -1e-9 * L.max()


# ------------------------------------------------------------------ contour-space ops
def test_gradient_and_keff_shapes():
    rng = np.random.default_rng(1)
    area = np.cumsum(rng.random((2, 31)) + 0.1, axis=1)
    ctr = np.linspace(0, 1, 31, dtype=np.float32)[None].repeat(2, 0)
    dq = O.cal_gradient_wrt_area(ctr, area)
    assert dq.dtype == np.float64 and np.all(dq > 0)
    assert np.allclose(dq[:, 1:-1], (ctr[:, 2:] - ctr[:, :-2]) / 2 / ((area[:, 2:] - area[:, :-2]) / 2))
    L2 = O.cal_sqared_equivalent_length(dq, dq)
    nk = O.cal_normalized_Keff(L2, np.full_like(L2, 1e-3), mask=1e5)
    assert np.isnan(nk).any() or np.all(nk < 1e5)


def test_table_roundtrip(vort):
    lat, lon, q = vort
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    tbl, c = O.cal_area_eqCoord_table_hist(lat, np.ones_like(q), dA, 0, True, True)
    assert np.all(np.diff(tbl) > 0) and np.array_equal(c, lat)
    mid = 0.5 * (tbl[1:] + tbl[:-1])
    back = O.table_lookup_coordinates(mid[None], tbl, c)[0]
    assert np.all((back > lat[:-1]) & (back < lat[1:]))


# ------------------------------------------------------------------ C ABI surface
def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "xcb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(xc_[a-z0-9_]+)\s*\(", hdr)))


def test_library_builds_and_exports_every_declared_symbol():
    from xcontour_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    assert set(_lib.SIGNATURES) == set(syms)
    lib.xc_abi_version.restype = ctypes.c_int
    header = open(os.path.join(ROOT, "include", "xcb200.h")).read()
    assert lib.xc_abi_version() == int(re.search(r"#define XC_ABI_VERSION (\d+)", header).group(1)) == 3


def test_size_queries_and_argument_errors_without_gpu():
    from xcontour_b200 import _lib
    lib = _lib.load()
    assert lib.xc_minmax_levels_workspace_bytes(4, 721 * 1440) > 0
    assert lib.xc_bin_accumulate_workspace_bytes(4, 721 * 1440, 361, 2) > 0
    assert lib.xc_keff_lwa_batch_workspace_bytes(4, 721, 1440, 361) > 0
    rc = lib.xc_leq2(None, None, 0, None, None)          # argument check happens before any launch
    assert rc != 0 and b"xc_leq2" in lib.xc_last_error()
    with pytest.raises(Exception, match="xc_leq2"):
        _lib.check(rc)


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from xcontour_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.require_cuda()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "xcontour_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


# ------------------------------------------------------------------ host logic
def test_constructor_errors_and_labeled_arrays():
    import xcontour_b200 as xb
    da = xb.DataArray(np.zeros((2, 3, 4), np.float32), dims=("time", "lat", "lon"),
                      coords={"lat": np.arange(3.0), "lon": np.arange(4.0)}, name="q")
    with pytest.raises(Exception, match="dimEq should be one dimension"):
        xb.Contour2D(da, da, dims={"X": "lon", "Y": "lat"}, dimEq={"Y": "lat", "Z": "z"})
    with pytest.raises(Exception, match="dims should be a 2D plane"):
        xb.Contour2D(da, da, dims={"X": "lon"}, dimEq={"Y": "lat"})
    c = xb.Contour2D(da, da, dims={"X": "lon", "Y": "lat"}, dimEq={"Y": "lat"})
    assert c.dimVs == ["lon", "lat"] and c.dimEqV == "lat" and c.lt is False
    # labelled container behaves like the slice of xarray the workflow needs
    s = (da + 1).isel(time=0).rename("x")
    assert s.dims == ("lat", "lon") and s.name == "x" and float(s.values[0, 0]) == 1.0
    w = da.where(da["lat"] > 0)
    assert np.isnan(w.values[:, 0]).all() and not np.isnan(w.values[:, 1:]).any()
    assert (da.isel(time=0) * xb.DataArray(np.arange(3.0), dims=("lat",))).shape == (3, 4)


def test_slice_range_partition():
    from xcontour_b200.pipeline import slice_range
    for S in (1, 7, 8, 324120):
        for world in (1, 2, 4, 8):
            ranges = [slice_range(S, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == S
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))


def test_gather_contour_space_gloo_world2(tmp_path):
    """N>1 host logic on CPU: two gloo ranks own slice ranges and all-gather."""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, torch, torch.distributed as dist\n"
        "sys.path.insert(0, %r)\n"
        "from xcontour_b200.pipeline import slice_range, gather_contour_space\n"
        "dist.init_process_group('gloo')\n"
        "r, w, S, N = dist.get_rank(), dist.get_world_size(), 7, 5\n"
        "lo, hi = slice_range(S, r, w)\n"
        "full = torch.arange(S * N, dtype=torch.float64).reshape(S, N)\n"
        "out = gather_contour_space({'area': full[lo:hi].clone()}, S)\n"
        "assert torch.equal(out['area'], full), out\n"
        "dist.barrier(); print('OK', r)\n" % ROOT)
    import socket
    with socket.socket() as sk:                               # a port that is free right now
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("OK") == 2


def test_equal_area_levels_oracle_vs_exact_quantiles(vort):
    """The weighted-quantile-histogram levels (north_star kernel (1), oracle side)
    sit within a fraction of a fine bin of the exact weighted quantiles by sorting."""
    lat, lon, q = vort
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    N, refine = 41, 8
    for increase, lt in ((True, True), (False, False), (True, False), (False, True)):
        lev = O.cal_contours_equal_area(q[None], dA, N, increase, lt, np.float32, refine)[0]
        step = (q.max() - q.min()) / ((N - 1) * refine)
        fr = np.linspace(0, 1, N)
        exact = O.weighted_quantile_levels(q, dA, fr if increase else 1 - fr)
        assert np.abs(lev[1:-1] - exact[1:-1]).max() <= 0.5 * step
        assert np.all(np.diff(lev) > 0) if increase else np.all(np.diff(lev) < 0)


def test_product_cell_area_helper_matches_oracle():
    from xcontour_b200.utils import latlon_cell_area
    lat = np.linspace(-90, 90, 73); lon = np.arange(144) * 2.5
    assert np.array_equal(latlon_cell_area(lat, lon), O.latlon_cell_area(lat, lon))
    assert np.array_equal(latlon_cell_area(lat[::-1], lon), O.latlon_cell_area(lat[::-1], lon))
    assert abs(latlon_cell_area(lat, lon).sum() / (4 * np.pi * O.Rearth ** 2) - 1) < 1e-12


def test_general_stencil_reduces_to_the_latlon_statement_and_pads_like_numpy():
    """oracle.squared_gradient (ghost cells by np.pad) == squared_gradient_latlon for (periodic, extend) with
    lat-lon metrics; spot values of the other ghost-cell rules."""
    rng = np.random.default_rng(5)
    lat = np.linspace(-90, 90, 19); lon = np.arange(36) * 10.0
    q = rng.standard_normal((2, 19, 36)).astype(np.float32)
    cx, cy = O.row_metrics_latlon(lat, lon)
    assert np.array_equal(O.squared_gradient(q, cx, cy, "periodic", "extend"), O.squared_gradient_latlon(q, lat, lon))
    y = np.arange(19) * 2.0; x = np.arange(36) * 0.5
    cx, cy = O.row_metrics_cartesian(y, x)
    q64 = q.astype(np.float64)
    g = O.squared_gradient(q, cx, cy, "reflect", "fill", fill=2.0)
    assert g[0, 3, 0] == ((q64[0, 3, 1] - q64[0, 3, 1]) * cx[3]) ** 2 + ((q64[0, 4, 0] - q64[0, 2, 0]) * cy[3]) ** 2
    assert g[1, 0, 5] == ((q64[1, 0, 6] - q64[1, 0, 4]) * cx[0]) ** 2 + ((q64[1, 1, 5] - 2.0) * cy[0]) ** 2
    g = O.squared_gradient(q, cx, cy, "extend", "periodic")
    assert g[0, 18, 35] == ((q64[0, 18, 35] - q64[0, 18, 34]) * cx[18]) ** 2 + ((q64[0, 0, 35] - q64[0, 17, 35]) * cy[18]) ** 2


def test_gradient_wrt_area_uses_the_contour_coordinate():
    """np.gradient against non-uniform level values (what differentiate('contour') does after cal_contours(array),
    core.py:264, 480-483) is not the unit-spacing quotient; against 0..N-1 it is."""
    rng = np.random.default_rng(9)
    lev = np.cumsum(0.5 + rng.random(12))
    f = np.sin(lev); a = np.cumsum(rng.random(12))
    got = O.cal_gradient_wrt_area(f, a, var_coord=lev, area_coord=lev)
    hd, hs = lev[2:] - lev[1:-1], lev[1:-1] - lev[:-2]
    def g(y):
        return (hs ** 2 * y[2:] + (hd ** 2 - hs ** 2) * y[1:-1] - hd ** 2 * y[:-2]) / (hs * hd * (hd + hs))
    assert np.allclose(got[1:-1], g(f) / g(a), rtol=1e-12)
    assert np.array_equal(O.cal_gradient_wrt_area(f, a), O.cal_gradient_wrt_area(f, a, var_coord=np.arange(12.0), area_coord=np.arange(12.0)))
    assert not np.allclose(got[1:-1], O.cal_gradient_wrt_area(f, a)[1:-1], rtol=1e-3)


def test_workspace_planning_terminates_for_every_benchmark_shape():
    """Host-side planning (no GPU needed): xc_keff_lwa_batch_workspace_bytes walks the row-march planner with the
    worst-case flags; at config 5 (N = 2048) no row count fits the two-CTA shared-memory budget for row-dependent
    areas and the shrink loop once stalled at 4 rows."""
    from xcontour_b200 import _lib
    lib = _lib.load()
    for args in ((1, 4096, 8192, 2048), (4, 4096, 8192, 2048), (32, 721, 1440, 361), (64, 721, 1440, 361),
                 (3, 83, 152, 47), (1, 8, 8, 2048), (5, 2, 8, 4), (100000, 721, 1440, 361)):
        assert lib.xc_keff_lwa_batch_workspace_bytes(*args) > 0


def test_lead_dims_are_paired_by_name_not_by_position():
    """interp_to_coords / cal_gradient_wrt_area pair slices the way xarray broadcasts them (core.py:1089-1097 runs
    through apply_ufunc): by dimension name, union of the lead dims in order of first appearance."""
    from xcontour_b200 import DataArray
    from xcontour_b200.core import Contour2D
    rng = np.random.default_rng(3)
    T, L, N = 3, 2, 5
    e = DataArray(rng.random((T, N)), dims=('time', 'contour'), coords={'time': np.arange(T)})
    v = DataArray(rng.random((L, T, N)), dims=('lev', 'time', 'contour'), coords={'lev': [10., 20.]})
    e2, v2, lead, lshape = Contour2D._align2(e, v)
    assert lead == ['time', 'lev'] and lshape == (T, L)
    assert e2.shape == (T * L, N) and v2.shape == (T * L, N)
    for t in range(T):
        for l in range(L):
            assert np.array_equal(e2[t * L + l], e.values[t])          # repeated along the dim it lacks
            assert np.array_equal(v2[t * L + l], v.values[l, t])       # transposed into the common order
    # same sizes, different dim order: positions would pair the wrong slices
    a = DataArray(rng.random((2, 2, N)), dims=('time', 'lev', 'contour'))
    b = DataArray(rng.random((2, 2, N)), dims=('lev', 'time', 'contour'))
    a2, b2, lead, _ = Contour2D._align2(a, b)
    assert lead == ['time', 'lev'] and np.array_equal(b2[1], b.values[1, 0]) and np.array_equal(a2[1], a.values[0, 1])
    with pytest.raises(Exception, match="cannot align"):
        Contour2D._align2(e, DataArray(rng.random((T + 1, N)), dims=('time', 'contour')))


def test_default_scalar_rules_follow_the_installed_numpy():
    """The per-'time' bin edges of _histogram depend on NumPy's scalar promotion (core.py:1277-1278); by default the
    product reproduces what the reference computes under the NumPy that is installed, and the switch is explicit."""
    code = ("import numpy as np; from xcontour_b200 import utils; "
            "print(utils.NUMPY_SCALAR_RULES, int(np.__version__.split('.')[0]))")
    env = {k: v for k, v in os.environ.items() if k != "XCB200_NUMPY_RULES"}
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, check=True).stdout.split()
    assert out[0] == ("numpy2" if int(out[1]) >= 2 else "numpy1")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(env, XCB200_NUMPY_RULES="numpy1"),
                         capture_output=True, text=True, check=True).stdout.split()
    assert out[0] == "numpy1"
    from xcontour_b200 import utils
    assert utils.scalar_rules("numpy2") == "numpy2" and utils.scalar_rules() == "numpy1"      # conftest pins numpy1
    with pytest.raises(Exception, match="scalar rules"):
        utils.scalar_rules("numpy3")


def test_contour_gather_packed_gloo_world2(tmp_path):
    """The collective of the N > 1 path itself (pipeline.ContourGather: one packed [9, S_local, N] buffer per rank,
    rotating receive buffers) on two gloo ranks: layout [world, 9, S_local, N], unpack() by name with the padded
    tail trimmed, a second batch landing in the other receive buffer while the first stays intact."""
    script = tmp_path / "g.py"
    script.write_text(
        "import sys, torch, torch.distributed as dist\n"
        "sys.path.insert(0, %r)\n"
        "from xcontour_b200.pipeline import ContourGather, CONTOUR_VARS, slice_range\n"
        "dist.init_process_group('gloo')\n"
        "r, w, S, N = dist.get_rank(), dist.get_world_size(), 7, 5\n"
        "per = (S + w - 1) // w\n"
        "lo, hi = slice_range(S, r, w)\n"
        "def full(b):\n"
        "    return torch.arange(9 * S * N, dtype=torch.float64).reshape(9, S, N) + 1000.0 * b\n"
        "g = ContourGather(per, N, 'cpu', nbuf=2)\n"
        "keep = []\n"
        "for b in range(3):\n"
        "    packed = torch.zeros((9, per, N), dtype=torch.float64)\n"
        "    packed[:, :hi - lo] = full(b)[:, lo:hi]\n"
        "    recv, idx = g.launch(packed)\n"
        "    assert idx == b %% 2 and tuple(recv.shape) == (w, 9, per, N)\n"
        "    out = ContourGather.unpack(recv, S)\n"
        "    assert list(out) == list(CONTOUR_VARS)\n"
        "    for i, name in enumerate(CONTOUR_VARS):\n"
        "        assert torch.equal(out[name], full(b)[i]), (b, name)\n"
        "    keep.append((recv, b))\n"
        "    if b == 1:\n"
        "        assert torch.equal(ContourGather.unpack(keep[0][0], S)['area'], full(0)[1])\n"
        "g.wait()\n"
        "try:\n"
        "    g.launch(torch.zeros((9, per + 1, N), dtype=torch.float64)); raise SystemExit('no error')\n"
        "except Exception as e:\n"
        "    assert 'packed buffer' in str(e)\n"
        "dist.barrier(); print('OK', r)\n" % ROOT)
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("OK") == 2


def test_bench_contract_pieces_checkable_without_a_gpu():
    """bench.py: the synthetic field is SURVEY 8(d)'s (and the one the parity tests use), both arms quote the same
    metric / workload strings, the reference arm's line carries the keys the driver reads, and without a CUDA device
    our arm fails loudly instead of measuring something else."""
    import bench
    lat, lon = bench.grid()
    _, _, q = synth_c4(1)
    assert np.array_equal(bench.synth_slice_np(0, lat, lon), q[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert (bench.NY, bench.NX, bench.NLEV) == (721, 1440, 361) and "721x1440" in bench.METRIC
    assert "slices" in base["metric"].lower() and "721" in base["metric"]
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"workload": WORKLOAD') + src.count('"workload": C5_WORKLOAD if c5 else WORKLOAD') >= 2
    for key in ('"impl": "reference"', '"cpu_baseline"', '"e2e"', '"roofline"', '"clocks"', '"gpu_launches"',
                '"higher_is_better"', '"scaling"', '"vs_baseline"', '"ms_per_step"'):
        assert key in src, key
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode != 0 and "no CPU fallback" in (res.stderr + res.stdout)
