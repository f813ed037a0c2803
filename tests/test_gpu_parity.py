"""
GPU parity suite (-m gpu): every entry point of the C ABI, called the way the
product calls it (ctypes -> libxcb200.so, device buffers from torch), against
the NumPy oracle on the same seeded inputs and on the committed fixtures.

Bars (BASELINE.json north_star): contour bin assignment bit-exact given identical
fp64 levels; levels / edges / np.interp / d/dA bit-exact; integrals and
Keff / LWA / LAPE fields within 1e-10 (relative to the field maximum -- the
tolerance is written next to each assertion).
"""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from conftest import GOLDEN, synth_c4
from oracle import xcontour_oracle as O

pytestmark = pytest.mark.gpu

RTOL_INT = 1e-12      # integrals (fp64 accumulation, different summation order)
RTOL_FIELD = 1e-10    # Keff / LWA / LAPE fields, relative to the field max


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xcontour_b200 import ops as _ops
    _ops.require_cuda()
    return _ops


def dev(ops, a):
    return ops.to_dev(np.ascontiguousarray(a))


def relmax(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = np.nanmax(np.abs(b)) if np.isfinite(np.nanmax(np.abs(b))) and np.nanmax(np.abs(b)) > 0 else 1.0
    return np.nanmax(np.abs(a - b)) / scale


def same_nan(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b))


# ---------------------------------------------------------------- (1) levels
def test_levels_golden_notebook_on_gpu(ops):
    g = json.load(open(os.path.join(GOLDEN, "contours_pv.json")))
    rows = np.array([[np.float32(x) for x in r] for r in g["printed"]])
    q = np.stack([rows[:, 0], rows[:, -1]], axis=1).astype(np.float32)       # [6, P=2]
    lv, mm = ops.minmax_levels(dev(ops, q), g["levels_N"], True, 0)
    got = lv.cpu().numpy().astype(np.float32)[:, g["columns"]]
    assert np.array_equal(got, rows)
    assert np.array_equal(mm.cpu().numpy(), q.astype(np.float64))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("increase", [True, False])
@pytest.mark.parametrize("shape", [(5, 37, 53), (3, 64, 128), (2, 721, 1440)])
def test_levels_bit_exact(ops, dtype, increase, shape):
    rng = np.random.default_rng(7)
    q = (rng.standard_normal(shape) * 3e-5 + 1e-4).astype(dtype)
    q[0, 0, :5] = np.nan
    q[1, -1, -1] = np.nan
    for out_dt, code in ((np.float32, 0), (np.float64, 1)):
        ref = O.cal_contours(q, 121, increase, out_dt)
        lv, _ = ops.minmax_levels(dev(ops, q.reshape(shape[0], -1)), 121, increase, code)
        assert np.array_equal(lv.cpu().numpy().astype(out_dt), ref)


@pytest.mark.parametrize("cdt", [np.float32, np.float64])
@pytest.mark.parametrize("time_branch", [True, False])
def test_hist_edges_bit_exact(ops, cdt, time_branch):
    rng = np.random.default_rng(3)
    q = (rng.standard_normal((4, 20, 30)) * 2e-4).astype(np.float32)
    for increase in (True, False):
        ctr = O.cal_contours(q, 33, increase, cdt)
        e, d = ops.hist_edges(dev(ops, ctr.astype(np.float64)), 0 if cdt == np.float32 else 1, time_branch)
        e, d = e.cpu().numpy(), d.cpu().numpy()
        for s in range(4):
            ref, binc = O.hist_edges(ctr[s], time_branch)
            nudged = np.concatenate((ref[:-1], ref[-1:] + 1e-8)).astype(ref.dtype)
            assert np.array_equal(e[s], nudged.astype(np.float64))
            assert d[s] == (0 if binc else 1)


# ---------------------------------------------------------------- (2) binning
def _edges_for(ctr, time_branch):
    ref, binc = O.hist_edges(ctr, time_branch)
    return np.concatenate((ref[:-1], ref[-1:] + 1e-8)).astype(ref.dtype).astype(np.float64), binc


@pytest.mark.parametrize("qdt", [np.float32, np.float64])
def test_bin_assignment_bit_exact(ops, vort, qdt):
    lat, lon, q = vort
    q = q.astype(qdt)
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    q3 = np.stack([q, q[::-1], q * 0.5])
    ctr = O.cal_contours(q3, 121, True)
    # plant cells exactly on levels, NaN, and out-of-range values
    q3[0, 10, :121] = ctr[0].astype(qdt)
    q3[1, 5, :7] = np.nan
    q3[2, 0, 0] = 1.0
    q3[2, 0, 1] = -1.0
    edges = np.stack([_edges_for(ctr[s], True)[0] for s in range(3)])
    cdf, pdf, idx = ops.bin_accumulate(dev(ops, q3.reshape(3, -1)), dev(ops, edges), dev(ops, dA.reshape(-1)),
                                       want_pdf=True, want_idx=True)
    idx = idx.cpu().numpy()
    for s in range(3):
        e, _ = O.hist_edges(ctr[s], True)
        ref = O.digitize_bins(q3[s].ravel(), e)
        assert np.array_equal(idx[s], ref)                        # bit-exact bin assignment
        refpdf = O.xhistogram_1d(q3[s], e, dA)
        assert relmax(pdf[s, 0].cpu().numpy(), refpdf) <= RTOL_INT


@pytest.mark.parametrize("increase,lt", [(True, True), (True, False), (False, True), (False, False)])
def test_cdf_hist_path(ops, vort, increase, lt):
    lat, lon, q = vort
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    rng = np.random.default_rng(11)
    q3 = np.stack([q, q[::-1]])
    g = np.abs(rng.standard_normal(q3.shape)).astype(np.float32) * 1e-9
    g[0, 3, 3] = np.nan                                           # fillna(0) on the weights
    ctr = O.cal_contours(q3, 121, increase)
    ref_a = O.cal_integral_within_contours_hist(q3, ctr, dA, lt)
    ref_g = O.cal_integral_within_contours_hist(q3, ctr, dA, lt, integrand=g)
    e, d = ops.hist_edges(dev(ops, ctr.astype(np.float64)), 0, True)
    from xcontour_b200._lib import SCAN_PREFIX, SCAN_TOTAL_MINUS
    cdf, _, _ = ops.bin_accumulate(dev(ops, q3.reshape(2, -1)), e, dev(ops, dA.reshape(-1)),
                                   acc_area=True, integrands=[dev(ops, g.reshape(2, -1))],
                                   scan_mode=SCAN_PREFIX if lt else SCAN_TOTAL_MINUS, decreasing=d)
    cdf = cdf.cpu().numpy()
    assert relmax(cdf[:, 0], ref_a) <= RTOL_INT
    assert relmax(cdf[:, 1], ref_g) <= RTOL_INT
    # run-to-run determinism (fixed reduction order, no atomics on this path)
    cdf2, _, _ = ops.bin_accumulate(dev(ops, q3.reshape(2, -1)), e, dev(ops, dA.reshape(-1)),
                                    acc_area=True, integrands=[dev(ops, g.reshape(2, -1))],
                                    scan_mode=SCAN_PREFIX if lt else SCAN_TOTAL_MINUS, decreasing=d)
    assert np.array_equal(cdf, cdf2.cpu().numpy())


def test_cdf_many_bins_shared_copy_path(ops):
    """N = 2048 with K = 2 does not fit 8 private copies -> shared-copy atomics."""
    rng = np.random.default_rng(5)
    q = rng.random((2, 128, 512)).astype(np.float32)
    dA = (rng.random((128, 512)) + 0.5)
    g = rng.random((2, 128, 512))
    ctr = O.cal_contours(q, 2048, True, np.float64)
    ref_a = O.cal_integral_within_contours_hist(q, ctr, dA, True)
    ref_g = O.cal_integral_within_contours_hist(q, ctr, dA, True, integrand=g)
    e, d = ops.hist_edges(dev(ops, ctr), 1, True)
    cdf, _, _ = ops.bin_accumulate(dev(ops, q.reshape(2, -1)), e, dev(ops, dA.reshape(-1)), acc_area=True,
                                   integrands=[dev(ops, g.reshape(2, -1))], decreasing=d)
    cdf = cdf.cpu().numpy()
    assert relmax(cdf[:, 0], ref_a) <= RTOL_INT and relmax(cdf[:, 1], ref_g) <= RTOL_INT


def test_cdf_ragged_and_tiny(ops):
    rng = np.random.default_rng(9)
    for shape, N in (((1, 1, 3), 2), ((3, 7, 13), 5), ((2, 1, 131), 17)):
        q = rng.standard_normal(shape).astype(np.float32)
        dA = (rng.random(shape[1:]) + 1).astype(np.float32)
        ctr = O.cal_contours(q, N, True)
        ref = O.cal_integral_within_contours_hist(q, ctr, dA, True, time_branch=True)
        e, d = ops.hist_edges(dev(ops, ctr.astype(np.float64)), 0, True)
        cdf, _, _ = ops.bin_accumulate(dev(ops, q.reshape(shape[0], -1)), e, dev(ops, dA.reshape(-1)), decreasing=d)
        assert relmax(cdf[:, 0].cpu().numpy(), ref) <= RTOL_INT


# ---------------------------------------------------------------- (3)-(5) contour space
def test_interp_bit_exact(ops):
    rng = np.random.default_rng(2)
    S, n, M = 6, 57, 91
    xp = np.cumsum(rng.random((S, n)), axis=1)
    xp[2, 10:14] = xp[2, 10]                                       # ties in the table
    fp = rng.standard_normal((S, n))
    x = rng.random((S, M)) * (xp[:, -1:] + 2) - 1                  # includes out-of-range
    x[0, :n] = xp[0]                                               # exact hits
    x[1, 0] = np.nan
    out = ops.interp(dev(ops, x), dev(ops, xp), dev(ops, fp), reverse=0).cpu().numpy()
    ref = np.stack([np.interp(x[s], xp[s], fp[s]) for s in range(S)])
    assert np.array_equal(out, ref, equal_nan=True)
    out = ops.interp(dev(ops, x), dev(ops, xp[:, ::-1].copy()), dev(ops, fp[:, ::-1].copy()), reverse=-1).cpu().numpy()
    assert np.array_equal(out, ref, equal_nan=True)                # auto-detected decreasing tables
    out = ops.interp(dev(ops, x[0]), dev(ops, xp), dev(ops, fp), reverse=0).cpu().numpy()
    assert np.array_equal(out, np.stack([np.interp(x[0], xp[s], fp[s]) for s in range(S)]), equal_nan=True)


def test_gradient_and_keff_epilogue(ops):
    rng = np.random.default_rng(4)
    S, N = 5, 121
    area = np.cumsum(rng.random((S, N)) * 1e12, axis=1)
    area[1, 40:43] = area[1, 40]                                   # empty bins -> 0/0, x/0
    ctr = O.cal_contours(rng.standard_normal((S, 9, 9)).astype(np.float32), N, True)
    intg = np.cumsum(rng.random((S, N)), axis=1)
    from xcontour_b200._lib import XC_F32, XC_F64, XC_F32_AS_F64
    with np.errstate(all="ignore"):
        ref_dq = O.cal_gradient_wrt_area(ctr, area)
        ref_dg = O.cal_gradient_wrt_area(intg, area)
        ref_l2 = O.cal_sqared_equivalent_length(ref_dg, ref_dq)
    dq = ops.gradient_wrt_area(dev(ops, ctr), XC_F32, dev(ops, area), XC_F64)
    dq2 = ops.gradient_wrt_area(dev(ops, ctr.astype(np.float64)), XC_F32_AS_F64, dev(ops, area), XC_F64)
    dg = ops.gradient_wrt_area(dev(ops, intg), XC_F64, dev(ops, area), XC_F64)
    assert np.array_equal(dq.cpu().numpy(), ref_dq, equal_nan=True)
    assert np.array_equal(dq2.cpu().numpy(), ref_dq, equal_nan=True)
    assert np.array_equal(dg.cpu().numpy(), ref_dg, equal_nan=True)
    l2 = ops.leq2(dg, dq).cpu().numpy()
    assert np.array_equal(l2, ref_l2, equal_nan=True)
    lat = rng.uniform(-90, 90, (S, N))
    ref_lm = O.latitude_lengths_at(lat)
    lm = ops.lmin(dev(ops, lat)).cpu().numpy()
    assert np.allclose(lm, ref_lm, rtol=4e-16, atol=1e-9)          # cos(): <= 1 ulp between libms
    ref_nk = O.cal_normalized_Keff(ref_l2, ref_lm)
    nk = ops.nkeff(dev(ops, ref_l2), dev(ops, ref_lm), 1e5).cpu().numpy()
    assert np.array_equal(nk, ref_nk, equal_nan=True)
    a = rng.random((S, N)) * 5.1e14
    assert np.allclose(ops.eqlat(dev(ops, a)).cpu().numpy(), O.equivalent_latitudes(a), rtol=1e-14, atol=1e-12)
    # fp32 / fp32 quotient is rounded in fp32 like NumPy's
    a32 = area.astype(np.float32)
    r = ops.gradient_wrt_area(dev(ops, ctr), XC_F32, dev(ops, a32), XC_F32).cpu().numpy().astype(np.float32)
    with np.errstate(all="ignore"):
        assert np.array_equal(r, O.cal_gradient_wrt_area(ctr, a32), equal_nan=True)


# ---------------------------------------------------------------- (6) LWA
def _sorted_profile(q3, lat, dA, N, increase, lt=True):
    ctr = O.cal_contours(q3, N, increase)
    area = O.cal_integral_within_contours_hist(q3, ctr, dA, lt)
    tbl, c = O.cal_area_eqCoord_table_hist(lat, np.ones_like(q3[0]), dA, 0, increase, lt)
    latEq = O.table_lookup_coordinates(area, tbl, c)
    return O.interp_to_coords(lat, latEq, ctr)


@pytest.mark.parametrize("increase", [True, False])
@pytest.mark.parametrize("part", ["all", "upper", "lower"])
def test_lwa_vs_reference_loop(ops, vort, increase, part):
    lat, lon, q = vort
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    q3 = np.stack([q, q[::-1]])
    Q = _sorted_profile(q3, lat, dA, 121, increase)
    ref = O.cal_local_wave_activity(q3, Q, dA, lat, increase, part)
    ww = ops.lwa_weights(dev(ops, dA.reshape(-1)))
    out = ops.lwa(dev(ops, q3), dev(ops, Q), ww, increase, part, 1).cpu().numpy()
    assert relmax(out, ref) <= RTOL_FIELD * 1e-2                    # measured ~1e-15
    assert (out >= -1e-12 * np.abs(ref).max()).all() if increase else (out <= 1e-12 * np.abs(ref).max()).all()


def test_lwa_unsorted_profile_nan_and_fp64(ops, vort):
    """A profile that is not sorted (or has NaN) takes the exact O(n^2) kernel;
    NaN tracer cells contribute nothing (skipna sum, core.py:1376)."""
    lat, lon, q = vort
    dA = O.latlon_cell_area(lat, lon)
    q3 = np.stack([q[::4, ::8], q[::4, ::8][::-1]]).astype(np.float64)
    lat4, dA4 = lat[::4], dA[::4, ::8].copy()
    q3[0, 5:9, 3:11] = np.nan
    Q = _sorted_profile(np.nan_to_num(q3), lat4, dA4, 41, True)
    Q[1] = Q[1][np.random.default_rng(0).permutation(Q.shape[1])]  # scrambled -> brute force
    Q[0, 7] = Q[0, 8]                                               # a tie stays on the fast path
    ref = O.cal_local_wave_activity(q3, Q, dA4, lat4, True)
    ww = ops.lwa_weights(dev(ops, dA4.reshape(-1)))
    out = ops.lwa(dev(ops, q3), dev(ops, Q), ww, True, "all", 1).cpu().numpy()
    assert relmax(out, ref) <= RTOL_FIELD * 1e-2
    Q[0, 3] = np.nan
    ref = O.cal_local_wave_activity(q3, Q, dA4, lat4, True)
    out = ops.lwa(dev(ops, q3), dev(ops, Q), ww, True, "all", 1).cpu().numpy()
    assert relmax(out, ref) <= RTOL_FIELD * 1e-2 and np.all(out[0, 3] == 0)


@pytest.mark.parametrize("increase", [True, False])
def test_lwa_fixed_point_ties_and_homogenised_patches(ops, increase):
    """Fixed-point LWA kernel on the inputs that stress its exact parts: a tracer
    quantised to few distinct values (dozens of cells of a column share one slot of
    the difference array), tracer values that hit profile values exactly, flat
    stretches in the profile, NaN cells and NaN weights -- against the reference
    j-loop.  Integer accumulation is order-independent: two runs are bit-identical."""
    rng = np.random.default_rng(5)
    ny, nx = 97, 53                                                  # ragged: 53 = 3 tiles of 16 + 5
    y = np.linspace(-1, 1, ny)[:, None]
    q = y + 0.4 * np.sin(np.linspace(0, 12.56, nx))[None, :] * (1 - y ** 2) + 0.03 * rng.standard_normal((ny, nx))
    q = (np.round(q * 16) / 16).astype(np.float32)                   # homogenised patches
    q3 = np.stack([q, q[::-1], -q])
    dA = (0.5 + rng.random((ny, nx))).astype(np.float64)
    dA[11, 7] = np.nan                                               # NaN weight: that cell contributes nothing
    Q = np.sort(rng.choice(np.unique(q3), size=(3, ny)).astype(np.float64), axis=1)   # exact hits + flat stretches
    if not increase:
        Q = Q[:, ::-1].copy()
    q3[1, 40:44, 10:20] = np.nan
    coord = np.arange(ny, dtype=np.float64)
    for part in ("all", "upper", "lower"):
        ref = O.cal_local_wave_activity(q3, Q, dA, coord, increase, part)
        ww = ops.lwa_weights(dev(ops, dA.reshape(-1)))
        out = ops.lwa(dev(ops, q3), dev(ops, Q), ww, increase, part, 1)
        assert relmax(out.cpu().numpy(), ref) <= RTOL_FIELD * 1e-2
        out2 = ops.lwa(dev(ops, q3), dev(ops, Q), ww, increase, part, 1)
        assert torch.equal(out, out2)


def test_lwa_fixed_point_hands_infinities_to_the_exact_loop(ops, vort):
    """A slice with an infinite tracer value cannot be scaled to fixed point: the
    preparation kernel flags it and the exact O(n^2) kernel computes it, the other
    slices stay on the fast path."""
    lat, lon, q = vort
    dA = O.latlon_cell_area(lat, lon)[::4, ::8].copy()
    q3 = np.stack([q[::4, ::8], q[::4, ::8][::-1], q[::4, ::8]]).astype(np.float32)
    lat4 = lat[::4]
    Q = _sorted_profile(q3, lat4, dA, 41, True)
    q3[1, 20, 5] = np.inf
    with np.errstate(invalid="ignore"):
        ref = O.cal_local_wave_activity(q3, Q, dA, lat4, True)
    ww = ops.lwa_weights(dev(ops, dA.reshape(-1)))
    out = ops.lwa(dev(ops, q3), dev(ops, Q), ww, True, "all", 1).cpu().numpy()
    for s in (0, 2):
        assert relmax(out[s], ref[s]) <= RTOL_FIELD * 1e-2
    fin = np.isfinite(ref[1])
    assert np.array_equal(np.isfinite(out[1]), fin)
    assert relmax(out[1][fin], ref[1][fin]) <= RTOL_FIELD * 1e-2
    assert np.array_equal(np.isinf(out[1]), np.isinf(ref[1])) and np.array_equal(np.isnan(out[1]), np.isnan(ref[1]))


def test_fixed_point_and_fp64_accumulators_agree(ops):
    """Cross-check inside the product: the exact integer accumulators (default: bin_rows.cu,
    k_lwa_cols) and the fp64 read-modify-write kernels (XCB200_LWA_FX=0, XCB200_NO_BIN_ROWS=1:
    k_lwa_fast, k_hist) are two implementations of the same sums; a fresh process with the
    switches off must reproduce the fused batch within the summation-order tolerance."""
    import subprocess, sys, tempfile
    from conftest import ROOT
    code = (
        "import sys, numpy as np, torch; sys.path.insert(0, %r)\n"
        "import bench\n"
        "from xcontour_b200.pipeline import KeffLwaPlan\n"
        "from xcontour_b200 import utils\n"
        "lat, lon = bench.grid()\n"
        "dA = utils.latlon_cell_area(lat, lon).astype(np.float32)\n"
        "plan = KeffLwaPlan(lat, lon, dA, bench.NLEV)\n"
        "q = torch.from_numpy(np.stack([bench.synth_slice_np(k, lat, lon) for k in range(3)])).cuda()\n"
        "out = plan.run(q)\n"
        "torch.cuda.synchronize()\n"
        "np.savez(sys.argv[1], **{k: v.cpu().numpy() for k, v in out.items()})\n" % ROOT)
    res = []
    with tempfile.TemporaryDirectory() as td:
        for i, env in enumerate(({}, {"XCB200_LWA_FX": "0", "XCB200_NO_BIN_ROWS": "1"})):
            f = os.path.join(td, "o%d.npz" % i)
            r = subprocess.run([sys.executable, "-c", code, f], env=dict(os.environ, **env),
                               capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
            res.append(dict(np.load(f)))
    a, b = res
    assert np.array_equal(a["ctr"], b["ctr"])
    assert relmax(a["area"], b["area"]) <= RTOL_INT and relmax(a["intgrdS"], b["intgrdS"]) <= 1e-11
    _close(a["dqdA"], b["dqdA"], 1e-10)
    _close(a["Qref"], b["Qref"], 1e-11)
    assert relmax(a["lwa"], b["lwa"]) <= RTOL_FIELD * 1e-2


def test_lwa_variant2_and_masks(ops, vort):
    lat, lon, q = vort
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    q3 = q[None, ::2, ::4].copy()
    lat2, dA2 = lat[::2], dA[::2, ::4].copy()
    for increase in (True, False):
        Q = _sorted_profile(q3, lat2, dA2, 61, increase)
        ww = ops.lwa_weights(dev(ops, dA2.reshape(-1)))
        for variant in (1, 2):
            ref, ctrs, masks = O.cal_local_wave_activity(q3, Q, dA2, lat2, increase, "all",
                                                         mask_idx=[3, 60, 100], variant=variant)
            out = ops.lwa(dev(ops, q3), dev(ops, Q), ww, increase, "all", variant).cpu().numpy()
            assert relmax(out, ref) <= RTOL_FIELD * 1e-2
            for j, m in zip([3, 60, 100], masks):
                got = ops.lwa_mask(dev(ops, q3), dev(ops, Q), j, increase, variant).cpu().numpy()
                assert np.array_equal(got, m)                      # integer masks bit-exact
            for part in ("upper", "lower"):
                refp = O.cal_local_wave_activity(q3, Q, dA2, lat2, increase, part, variant=variant)
                outp = ops.lwa(dev(ops, q3), dev(ops, Q), ww, increase, part, variant).cpu().numpy()
                assert relmax(outp, refp) <= RTOL_FIELD * 1e-2
        # variant 2 with an unsorted profile and NaN cells takes the exact kernel
        Qs = Q.copy(); Qs[0, 5], Qs[0, 50] = Qs[0, 50], Qs[0, 5]
        q4 = q3.copy(); q4[0, 7, 3:9] = np.nan
        ref = O.cal_local_wave_activity(q4, Qs, dA2, lat2, increase, "all", variant=2)
        out = ops.lwa(dev(ops, q4), dev(ops, Qs), ww, increase, "all", 2).cpu().numpy()
        assert relmax(out, ref) <= RTOL_FIELD * 1e-2
        ref = O.cal_local_wave_activity(q4, Q, dA2, lat2, increase, "all", variant=2)
        out = ops.lwa(dev(ops, q4), dev(ops, Q), ww, increase, "all", 2).cpu().numpy()
        assert relmax(out, ref) <= RTOL_FIELD * 1e-2


def test_lape_xz_plane_with_topography(ops):
    """Config 3 stand-in (Data/internalwave.nc is missing from the mount): X-Z
    plane, increase=False, NaN topography, time-varying contours."""
    rng = np.random.default_rng(21)
    nz, nx, T = 100, 448, 3
    z = -np.arange(nz) * 2.0 - 1.0                                   # descending coordinate
    x = np.arange(nx) * 20.0
    b = np.empty((T, nz, nx), np.float32)
    for t in range(T):
        eta = 8.0 * np.sin(2 * np.pi * x / 3000.0 + t)[None, :] * np.exp(-((z[:, None] + 80) / 60.0) ** 2)
        b[t] = (2e-4 * 9.81 * (10.0 * np.tanh((z[:, None] - eta + 100) / 40.0)) +
                1e-5 * rng.standard_normal((nz, nx))).astype(np.float32)
    b[:, 80:, 300:] = np.nan                                         # topography
    dA = np.full((nz, nx), 40.0, np.float32)
    mask = (~np.isnan(b[0])).astype(np.float32)
    ctr = O.cal_contours(b, 121, increase=False)
    area = O.cal_integral_within_contours_hist(b, ctr, dA, False)
    tbl, c = O.cal_area_eqCoord_table_hist(z.astype(np.float32), mask, dA, 0, False, False)
    zEq = O.table_lookup_coordinates(area, tbl, c)
    Q = O.interp_to_coords(z.astype(np.float32), zEq, ctr)
    ref = O.cal_local_wave_activity(b, Q, dA, z, False)
    ww = ops.lwa_weights(dev(ops, dA.reshape(-1)))
    out = ops.lwa(dev(ops, b), dev(ops, Q), ww, False, "all", 1).cpu().numpy()
    assert relmax(out, ref) <= RTOL_FIELD * 1e-2
    assert (out <= 0).all()                                          # LAPE is plotted as -lape >= 0


# ---------------------------------------------------------------- (7) stencil
def test_grad2_latlon(ops):
    lat, lon, q = synth_c4(2, 91, 180)
    ref = O.squared_gradient_latlon(q, lat, lon)
    lat_rad = np.deg2rad(lat.astype(np.float64))
    lam = np.deg2rad(lon.astype(np.float64))
    out = ops.grad2_latlon(dev(ops, q), dev(ops, lat_rad), float(lam[1] - lam[0])).cpu().numpy()
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(out), fin)
    assert np.allclose(out[fin], ref[fin], rtol=1e-13, atol=0)       # cos() ulp at the rows only


# ---------------------------------------------------------------- Contour2D API, reference call order
def _oracle_keff_chain(q3, lat, lon, dA, grdS, N, increase, lt):
    ctr = O.cal_contours(q3, N, increase)
    tbl, c = O.cal_area_eqCoord_table_hist(lat, np.ones_like(q3[0]), dA, 0, increase, lt)
    per = q3.shape[0] > 1
    cc = ctr if per else ctr[0]
    area = O.cal_integral_within_contours_hist(q3, cc, dA, lt)
    intg = O.cal_integral_within_contours_hist(q3, cc, dA, lt, integrand=grdS)
    latEq = O.table_lookup_coordinates(area, tbl, c)
    with np.errstate(all="ignore"):
        Lmin = O.latitude_lengths_at(latEq)
        dintSdA = O.cal_gradient_wrt_area(intg, area)
        dqdA = O.cal_gradient_wrt_area(ctr, area)
        Leq2 = O.cal_sqared_equivalent_length(dintSdA, dqdA)
        nkeff = O.cal_normalized_Keff(Leq2, Lmin)
    return dict(ctr=ctr, table=tbl, area=area, intgrdS=intg, latEq=latEq, Lmin=Lmin,
                dintSdA=dintSdA, dqdA=dqdA, Leq2=Leq2, nkeff=nkeff)


def _close(a, b, rtol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.array_equal(np.isinf(a), np.isinf(b))
    f = np.isfinite(b)
    assert np.allclose(a[f], b[f], rtol=rtol, atol=0), np.abs(a[f] - b[f]).max()


@pytest.mark.parametrize("increase,lt", [(True, True), (True, False), (False, True), (False, False)])
def test_contour2d_keff_lwa_workflow(ops, vort, increase, lt):
    """tests/test_Keff_atmos.py:76-92 and tests/test_LWA.py:57-77 call order on
    Data/barotropic_vorticity.nc through the drop-in classes."""
    import xcontour_b200 as xb
    lat, lon, q = vort
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    grd = O.squared_gradient_latlon(q, lat, lon).astype(np.float32)
    coords = {"latitude": lat, "longitude": lon}
    tracer = xb.DataArray(q, dims=("latitude", "longitude"), coords=coords, name="absolute_vorticity")
    grdS = xb.DataArray(grd, dims=("latitude", "longitude"), coords=coords, name="grdS")
    dAx = xb.DataArray(dA, dims=("latitude", "longitude"), coords=coords)
    mask = xb.DataArray(np.ones_like(q), dims=("latitude", "longitude"), coords=coords)
    N = 121
    an = xb.Contour2D(tracer, dAx, dims={"X": "longitude", "Y": "latitude"}, dimEq={"Y": "latitude"},
                      increase=increase, lt=lt)
    ctr = an.cal_contours(N)
    table = an.cal_area_eqCoord_table_hist(mask)
    area = an.cal_integral_within_contours_hist(ctr).rename("intArea")
    intgrdS = an.cal_integral_within_contours_hist(ctr, integrand=grdS).rename("intgrdS")
    latEq = table.lookup_coordinates(area).rename("latEq")
    Lmin = xb.latitude_lengths_at(latEq).rename("Lmin")
    dintSdA = an.cal_gradient_wrt_area(intgrdS, area).rename("dintSdA")
    dqdA = an.cal_gradient_wrt_area(ctr, area)
    assert dqdA.name == "dabsolute_vorticitydA"                                   # core.py:488
    dqdA = dqdA.rename("dqdA")
    Leq2 = an.cal_sqared_equivalent_length(dintSdA, dqdA)
    nkeff = an.cal_normalized_Keff(Leq2, Lmin)
    ref = _oracle_keff_chain(q[None], lat, lon, dA, grd[None], N, increase, lt)

    assert ctr.dims == ("contour",) and ctr.name == "absolute_vorticity" and ctr.dtype == np.float32
    assert np.array_equal(ctr["contour"].values, np.arange(N, dtype=np.float32))
    assert np.array_equal(ctr.values, ref["ctr"][0])                              # bit-exact levels
    assert relmax(table._table.values, ref["table"]) <= RTOL_INT
    assert relmax(area.values, ref["area"][0]) <= RTOL_INT
    assert relmax(intgrdS.values, ref["intgrdS"][0]) <= RTOL_INT
    _close(latEq.values, ref["latEq"][0], 1e-11)
    _close(Lmin.values, ref["Lmin"][0], 1e-9)
    _close(dqdA.values, ref["dqdA"][0], 1e-10)
    _close(dintSdA.values, ref["dintSdA"][0], 1e-10)
    _close(Leq2.values, ref["Leq2"][0], 1e-9)
    assert nkeff.name == "nkeff" and Leq2.name == "Leq2"
    _close(nkeff.values, ref["nkeff"][0], 1e-8)

    ds_contour = xb.merge([ctr, area, latEq])
    preLats = tracer["latitude"].astype(np.float32)          # tests/test_LWA.py:72
    ds_latEq = an.interp_to_dataset(preLats, latEq, ds_contour)
    Q = ds_latEq.absolute_vorticity
    assert Q.dims == ("latitude",)
    refQ = O.interp_to_coords(preLats.values, ref["latEq"], ref["ctr"])
    _close(Q.values, refQ[0], 1e-11)
    lwa, ctrs, masks = an.cal_local_wave_activity(tracer, Q, mask_idx=[37, 125, 170, 213], part="all")
    refL, rc, rm = O.cal_local_wave_activity(q[None], Q.values[None], dA, lat, increase, "all",
                                             mask_idx=[37, 125, 170, 213])
    assert lwa.name == "LWA" and lwa.dims == tracer.dims
    assert relmax(lwa.values, refL[0]) <= RTOL_FIELD * 1e-2
    for m, r in zip(masks, rm):
        assert np.array_equal(m.values, r[0])
    lape = an.cal_local_APE(tracer, Q)
    assert lape.name == "LAPE" and np.array_equal(lape.values, lwa.values)
    with pytest.raises(Exception, match="invalid part"):
        an.cal_local_wave_activity(tracer, Q, part="middle")
    with pytest.raises(Exception, match="out of boundary"):
        an.cal_local_wave_activity(tracer, Q, mask_idx=[256])


def test_contour2d_strict_path_and_tables(ops, vort):
    """cal_integral_within_contours / cal_area_eqCoord_table (core.py:73-147,
    363-409) against the oracle's literal broadcast, incl. a land mask and a
    level-varying (3-D) tracer -- the case the reference says xhistogram cannot do
    (notebooks/1.Keff_atmos.ipynb cell 4)."""
    import xcontour_b200 as xb
    lat, lon, q = vort
    q = q[::2, ::2].copy(); lat = lat[::2].copy(); lon = lon[::2].copy()
    dA = O.latlon_cell_area(lat, lon)
    q3 = np.stack([q, 0.7 * q[::-1] + 1e-5])
    coords = {"level": np.array([300, 350]), "latitude": lat, "longitude": lon}
    tracer = xb.DataArray(q3, dims=("level", "latitude", "longitude"), coords=coords, name="pv")
    dAx = xb.DataArray(dA, dims=("latitude", "longitude"))
    m = np.ones_like(q); m[40:60, 100:140] = 0
    mask = xb.DataArray(m, dims=("latitude", "longitude"), coords={"latitude": lat, "longitude": lon})
    for increase, lt in ((True, True), (True, False), (False, False)):
        an = xb.Contour2D(tracer, dAx, dims={"X": "longitude", "Y": "latitude"}, dimEq={"Y": "latitude"},
                          increase=increase, lt=lt)
        ctr = an.cal_contours(61)
        assert ctr.dims == ("level", "contour")
        area = an.cal_integral_within_contours(ctr)
        ref = O.cal_integral_within_contours(q3, ctr.values, dA, lt)
        assert relmax(area.values, ref) <= RTOL_INT
        areah = an.cal_integral_within_contours_hist(ctr)
        refh = O.cal_integral_within_contours_hist(q3, ctr.values, dA, lt)
        assert relmax(areah.values, refh) <= RTOL_INT
        t1 = an.cal_area_eqCoord_table(mask)
        r1, _ = O.cal_area_eqCoord_table(lat, m, dA, 0, increase, lt)
        assert relmax(t1._table.values, r1) <= RTOL_INT
        t2 = an.cal_area_eqCoord_table_hist(mask)
        r2, _ = O.cal_area_eqCoord_table_hist(lat, m, dA, 0, increase, lt)
        assert relmax(t2._table.values, r2) <= RTOL_INT
        y = t2.lookup_coordinates(areah)
        assert y.dims == ("level", "contour")
        _close(y.values, O.table_lookup_coordinates(refh, r2, lat), 1e-11)


# ---------------------------------------------------------------- (8) fused batch
@pytest.mark.parametrize("increase,lt", [(True, True), (False, False)])
def test_fused_batch_matches_oracle_chain(ops, increase, lt):
    from xcontour_b200.pipeline import KeffLwaPlan
    lat, lon, q = synth_c4(3, 91, 180)
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    N = 41
    plan = KeffLwaPlan(lat, lon, dA, N, increase=increase, lt=lt)
    out = plan.run(dev(ops, q))
    torch.cuda.synchronize()
    grd = O.squared_gradient_latlon(q, lat, lon)                     # fp64, never rounded to fp32
    ref = _oracle_keff_chain(q, lat, lon, dA, grd, N, increase, lt)
    assert np.array_equal(out["ctr"].cpu().numpy().astype(np.float32), ref["ctr"])
    assert relmax(out["area"].cpu().numpy(), ref["area"]) <= RTOL_INT
    assert relmax(out["intgrdS"].cpu().numpy(), ref["intgrdS"]) <= 1e-11
    _close(out["latEq"].cpu().numpy(), ref["latEq"], 1e-11)
    _close(out["dqdA"].cpu().numpy(), ref["dqdA"], 1e-10)
    _close(out["dintSdA"].cpu().numpy(), ref["dintSdA"], 1e-9)
    _close(out["nkeff"].cpu().numpy(), ref["nkeff"], 1e-8)
    Qref = O.interp_to_coords(lat.astype(np.float32), ref["latEq"], ref["ctr"])
    _close(out["Qref"].cpu().numpy(), Qref, 1e-11)
    refL = O.cal_local_wave_activity(q, out["Qref"].cpu().numpy(), dA, lat, increase)
    assert relmax(out["lwa"].cpu().numpy(), refL) <= RTOL_FIELD * 1e-2


def test_full_size_properties(ops):
    """BASELINE config 4 shape (721x1440, N=361): size-independent properties."""
    from xcontour_b200.pipeline import KeffLwaPlan
    lat, lon, q = synth_c4(2)
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    plan = KeffLwaPlan(lat, lon, dA, 361, increase=True, lt=True)
    qd = dev(ops, q)
    out = plan.run(qd)
    torch.cuda.synchronize()
    area = out["area"].cpu().numpy()
    tot = dA.astype(np.float64).sum()
    assert np.all(np.diff(area, axis=1) >= 0)                        # CDF monotone
    assert np.all(np.abs(area[:, -1] - tot) / tot < 1e-6)            # area[-1] = sum dA (core.py:133-140)
    Q = out["Qref"].cpu().numpy()
    assert np.all(np.diff(Q, axis=1) >= 0)                           # sorted profile
    lwa = out["lwa"].cpu().numpy()
    assert lwa.min() >= -1e-10 * lwa.max()                           # LWA >= 0 for part='all'
    # spot-check 5 reference rows of slice 0 against the reference's own j-loop
    rows = [0, 100, 360, 600, 720]
    ref = O.cal_local_wave_activity(q[:1], Q[:1], dA, lat, True, rows=rows)
    for j in rows:
        assert np.abs(lwa[0, j] - ref[0, j]).max() <= RTOL_FIELD * lwa.max()
    # bit-exact bins at full size on one slice
    ctr = out["ctr"].cpu().numpy().astype(np.float32)
    e, _ = O.hist_edges(ctr[0], True)
    ee, dd = ops.hist_edges(out["ctr"][:1].contiguous(), 0, True)
    _, _, idx = ops.bin_accumulate(qd[:1].reshape(1, -1), ee, dev(ops, dA.reshape(-1)), decreasing=dd, want_idx=True)
    assert np.array_equal(idx.cpu().numpy()[0], O.digitize_bins(q[0].ravel(), e))
    # determinism
    out2 = plan.run(qd, out=plan.alloc_outputs(2))
    torch.cuda.synchronize()
    # contour space is bit-reproducible (fixed reduction order, MATCH.ANY peel serves
    # the lowest lane first); the LWA tag election may in principle pick another
    # winner order, so the field is held to rounding level rather than bit equality
    assert torch.equal(out2["nkeff"].nan_to_num(), out["nkeff"].nan_to_num())
    assert torch.equal(out2["area"], out["area"]) and torch.equal(out2["Qref"], out["Qref"])
    assert float((out2["lwa"] - out["lwa"]).abs().max()) <= 1e-13 * float(out["lwa"].abs().max())


# ---------------------------------------------------------------- more shapes / options
def test_fused_batch_fp64_tracer_grds_input_odd_sizes(ops):
    """fp64 tracer, |grad q|^2 given as an input array (fp32 and fp64), nx not a
    multiple of 4, descending latitude coordinate, fp64 contour dtype, part='upper'."""
    from xcontour_b200.pipeline import KeffLwaPlan
    rng = np.random.default_rng(17)
    ny, nx, S, N = 73, 145, 3, 37
    lat = np.linspace(88.0, -88.0, ny)                       # descending
    lon = np.arange(nx) * (360.0 / nx)
    phi, lam = np.deg2rad(lat)[:, None], np.deg2rad(lon)[None, :]
    q = np.stack([np.sin(phi) + 0.25 * np.cos(phi) ** 2 * np.sin(5 * lam + s) +
                  0.02 * rng.standard_normal((ny, nx)) for s in range(S)])          # fp64
    dA = O.latlon_cell_area(lat, lon)                                              # fp64
    for gdt in (np.float32, np.float64):
        grd = np.abs(rng.standard_normal(q.shape)).astype(gdt) * 1e-10
        for increase, lt, part in ((False, True, "all"), (False, False, "upper")):
            plan = KeffLwaPlan(lat, lon, dA, N, increase=increase, lt=lt, dtype=np.float64, part=part)
            out = plan.run(dev(ops, q), grdS=dev(ops, grd))
            torch.cuda.synchronize()
            ctr = O.cal_contours(q, N, increase, np.float64)
            assert np.array_equal(out["ctr"].cpu().numpy(), ctr)
            tbl, c = O.cal_area_eqCoord_table_hist(lat, np.ones((ny, nx)), dA, 0, increase, lt)
            area = O.cal_integral_within_contours_hist(q, ctr, dA, lt)
            intg = O.cal_integral_within_contours_hist(q, ctr, dA, lt, integrand=grd)
            assert relmax(out["area"].cpu().numpy(), area) <= RTOL_INT
            assert relmax(out["intgrdS"].cpu().numpy(), intg) <= RTOL_INT
            latEq = O.table_lookup_coordinates(area, tbl, c)
            _close(out["latEq"].cpu().numpy(), latEq, 1e-11)
            Qref = O.interp_to_coords(lat, latEq, ctr)
            _close(out["Qref"].cpu().numpy(), Qref, 1e-11)
            refL = O.cal_local_wave_activity(q, out["Qref"].cpu().numpy(), dA, lat, increase, part)
            assert relmax(out["lwa"].cpu().numpy(), refL) <= RTOL_FIELD * 1e-2


def test_contour2d_plane_orientations_and_leading_dims(ops):
    """Equivalent dimension stored last (lon, lat) and a (time, level) stack: the
    host layer transposes for the kernels and restores the caller's dim order."""
    import xcontour_b200 as xb
    lat, lon, q = synth_c4(4, 46, 90)
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    q4 = q.reshape(2, 2, 46, 90)
    qT = np.ascontiguousarray(q4.transpose(0, 1, 3, 2))                # (time, level, lon, lat)
    coords = {"time": np.arange(2), "level": np.array([300, 350]), "lat": lat, "lon": lon}
    tr = xb.DataArray(qT, dims=("time", "level", "lon", "lat"), coords=coords, name="pv")
    dAx = xb.DataArray(dA, dims=("lat", "lon"), coords={"lat": lat, "lon": lon})
    mask = xb.DataArray(np.ones((46, 90), np.float32), dims=("lat", "lon"), coords={"lat": lat, "lon": lon})
    an = xb.Contour2D(tr, dAx, dims={"X": "lon", "Y": "lat"}, dimEq={"Y": "lat"}, increase=True, lt=True)
    ctr = an.cal_contours(31)
    assert ctr.dims == ("time", "level", "contour")
    ref_ctr = O.cal_contours(q, 31, True).reshape(2, 2, 31)
    assert np.array_equal(ctr.values, ref_ctr)
    area = an.cal_integral_within_contours_hist(ctr)
    ref_area = O.cal_integral_within_contours_hist(q, ref_ctr.reshape(4, 31), dA, True).reshape(2, 2, 31)
    assert relmax(area.values, ref_area) <= RTOL_INT
    table = an.cal_area_eqCoord_table_hist(mask)
    latEq = table.lookup_coordinates(area)
    Q = an.interp_to_coords(xb.DataArray(lat, dims=("lat",), coords={"lat": lat}), latEq, ctr)
    assert Q.dims == ("time", "level", "lat")
    lwa = an.cal_local_wave_activity(tr, Q)
    assert lwa.dims == tr.dims
    refL = O.cal_local_wave_activity(q, Q.values.reshape(4, 46), dA, lat, True).reshape(2, 2, 46, 90)
    assert relmax(lwa.values.transpose(0, 1, 3, 2), refL) <= RTOL_FIELD * 1e-2
    lwa2 = an.cal_local_wave_activity2(tr, Q)
    refL2 = O.cal_local_wave_activity(q, Q.values.reshape(4, 46), dA, lat, True, variant=2).reshape(2, 2, 46, 90)
    assert relmax(lwa2.values.transpose(0, 1, 3, 2), refL2) <= RTOL_FIELD * 1e-2


def test_contours_at_and_contour_means(ops, vort):
    """cal_contours_at_hist (core.py:316-360) and the along-contour means
    (core.py:491-616) are compositions of the kernels above."""
    import xcontour_b200 as xb
    lat, lon, q = vort
    q = q[::2, ::2].copy(); lat = lat[::2].copy(); lon = lon[::2].copy()
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    rng = np.random.default_rng(3)
    g = np.abs(rng.standard_normal(q.shape)).astype(np.float32)
    f = rng.standard_normal(q.shape).astype(np.float32)
    coords = {"latitude": lat, "longitude": lon}
    tr = xb.DataArray(q, dims=("latitude", "longitude"), coords=coords, name="vor")
    gx = xb.DataArray(g, dims=("latitude", "longitude"), coords=coords, name="grd")
    fx = xb.DataArray(f, dims=("latitude", "longitude"), coords=coords, name="f")
    dAx = xb.DataArray(dA, dims=("latitude", "longitude"), coords=coords)
    mask = xb.DataArray(np.ones_like(q), dims=("latitude", "longitude"), coords=coords)
    an = xb.Contour2D(tr, dAx, dims={"X": "longitude", "Y": "latitude"}, dimEq={"Y": "latitude"},
                      increase=True, lt=True)
    table = an.cal_area_eqCoord_table_hist(mask)
    pre = np.linspace(-80, 80, 33).astype(np.float32)
    qat = an.cal_contours_at_hist(pre, table)
    ctr = O.cal_contours(q[None], 33, True)
    area = O.cal_integral_within_contours_hist(q[None], ctr[0], dA, True)
    tbl, c = O.cal_area_eqCoord_table_hist(lat, np.ones_like(q), dA, 0, True, True)
    ref = O.interp_to_coords(pre, O.table_lookup_coordinates(area, tbl, c), ctr)[0]
    assert qat.dims == ("contour",) and qat.name == "vor"
    _close(qat.values, ref, 1e-11)
    c61 = an.cal_contours(61)
    cm = an.cal_contour_mean_hist(c61, fx, gx)
    ref61 = O.cal_contours(q[None], 61, True)
    a61 = O.cal_integral_within_contours_hist(q[None], ref61[0], dA, True)
    with np.errstate(all="ignore"):
        up = O.cal_gradient_wrt_area(O.cal_integral_within_contours_hist(q[None], ref61[0], dA, True, integrand=(f * g)[None]), a61)
        lo = O.cal_gradient_wrt_area(O.cal_integral_within_contours_hist(q[None], ref61[0], dA, True, integrand=g[None]), a61)
        refcm = (up / lo)[0]
    assert cm.name == "cmf"
    _close(cm.values, refcm, 1e-8)


def test_error_conventions_and_degenerate_inputs(ops):
    """Bare Exceptions with the reference's messages; degenerate slices."""
    import xcontour_b200 as xb
    lat = np.linspace(-60, 60, 13).astype(np.float32); lon = np.arange(24, dtype=np.float32) * 15
    q = np.zeros((13, 24), np.float32) + 2.5                       # constant tracer
    tr = xb.DataArray(q, dims=("lat", "lon"), coords={"lat": lat, "lon": lon}, name="c")
    dAx = xb.DataArray(np.ones((13, 24), np.float32), dims=("lat", "lon"))
    an = xb.Contour2D(tr, dAx, dims={"X": "lon", "Y": "lat"}, dimEq={"Y": "lat"}, increase=True, lt=True)
    ctr = an.cal_contours(5)
    assert np.array_equal(ctr.values, np.full(5, 2.5, np.float32))
    with pytest.raises(Exception, match="non monotonic bins"):       # core.py:1233-1240
        an.cal_integral_within_contours_hist(ctr)
    with pytest.raises(Exception, match="predef should be a 1D array"):
        an.cal_contours_at_hist(np.zeros((2, 2)), None)
    # strict path on the constant tracer: nothing is < 2.5, everything is < 3
    lv = np.array([2.0, 2.5, 3.0], np.float32)
    a = an.cal_integral_within_contours(lv)
    assert a.values.tolist() == [0.0, 0.0, 13 * 24.0]
    # all-NaN slice: levels are NaN, like nanmin/nanmax of an empty set
    qn = np.full((2, 4, 8), np.nan, np.float32); qn[1] = np.arange(32, dtype=np.float32).reshape(4, 8)
    lv, mm = ops.minmax_levels(dev(ops, qn.reshape(2, -1)), 4, True, 0)
    lv = lv.cpu().numpy()
    assert np.isnan(lv[0]).all() and np.array_equal(lv[1].astype(np.float32), O.cal_contours(qn[1:], 4, True)[0])
    # argument errors come back through xc_last_error
    with pytest.raises(Exception, match="N>=2"):
        ops.minmax_levels(dev(ops, qn.reshape(2, -1)), 1, True, 0)
    t = xb.Table(xb.DataArray(np.array([0.0, 1.0, 3.0]), dims=("lat",), coords={"lat": np.array([-1.0, 0.0, 1.0])}), "lat")
    assert np.allclose(t.lookup_coordinates(np.array([0.5, 2.0, 9.0])), [-0.5, 0.5, 1.0])
    assert np.allclose(t.lookup_values(np.array([-0.5, 0.5])), [0.5, 2.0])
    with pytest.raises(Exception, match="not every time or level"):
        xb.Table(xb.DataArray(np.array([[0.0, 1.0], [1.0, 0.0]]), dims=("t", "lat"),
                              coords={"lat": np.array([0.0, 1.0])}), "lat")


@pytest.mark.parametrize("increase,lt", [(True, True), (False, False)])
def test_equal_area_levels_weighted_quantile_histogram(ops, vort, increase, lt):
    """north_star kernel (1): equal-area levels from a weighted-quantile histogram.
    Parity with the oracle's restatement, and the levels sit within one fine bin of
    the exact weighted quantiles obtained by sorting."""
    import xcontour_b200 as xb
    lat, lon, q = vort
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    q3 = np.stack([q, 0.5 * q[::-1]])
    tr = xb.DataArray(q3, dims=("time", "lat", "lon"), coords={"lat": lat, "lon": lon}, name="vor")
    an = xb.Contour2D(tr, xb.DataArray(dA, dims=("lat", "lon")), dims={"X": "lon", "Y": "lat"},
                      dimEq={"Y": "lat"}, increase=increase, lt=lt)
    N, refine = 41, 8
    lev = an.cal_contours_equal_area(N, refine=refine)
    assert lev.dims == ("time", "contour") and lev.dtype == np.float32
    ref = O.cal_contours_equal_area(q3, dA, N, increase, lt, np.float32, refine)
    assert np.allclose(lev.values, ref, rtol=0, atol=2e-12)           # fp32 levels: identical up to 1 ulp
    assert np.abs(lev.values.view(np.int32) - ref.view(np.int32)).max() <= 1
    # enclosed areas of the returned levels are equally spaced to within one fine bin
    area = an.cal_integral_within_contours_hist(lev).values
    d = np.diff(area, axis=1)
    fine_bin = np.abs(area[:, -1] - area[:, 0])[:, None] / ((N - 1) * refine)
    assert np.all(np.abs(d - d.mean(axis=1, keepdims=True)) <= 4 * fine_bin)
    # and they bracket the exact weighted quantiles (sorting) within one fine level step
    for s in range(2):
        step = (q3[s].max() - q3[s].min()) / ((N - 1) * refine)
        fr = np.linspace(0, 1, N)
        fr = fr if increase else 1 - fr          # decreasing levels enclose the complementary fraction
        exact = O.weighted_quantile_levels(q3[s], dA, fr)
        assert np.abs(lev.values[s][1:-1] - exact[1:-1]).max() <= 2 * step
    # a single slice takes the static-bins branch of _histogram (edges in the contour dtype); straight through the C ABI
    one = ops.equal_area_levels(dev(ops, q3[:1].reshape(1, -1)), dev(ops, dA.reshape(-1)), N, refine, increase, lt,
                                0, False).cpu().numpy().astype(np.float32)
    ref1 = O.cal_contours_equal_area(q3[:1], dA, N, increase, lt, np.float32, refine)
    assert np.abs(one.view(np.int32) - ref1.view(np.int32)).max() <= 1
    # fp64 levels, NumPy >= 2 edge rules
    two = ops.equal_area_levels(dev(ops, q3.reshape(2, -1)), dev(ops, dA.reshape(-1)), N, refine, increase, lt,
                                1, True).cpu().numpy()
    fine = O.cal_contours(q3, (N - 1) * refine + 1, increase, np.float64)
    area = O.cal_integral_within_contours_hist(q3, fine, dA, lt, scalar_rules="numpy2")
    tgt = area[:, :1] + (area[:, -1:] - area[:, :1]) * np.linspace(0.0, 1.0, N)[None, :]
    ref2 = np.stack([O.interp1d(tgt[s], area[s], fine[s], bool(area[0, 0] < area[0, -1])) for s in range(2)])
    assert np.allclose(two, ref2, rtol=1e-12, atol=0)


def test_device_cell_area(ops):
    """f4: lat-lon cell areas built on the device equal the host helper."""
    from xcontour_b200.utils import latlon_cell_area
    for lat in (np.linspace(-90, 90, 73), np.linspace(88, -88, 45), np.array([-60.0, -20.0, 10.0, 35.0, 80.0])):
        nx, dlon = 48, 7.5
        ref = latlon_cell_area(lat, np.arange(nx) * dlon)
        out = ops.latlon_cell_area(dev(ops, lat.astype(np.float64)), nx, dlon).cpu().numpy()
        assert np.allclose(out, ref, rtol=1e-13, atol=0)
        out32 = ops.latlon_cell_area(dev(ops, lat.astype(np.float64)), nx, dlon, torch.float32).cpu().numpy()
        assert np.allclose(out32, ref.astype(np.float32), rtol=2e-7, atol=0)


def test_contour2d_1d_area_explicit_levels_check_mono(ops, vort):
    """dA given on the latitude axis only, explicit level arrays (core.py:251-264),
    decreasing explicit levels, and the check_mono flag (core.py:1328-1355)."""
    import xcontour_b200 as xb
    lat, lon, q = vort
    q = q[::4, ::4].copy(); lat = lat[::4].copy(); lon = lon[::4].copy()
    dA2 = O.latlon_cell_area(lat, lon)
    coords = {"lat": lat, "lon": lon}
    tr = xb.DataArray(q, dims=("lat", "lon"), coords=coords, name="vor")
    dA1 = xb.DataArray(dA2[:, 0].copy(), dims=("lat",), coords={"lat": lat})      # broadcast along lon
    an = xb.Contour2D(tr, dA1, dims={"X": "lon", "Y": "lat"}, dimEq={"Y": "lat"}, increase=True, lt=True)
    levs = np.linspace(q.min(), q.max(), 17).astype(np.float32)
    ctr = an.cal_contours(levs)
    assert ctr.dims == ("contour",) and np.array_equal(ctr.values, levs)
    assert np.array_equal(ctr["contour"].values, levs)
    a = an.cal_integral_within_contours_hist(ctr)
    ref = O.cal_integral_within_contours_hist(q[None], levs, dA2, True, time_branch=False)[0]
    assert relmax(a.values, ref) <= RTOL_INT
    # decreasing explicit levels: same numbers in reversed order (core.py:454-455)
    a_dec = an.cal_integral_within_contours_hist(levs[::-1].copy())
    ref_dec = O.cal_integral_within_contours_hist(q[None], levs[::-1].copy(), dA2, True, time_branch=False)[0]
    assert relmax(a_dec.values, ref_dec) <= RTOL_INT
    # non-uniform explicit levels exercise the binary-search path of the binning kernel
    nl = np.sort(np.concatenate([levs[:3], levs[3] + (levs[-1] - levs[3]) * np.linspace(0, 1, 9) ** 3])).astype(np.float32)
    nl = np.unique(nl)
    a_nu = an.cal_integral_within_contours(nl)
    assert relmax(a_nu.values, O.cal_integral_within_contours(q[None], nl, dA2, True)[0]) <= RTOL_INT
    # check_mono: many levels on a small field leave empty bins -> zero differences
    an2 = xb.Contour2D(tr, dA1, dims={"X": "lon", "Y": "lat"}, dimEq={"Y": "lat"}, increase=True, lt=True,
                       check_mono=True)
    with pytest.raises(Exception, match="not monotonic var"):
        an2.cal_integral_within_contours_hist(an2.cal_contours(4001))


@pytest.mark.parametrize("env", [
    {"XCB200_LWA_FX": "0"},                                      # fp64 read-modify-write LWA kernel (k_lwa_fast)
    {"XCB200_NO_LWA_COLS": "1"},                                 # fixed-point LWA kernel for general weights (k_lwa_fx)
    {"XCB200_NO_BIN_ROWS": "1"},                                 # general fp64 binning kernel (k_hist) in the fused Keff pass
    {"XCB200_NO_BIN_ROWS": "1", "XCB200_OVERLAP": "0"},          # ... on the serial schedule
    {"XCB200_NO_BULK": "1"},                                     # register-staged min/max instead of the bulk-copy ring
    {"XCB200_SUB_BATCH": "1"},                                   # one slice per pass, two passes in flight
])
def test_alternate_code_paths_smoke(ops, env):
    """The run-time selectable variants (read once per process from the environment)
    each pass the oracle-checked smoke run in a fresh process."""
    import subprocess, sys
    from conftest import ROOT
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "__graft_entry__.py"), "--smoke-only"],
                       env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "smoke ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
