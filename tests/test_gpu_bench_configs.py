"""
GPU parity at the sizes bench.py is quoted on (-m gpu): BASELINE.json config 4
(721x1440, 361 contours, Keff + LWA) and config 5 (4096x8192, 2048 contours, Keff
part) against the NumPy oracle on the same seeded inputs, plus the callers around
the fused batch (HostStreamer, torch / DLPack inputs of the drop-in class).

Reference path being checked: xcontour/core.py:205-266 (levels), :412-460 +
:1202-1325 (histogram / CDF), :463-488 (d/dA), :619-637 + :945-966 (Keff),
:1050-1100 (Q), :696-799 (LWA).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from conftest import synth_c4
from oracle import xcontour_oracle as O
from test_gpu_parity import _close, _oracle_keff_chain, relmax

pytestmark = pytest.mark.gpu

RTOL_INT = 1e-12      # integrals, relative to the array maximum (bar of north_star: 1e-10)
RTOL_FIELD = 1e-12    # LWA, relative to the field maximum (bar: 1e-10)


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xcontour_b200 import ops as _ops
    _ops.require_cuda()
    return _ops


def dev(ops, a):
    return ops.to_dev(np.ascontiguousarray(a))


# ---------------------------------------------------------------- config 4
@pytest.fixture(scope="module")
def c4(ops):
    """Two full-size slices through the fused batch (the call bench.py times)."""
    from xcontour_b200.pipeline import KeffLwaPlan
    lat, lon, q = synth_c4(2)
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    plan = KeffLwaPlan(lat, lon, dA, 361, increase=True, lt=True)
    qd = dev(ops, q)
    out = plan.run(qd)
    torch.cuda.synchronize()
    return lat, lon, q, dA, plan, qd, out


def test_c4_contour_space_matches_oracle_chain(ops, c4):
    """Every contour-space result of both slices at 721x1440 / N = 361 against the oracle chain."""
    lat, lon, q, dA, plan, qd, out = c4
    grd = O.squared_gradient_latlon(q, lat, lon)
    ref = _oracle_keff_chain(q, lat, lon, dA, grd, 361, True, True)
    got = {k: v.cpu().numpy() for k, v in out.items() if k != "lwa"}
    assert np.array_equal(got["ctr"].astype(np.float32), ref["ctr"])          # levels bit-exact
    assert relmax(got["area"], ref["area"]) <= RTOL_INT
    # |grad q|^2 dA: the polar rows put single terms 2^100 above the mid-latitude ones, so the CDF is held
    # (i) to its own maximum and (ii) bin by bin (pdf = diff of the CDF) to each bin's own value plus the
    # cancellation noise of that difference (a few ulp of the CDF reached so far)
    assert relmax(got["intgrdS"], ref["intgrdS"]) <= RTOL_INT
    pdf_got = np.diff(got["intgrdS"], axis=1, prepend=0.0)
    pdf_ref = np.diff(ref["intgrdS"], axis=1, prepend=0.0)
    cum = np.maximum.accumulate(np.abs(ref["intgrdS"]), axis=1)
    assert np.all(np.abs(pdf_got - pdf_ref) <= 1e-10 * np.abs(pdf_ref) + 8e-16 * cum)
    _close(got["latEq"], ref["latEq"], 1e-11)
    _close(got["Lmin"], ref["Lmin"], 1e-9)
    _close(got["dqdA"], ref["dqdA"], 1e-10)
    _close(got["dintSdA"], ref["dintSdA"], 1e-9)
    _close(got["Leq2"], ref["Leq2"], 1e-8)
    _close(got["nkeff"], ref["nkeff"], 1e-8)
    Qref = O.interp_to_coords(lat.astype(np.float32), ref["latEq"], ref["ctr"])
    _close(got["Qref"], Qref, 1e-11)


def test_c4_bins_bit_exact_both_slices(ops, c4):
    lat, lon, q, dA, plan, qd, out = c4
    ctr = out["ctr"].cpu().numpy().astype(np.float32)
    ee, dd = ops.hist_edges(out["ctr"].contiguous(), 0, True)
    _, _, idx = ops.bin_accumulate(qd.reshape(2, -1), ee, dev(ops, dA.reshape(-1)), decreasing=dd, want_idx=True)
    idx = idx.cpu().numpy()
    for s in range(2):
        e, _ = O.hist_edges(ctr[s], True)
        assert np.array_equal(idx[s], O.digitize_bins(q[s].ravel(), e))


def test_c4_lwa_rows_match_reference_loop(ops, c4):
    """40 rows of each slice (both poles, their neighbours, an even spread) against the reference's own
    j-loop (core.py:752-794) at full size."""
    lat, lon, q, dA, plan, qd, out = c4
    rows = sorted(set([0, 1, 2, 359, 360, 361, 718, 719, 720] + list(range(5, 721, 23))))
    assert len(rows) >= 32
    Q = out["Qref"].cpu().numpy()
    lwa = out["lwa"].cpu().numpy()
    ref = O.cal_local_wave_activity(q, Q, dA, lat, True, rows=rows)
    for s in range(2):
        for j in rows:
            assert np.abs(lwa[s, j] - ref[s, j]).max() <= RTOL_FIELD * lwa[s].max(), (s, j)
    assert lwa.min() >= -1e-12 * lwa.max()


def test_c4_run_to_run_bit_identical(ops, c4):
    """Integer accumulators: every output, LWA included, is bit-reproducible whatever the schedule."""
    lat, lon, q, dA, plan, qd, out = c4
    for _ in range(2):
        out2 = plan.run(qd, out=plan.alloc_outputs(2))
        torch.cuda.synchronize()
        for k in out:
            assert torch.equal(out2[k].nan_to_num(), out[k].nan_to_num()), k


def test_c4_batch_split_does_not_change_results(ops, c4):
    """sub_batch (slices per pass) and the position of a slice in the batch are invisible in the results."""
    from xcontour_b200.pipeline import KeffLwaPlan
    lat, lon, q, dA, plan, qd, out = c4
    plan1 = KeffLwaPlan(lat, lon, dA, 361, increase=True, lt=True, sub_batch=1)
    out1 = plan1.run(qd.flip(0).contiguous())
    torch.cuda.synchronize()
    for k in out:
        assert torch.equal(out1[k].flip(0).nan_to_num(), out[k].nan_to_num()), k


def test_host_streamer_equals_plan_run(ops):
    """pipeline.HostStreamer (pinned host -> H2D -> fused batch -> D2H, buffers in flight): every host array it
    hands to `consume` equals plan.run on the same slices, including a ragged last batch."""
    from xcontour_b200.pipeline import HostStreamer, KeffLwaPlan
    lat, lon, q = synth_c4(7, 181, 360)
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    plan = KeffLwaPlan(lat, lon, dA, 91, increase=True, lt=True)
    ref = {k: v.cpu() for k, v in plan.run(dev(ops, q)).items()}
    torch.cuda.synchronize()
    qh = torch.from_numpy(q).pin_memory()
    for batch, nbuf in ((3, 2), (2, 3), (7, 2)):
        hs = HostStreamer(plan, batch, nbuf=nbuf)
        seen = []

        def consume(s0, s1, host):
            seen.append((s0, s1))
            for k, v in host.items():
                assert torch.equal(v.nan_to_num(), ref[k][s0:s1].nan_to_num()), (k, s0, s1)
        hs.run(qh, consume=consume)
        assert sorted(seen) == [(b, min(7, b + batch)) for b in range(0, 7, batch)]
        assert hs.h2d_bytes == q.nbytes
        assert hs.d2h_bytes == sum(v.numel() * v.element_size() for v in ref.values())


class _DLPackOnly(object):
    """An array that can ONLY be consumed through the DLPack protocol (no __array__, no torch type)."""

    def __init__(self, t):
        self._t = t
        self.shape, self.ndim = tuple(t.shape), t.dim()

    def __dlpack__(self, *a, **k):
        return self._t.__dlpack__(*a, **k)

    def __dlpack_device__(self):
        return self._t.__dlpack_device__()


def test_contour2d_accepts_torch_and_dlpack_tracers(ops, vort):
    """north_star: 'exchanges buffers with torch via DLPack'.  The drop-in class takes a numpy array, a CUDA /
    CPU torch.Tensor or any DLPack exporter as the tracer's values and returns the same numbers."""
    import xcontour_b200 as xb
    lat, lon, q = vort
    q = q[::4, ::4].copy(); lat = lat[::4].copy(); lon = lon[::4].copy()
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    coords = {"latitude": lat, "longitude": lon}
    dAx = xb.DataArray(dA, dims=("latitude", "longitude"), coords=coords)

    def chain(values):
        tr = xb.DataArray(values, dims=("latitude", "longitude"), coords=coords, name="pv")
        an = xb.Contour2D(tr, dAx, dims={"X": "longitude", "Y": "latitude"}, dimEq={"Y": "latitude"},
                          increase=True, lt=True)
        ctr = an.cal_contours(41)
        area = an.cal_integral_within_contours_hist(ctr)
        table = an.cal_area_eqCoord_table_hist(xb.DataArray(np.ones_like(dA), dims=("latitude", "longitude"), coords=coords))
        latEq = table.lookup_coordinates(area)
        Q = an.interp_to_coords(xb.DataArray(lat, dims=("latitude",), coords={"latitude": lat}), latEq, ctr)
        lwa = an.cal_local_wave_activity(tr, Q)
        return [np.asarray(x.values) for x in (ctr, area, latEq, Q, lwa)]

    base = chain(q)
    assert relmax(base[1], O.cal_integral_within_contours_hist(q[None], base[0], dA, True)[0]) <= RTOL_INT
    for values in (torch.from_numpy(q), torch.from_numpy(q).cuda(), _DLPackOnly(torch.from_numpy(q).cuda()),
                   _DLPackOnly(torch.from_numpy(q))):
        got = chain(values)
        for a, b in zip(got, base):
            assert np.array_equal(a, b, equal_nan=True)


# ---------------------------------------------------------------- config 5
def _c5_field(S=1, ny=4096, nx=8192):
    y = (np.arange(ny) + 0.5) / ny
    x = (np.arange(nx) + 0.5) / nx
    out = np.empty((S, ny, nx), np.float32)
    for s in range(S):
        rng = np.random.default_rng(4321 + s)
        out[s] = (y[:, None] + 0.2 * np.sin(8 * np.pi * x)[None, :] * np.sin(4 * np.pi * y)[:, None]
                  + 0.01 * rng.standard_normal((ny, nx))).astype(np.float32)
    return y, x, out


def test_c5_histogram_scan_stress_matches_oracle(ops):
    """BASELINE config 5: 4096x8192 Cartesian tracer, 2048 contour levels, area + an fp32 integrand.
    Levels and bin assignment bit-exact, both CDFs <= 1e-12 of their maximum."""
    ny, nx, N = 4096, 8192, 2048
    y, x, q = _c5_field(1, ny, nx)
    rng = np.random.default_rng(99)
    dA = np.full((ny, nx), 1.0 / (ny * nx))
    g = rng.random((1, ny, nx)).astype(np.float32)
    qd, dAd, gd = dev(ops, q.reshape(1, -1)), dev(ops, dA.reshape(-1)), dev(ops, g.reshape(1, -1))
    lv, _ = ops.minmax_levels(qd, N, True, 0)
    ctr = O.cal_contours(q, N, True)
    assert np.array_equal(lv.cpu().numpy().astype(np.float32), ctr)
    e, d = ops.hist_edges(lv, 0, True)
    cdf, _, idx = ops.bin_accumulate(qd, e, dAd, acc_area=True, integrands=[gd], decreasing=d, want_idx=True)
    eh, _ = O.hist_edges(ctr[0], True)
    assert np.array_equal(idx.cpu().numpy()[0], O.digitize_bins(q[0].ravel(), eh))
    ref_a = O.cal_integral_within_contours_hist(q, ctr, dA, True)
    ref_g = O.cal_integral_within_contours_hist(q, ctr, dA, True, integrand=g)
    c = cdf.cpu().numpy()
    assert relmax(c[:, 0], ref_a) <= RTOL_INT
    assert relmax(c[:, 1], ref_g) <= RTOL_INT


# ---------------------------------------------------------------- general stencil (A9) through the row-march kernel
def _plan_chain_reference(q, y, dA, grd, N, increase, lt):
    ctr = O.cal_contours(q, N, increase)
    area = O.cal_integral_within_contours_hist(q, ctr, dA, lt)
    intg = O.cal_integral_within_contours_hist(q, ctr, dA, lt, integrand=grd)
    return ctr, area, intg


@pytest.mark.parametrize("bcx,bcy", [("periodic", "extend"), ("extend", "reflect"), ("fill", "fill"),
                                     ("reflect", "periodic")])
def test_cartesian_stencil_boundaries_in_flight(ops, bcx, bcy):
    """|grad q|^2 with every ghost-cell rule on an X-Z / Cartesian plane (non-uniform, DESCENDING row
    coordinate, row-dependent cell areas, NaN cells = topography), computed inside the binning pass,
    against the oracle's np.pad statement (callers' BCs: tests/test_Keff_ocean.py:26-32,
    tests/test_clength.py:39-45)."""
    from xcontour_b200.pipeline import KeffLwaPlan
    from xcontour_b200.utils import row_metrics_cartesian
    rng = np.random.default_rng(11)
    ny, nx, S, N = 83, 152, 3, 47
    z = -np.cumsum(1.0 + rng.random(ny))                       # descending, non-uniform
    x = np.arange(nx) * 250.0
    q = (np.linspace(20, 2, ny)[None, :, None] + 0.6 * np.sin(2 * np.pi * x / x[-1] * 3 + np.arange(S)[:, None, None])
         + 0.1 * rng.standard_normal((S, ny, nx))).astype(np.float32)
    q[0, 60:, 40:70] = np.nan                                   # topography
    q[2, 0, 0] = np.nan
    dA = np.repeat((np.abs(np.gradient(z)) * 250.0)[:, None], nx, axis=1)          # fp64, row-constant
    cx, cy = row_metrics_cartesian(z, x)
    assert np.array_equal(np.stack([cx, cy]), np.stack(O.row_metrics_cartesian(z, x)))
    grd = O.squared_gradient(q, cx, cy, bcx, bcy, fill=1.5)
    for increase, lt in ((False, True), (True, False)):
        plan = KeffLwaPlan(z, x, dA, N, increase=increase, lt=lt, metrics=(cx, cy), boundary=(bcx, bcy), fill_value=1.5)
        out = plan.run(dev(ops, q))
        torch.cuda.synchronize()
        ctr, area, intg = _plan_chain_reference(q, z, dA, np.nan_to_num(grd, nan=0.0), N, increase, lt)
        assert np.array_equal(out["ctr"].cpu().numpy().astype(np.float32), ctr)
        assert relmax(out["area"].cpu().numpy(), area) <= RTOL_INT
        assert relmax(out["intgrdS"].cpu().numpy(), intg) <= RTOL_INT
        out2 = plan.run(dev(ops, q))
        torch.cuda.synchronize()
        assert torch.equal(out2["intgrdS"], out["intgrdS"]) and torch.equal(out2["area"], out["area"])


def test_row_march_kernel_equals_general_kernel(ops, monkeypatch):
    """The row-march binning kernel (bin_rows.cu) against the general fp64 read-modify-write kernel (hist.cu) on
    the same call: bin-by-bin agreement of the pdfs to 1e-13 of each bin's own value, incl. the pole rows."""
    from xcontour_b200.pipeline import KeffLwaPlan
    lat, lon, q = synth_c4(3, 181, 360)
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    plan = KeffLwaPlan(lat, lon, dA, 91, increase=True, lt=True)
    a = {k: v.cpu().numpy() for k, v in plan.run(dev(ops, q)).items()}
    torch.cuda.synchronize()
    plan.dA_row = None                                          # no row-constancy hint: general kernels
    b = {k: v.cpu().numpy() for k, v in plan.run(dev(ops, q)).items()}
    torch.cuda.synchronize()
    for k in ("area", "intgrdS"):
        pa, pb = np.diff(a[k], axis=1, prepend=0.0), np.diff(b[k], axis=1, prepend=0.0)
        cum = np.maximum.accumulate(np.abs(b[k]), axis=1)
        assert np.all(np.abs(pa - pb) <= 1e-13 * np.abs(pb) + 4e-16 * cum), k
    assert np.array_equal(a["ctr"], b["ctr"])


def test_c5_keff_in_flight_cartesian(ops):
    """BASELINE config 5 through the fused batch (Keff part, no LWA): 4096x8192 Cartesian tracer, 2048 levels,
    |grad q|^2 in flight (periodic x, edge value in y), uniform cell area."""
    from xcontour_b200.pipeline import KeffLwaPlan
    from xcontour_b200.utils import row_metrics_cartesian
    ny, nx, N = 4096, 8192, 2048
    y, x, q = _c5_field(1, ny, nx)
    dA = np.full((ny, nx), 1.0 / (ny * nx))
    cx, cy = row_metrics_cartesian(y, x)
    plan = KeffLwaPlan(y, x, dA, N, increase=True, lt=True, metrics=(cx, cy), boundary=("periodic", "extend"))
    assert plan.uniform_dA and not plan.any_degenerate
    out = plan.alloc_outputs(1, lwa=False)
    plan.run(dev(ops, q), out=out)
    torch.cuda.synchronize()
    grd = O.squared_gradient(q, cx, cy, "periodic", "extend")
    ctr = O.cal_contours(q, N, True)
    assert np.array_equal(out["ctr"].cpu().numpy().astype(np.float32), ctr)
    area = O.cal_integral_within_contours_hist(q, ctr, dA, True)
    intg = O.cal_integral_within_contours_hist(q, ctr, dA, True, integrand=grd)
    assert relmax(out["area"].cpu().numpy(), area) <= RTOL_INT
    assert relmax(out["intgrdS"].cpu().numpy(), intg) <= RTOL_INT
    pa, pb = np.diff(out["intgrdS"].cpu().numpy(), axis=1, prepend=0.0), np.diff(intg, axis=1, prepend=0.0)
    assert np.all(np.abs(pa - pb) <= 1e-11 * np.abs(pb) + 1e-15 * intg.max())


# ---------------------------------------------------------------- d/dA against explicit (non-uniform) levels
@pytest.mark.parametrize("ldt", [np.float32, np.float64])
def test_gradient_wrt_area_follows_the_contour_coordinate(ops, vort, ldt):
    """cal_contours(array) labels the contour axis with the level VALUES (core.py:264) and
    cal_gradient_wrt_area differentiates against that coordinate (core.py:480-483, np.gradient's non-uniform
    interior formula).  The reference's own explicit-levels branch raises inside apply_ufunc (core.py:256-262
    returns an (N, 1) array), so this is held to the oracle's statement of the evident intent."""
    import xcontour_b200 as xb
    lat, lon, q = vort
    q = q[::4, ::4].copy(); lat = lat[::4].copy(); lon = lon[::4].copy()
    dA = O.latlon_cell_area(lat, lon)                           # fp64: the strict path then sums in fp64 in the reference too
    coords = {"latitude": lat, "longitude": lon}
    tr = xb.DataArray(q, dims=("latitude", "longitude"), coords=coords, name="pv")
    dAx = xb.DataArray(dA, dims=("latitude", "longitude"), coords=coords)
    lo, hi = float(q.min()), float(q.max())
    rng = np.random.default_rng(2)
    g = np.abs(rng.standard_normal(q.shape))
    gx = xb.DataArray(g, dims=("latitude", "longitude"), coords=coords, name="g")
    for levels in ((lo + (hi - lo) * np.linspace(0.02, 0.98, 23) ** 1.7).astype(ldt),          # non-uniform
                   np.linspace(lo, hi, 23).astype(ldt)[2:-2] if ldt == np.float64 else
                   (np.arange(19) * np.float32(2.0 ** -17) + np.float32(-2.0 ** -14)).astype(ldt)):  # uniform, spacing != 1
        an = xb.Contour2D(tr, dAx, dims={"X": "longitude", "Y": "latitude"}, dimEq={"Y": "latitude"},
                          increase=True, lt=True, dtype=ldt)
        ctr = an.cal_contours(levels)
        assert np.array_equal(ctr["contour"].values, levels)
        area = an.cal_integral_within_contours(ctr)
        intg = an.cal_integral_within_contours(ctr, integrand=gx)
        assert np.array_equal(area["contour"].values, levels)
        dq = an.cal_gradient_wrt_area(ctr, area)
        dint = an.cal_gradient_wrt_area(intg, area)
        ref_area = O.cal_integral_within_contours(q[None], levels[None], dA, True)[0]
        assert relmax(area.values, ref_area) <= RTOL_INT
        with np.errstate(all="ignore"):
            ref_dq = O.cal_gradient_wrt_area(ctr.values, area.values, var_coord=levels, area_coord=levels)
            ref_dint = O.cal_gradient_wrt_area(intg.values, area.values, var_coord=levels, area_coord=levels)
        assert dq.dtype == ref_dq.dtype and dint.dtype == ref_dint.dtype
        assert np.array_equal(dq.values, ref_dq, equal_nan=True)                 # bit-exact, like the unit-spacing case
        assert np.array_equal(dint.values, ref_dint, equal_nan=True)
        # ... which is not what unit spacing gives (the round-1 behaviour) when the levels are not uniform -- both
        # estimate dq/dA, so they agree to second order, not bit for bit
        if not np.allclose(np.diff(levels), np.diff(levels)[0]):
            unit = O.cal_gradient_wrt_area(ctr.values, area.values)
            assert not np.array_equal(dq.values[1:-1], unit[1:-1])


# ---------------------------------------------------------------- the one collective of the path
def test_contour_gather_nccl_single_rank_roundtrip(ops):
    """ContourGather (packed [9, S, N] all-gather on a side stream, NCCL) on a one-rank group: what lands in the
    receive buffer is bit-for-bit what plan.run produced, for two batches in flight; the world-size-2 logic of the
    dict-based gather is covered on gloo in tests/test_oracle.py.  Runs in its own process under a hard timeout
    (a collective that cannot initialise must fail this test, not hang the suite)."""
    import subprocess, sys
    from conftest import ROOT
    code = (
        "import os, sys, numpy as np, torch; sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'tests'))\n"
        "import torch.distributed as dist\n"
        "from conftest import synth_c4\n"
        "from xcontour_b200 import ops\n"
        "from xcontour_b200.utils import latlon_cell_area\n"
        "from xcontour_b200.pipeline import CONTOUR_VARS, ContourGather, KeffLwaPlan\n"
        "lat, lon, q = synth_c4(6, 91, 180)\n"
        "dA = latlon_cell_area(lat, lon).astype(np.float32)\n"
        "plan = KeffLwaPlan(lat, lon, dA, 41)\n"
        "os.environ.setdefault('MASTER_ADDR', '127.0.0.1'); os.environ.setdefault('MASTER_PORT', '29577')\n"
        "dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device('cuda', 0))\n"
        "g = ContourGather(3, 41, torch.device('cuda', 0), nbuf=2)\n"
        "qd = ops.to_dev(q)\n"
        "outs = [plan.alloc_outputs(3, lwa=False), plan.alloc_outputs(3, lwa=False)]\n"
        "recvs = []\n"
        "for b in range(2):\n"
        "    plan.run(qd[3 * b:3 * b + 3], out=outs[b])\n"
        "    recvs.append(g.launch(outs[b].packed)[0])\n"
        "g.wait(); torch.cuda.synchronize()\n"
        "for b in range(2):\n"
        "    assert recvs[b].shape == (1, len(CONTOUR_VARS), 3, 41)\n"
        "    assert torch.equal(recvs[b][0].nan_to_num(), outs[b].packed.nan_to_num())\n"
        "    un = ContourGather.unpack(recvs[b])\n"
        "    for k in CONTOUR_VARS:\n"
        "        assert torch.equal(un[k].nan_to_num(), outs[b][k].nan_to_num())\n"
        "dist.destroy_process_group()\n"
        "print('gather ok')\n" % (ROOT, ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0 and "gather ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_lwa_f32_opt_in_is_the_fp64_field_rounded_once(ops):
    """alloc_outputs(lwa_dtype=float32) (opt-in, documented as not a drop-in): the field equals the fp64 result
    cast to fp32, bit for bit; everything else is unchanged."""
    from xcontour_b200.pipeline import KeffLwaPlan
    lat, lon, q = synth_c4(3, 181, 360)
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    plan = KeffLwaPlan(lat, lon, dA, 91)
    qd = dev(ops, q)
    a = plan.run(qd)
    b = plan.run(qd, out=plan.alloc_outputs(3, lwa_dtype=torch.float32))
    torch.cuda.synchronize()
    assert b["lwa"].dtype == torch.float32
    assert torch.equal(b["lwa"], a["lwa"].to(torch.float32))
    for k in a:
        if k != "lwa":
            assert torch.equal(a[k].nan_to_num(), b[k].nan_to_num()), k


# ---------------------------------------------------------------- lead dims paired by name (xarray's broadcasting)
def test_interp_to_coords_pairs_slices_by_dimension_name(ops, vort):
    """eqCoords (time, contour) against var (lev, time, contour): apply_ufunc (core.py:1089-1097) broadcasts by name,
    so every (time, lev) slice of var is interpolated with the eqCoords row of ITS time; result dims (time, lev, new).
    A second var whose lead sizes cannot be paired with eqCoords raises."""
    import xcontour_b200 as xb
    lat, lon, q = vort
    tr = xb.DataArray(q[::8, ::8].copy(), dims=("latitude", "longitude"),
                      coords={"latitude": lat[::8].copy(), "longitude": lon[::8].copy()}, name="pv")
    an = xb.Contour2D(tr, tr, dims={"X": "longitude", "Y": "latitude"}, dimEq={"Y": "latitude"})
    rng = np.random.default_rng(11)
    T, L, N, M = 3, 2, 17, 9
    e = np.sort(rng.random((T, N)) * 90.0, axis=-1)
    v = rng.standard_normal((L, T, N))
    pre = np.linspace(-5.0, 95.0, M)
    ex = xb.DataArray(e, dims=("time", "contour"), coords={"time": np.arange(T)}, name="latEq")
    vx = xb.DataArray(v, dims=("lev", "time", "contour"), coords={"lev": np.array([850.0, 500.0])}, name="v")
    out = an.interp_to_coords(pre, ex, vx)
    assert out.dims == ("time", "lev", "new") and out.shape == (T, L, M) and out.name == "v"
    assert np.array_equal(out["lev"].values, [850.0, 500.0]) and np.array_equal(out["new"].values, pre)
    for t in range(T):
        for l in range(L):
            assert np.array_equal(out.values[t, l], O.interp1d(pre, e[t], v[l, t], True))
    with pytest.raises(Exception, match="cannot align"):
        an.interp_to_coords(pre, ex, xb.DataArray(rng.random((T + 1, N)), dims=("time", "contour"), name="w"))


# ---------------------------------------------------------------- NumPy regime of the per-'time' edges, fused path
@pytest.mark.parametrize("increase,lt", [(True, True), (False, False), (True, False)])
def test_fused_batch_in_both_numpy_regimes(ops, monkeypatch, increase, lt):
    """core.py:1273-1281 builds the per-'time' edges with NumPy scalars: fp64 edges under NumPy 1.x, fp32 under
    NEP 50.  xc_keff_lwa_batch (ABI 3: numpy2_rules) reproduces either; each is held to the oracle's statement of
    the same regime, and the Contour2D hist path agrees with the fused path in both."""
    import xcontour_b200 as xb
    from xcontour_b200 import utils as xutils
    from xcontour_b200.pipeline import KeffLwaPlan
    lat, lon, q = synth_c4(3, 91, 180)
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    N = 41
    ctr = O.cal_contours(q, N, increase)
    grd = O.squared_gradient_latlon(q, lat, lon)
    areas = {}
    for rules in ("numpy1", "numpy2"):
        plan = KeffLwaPlan(lat, lon, dA, N, increase=increase, lt=lt, scalar_rules=rules)
        out = plan.run(torch.as_tensor(q, device="cuda"))
        torch.cuda.synchronize()
        ref_a = O.cal_integral_within_contours_hist(q, ctr, dA, lt, scalar_rules=rules)
        ref_g = O.cal_integral_within_contours_hist(q, ctr, dA, lt, integrand=grd, scalar_rules=rules)
        assert np.array_equal(out["ctr"].cpu().numpy().astype(np.float32), ctr)
        assert relmax(out["area"].cpu().numpy(), ref_a) <= RTOL_INT
        assert relmax(out["intgrdS"].cpu().numpy(), ref_g) <= 1e-11
        areas[rules] = out["area"].cpu().numpy()
        # the drop-in API in the same regime
        monkeypatch.setattr(xutils, "NUMPY_SCALAR_RULES", rules)
        coords = {"time": np.arange(3), "latitude": lat, "longitude": lon}
        tr = xb.DataArray(q, dims=("time", "latitude", "longitude"), coords=coords, name="q")
        an = xb.Contour2D(tr, xb.DataArray(dA, dims=("latitude", "longitude")), dims={"X": "longitude", "Y": "latitude"},
                          dimEq={"Y": "latitude"}, increase=increase, lt=lt)
        a_api = an.cal_integral_within_contours_hist(an.cal_contours(N))
        assert relmax(a_api.values, areas[rules]) <= RTOL_INT          # (general kernel: fp64 partial sums)
    # the regimes differ exactly where the reference's do (the cell that holds the maximum, last-edge nudge)
    ref1 = O.cal_integral_within_contours_hist(q, ctr, dA, lt, scalar_rules="numpy1")
    ref2 = O.cal_integral_within_contours_hist(q, ctr, dA, lt, scalar_rules="numpy2")
    assert np.array_equal(areas["numpy1"] != areas["numpy2"], ref1 != ref2)
