"""
The oracle against golden vectors produced by the reference's OWN code
(tests/golden/ref_*.npz, written by tests/golden/make_reference_golden.py from the
unmodified /root/reference/xcontour/core.py running on the oracle/refshim stand-ins
for xarray / xhistogram).  CPU only; the GPU twin is
tests/test_gpu_parity.py::test_product_matches_reference_fixtures.

Bars: everything on the histogram path, the tables, np.interp, d/dA, Keff, the LWA /
LAPE j-loop (both variants) and the integer masks are BIT-EXACT (the oracle performs
the same NumPy operations in the same order); the strict broadcast path sums in the
operands' dtype, so with fp32 cell areas it agrees to fp32 summation-order level.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import xcontour_oracle as O

sys.path.insert(0, GOLDEN)
import make_reference_golden as G       # noqa: E402  (inputs/keys of the fixtures; never touches /root/reference at import)

STRICT_BAR = 5e-4     # fp32 summation order, amplified by the differences d/dA divides (measured: <= 5e-5)

CASES = {                                # name -> (eq dim, leading 'time' dim, with |grad q|^2 chain)
    "ref_vort32": ("Y", False, True),
    "ref_time3": ("Y", True, True),
    "ref_lape": ("Z", True, False),
}


def oracle_chain(fx, eq, lead, with_grd, increase, lt):
    """The same call order as make_reference_golden._run_chain, on the oracle."""
    rules = str(fx["meta/scalar_rules"])
    q = fx["in/q"] if lead else fx["in/q"][None]
    dA, coord, N = fx["in/dA"], fx["in/" + eq], int(fx["in/N"])
    mask = fx["in/mask"] if "in/mask" in fx else np.ones(dA.shape, np.float32)
    out = {}
    ctr = O.cal_contours(q, N, increase)
    per_slice = ctr if lead else ctr[0]                     # contours along 'time' take the per-'time' loop
    tbl, tc = O.cal_area_eqCoord_table_hist(coord, mask, dA, 0, increase, lt, scalar_rules=rules)
    area = O.cal_integral_within_contours_hist(q, per_slice, dA, lt, time_branch=lead, scalar_rules=rules)
    eqc = O.table_lookup_coordinates(area, tbl, tc)
    out.update(ctr=ctr, contour_coord=O.contour_coord(N), table=tbl, table_coord=tc, area=area, eqCoord=eqc)
    with np.errstate(all="ignore"):
        out["dqdA"] = O.cal_gradient_wrt_area(ctr, area)
        if with_grd:
            g = fx["in/grdS"] if lead else fx["in/grdS"][None]
            intg = O.cal_integral_within_contours_hist(q, per_slice, dA, lt, integrand=g, time_branch=lead,
                                                       scalar_rules=rules)
            Lmin = O.latitude_lengths_at(eqc)
            dint = O.cal_gradient_wrt_area(intg, area)
            Leq2 = O.cal_sqared_equivalent_length(dint, out["dqdA"])
            out.update(intgrdS=intg, Lmin=Lmin, dintSdA=dint, Leq2=Leq2, nkeff=O.cal_normalized_Keff(Leq2, Lmin))
            if not lead:                                    # f2 / f3 of SURVEY §8(f), composed as core.py:316-360, 491-616 do
                def hist(integrand):
                    return O.cal_integral_within_contours_hist(q, per_slice, dA, lt, integrand=integrand,
                                                               time_branch=False, scalar_rules=rules)
                out["lwm_hist"] = dint
                out["cm_hist"] = O.cal_gradient_wrt_area(hist(q * g), area) / dint
                predef = fx["out/%s/ctr_at_predef" % G.tag(increase, lt)]
                c2 = O.cal_contours(q, len(predef), increase)
                a2 = O.cal_integral_within_contours_hist(q, c2[0], dA, lt, time_branch=False, scalar_rules=rules)
                out["ctr_at_hist"] = O.interp_to_coords(predef, O.table_lookup_coordinates(a2, tbl, tc), c2)
                out["ctr_at_predef"] = predef[None]
                # conditional-integration twins: the reference sums them in fp32 (loose bars below)
                a_s = O.cal_integral_within_contours(q, per_slice, dA, lt)

                def strict(integrand):
                    return O.cal_integral_within_contours(q, per_slice, dA, lt, integrand=integrand)
                out["lwm_strict"] = O.cal_gradient_wrt_area(strict(g), a_s)
                out["cm_strict"] = O.cal_gradient_wrt_area(strict(q * g), a_s) / out["lwm_strict"]
                t_s, c_s = O.cal_area_eqCoord_table(coord, mask, dA, 0, increase, lt)
                out["ctr_at_strict"] = O.interp_to_coords(
                    predef, O.table_lookup_coordinates(O.cal_integral_within_contours(q, c2[0], dA, lt), t_s, c_s), c2)
    out["table_strict"] = O.cal_area_eqCoord_table(coord, mask, dA, 0, increase, lt)[0]
    out["area_strict"] = O.cal_integral_within_contours(q, per_slice, dA, lt)
    Q = O.interp_to_coords(coord.astype(np.float32), eqc, ctr)
    out["Q"] = Q
    idx = [int(i) for i in fx["in/mask_idx"]]
    for part in ("all", "upper", "lower"):
        if part == "all":
            lwa, cs, ms = O.cal_local_wave_activity(q, Q, dA, coord, increase, part, mask_idx=idx)
            out["lwa_masks"] = np.stack(ms).astype(np.int8)
            out["lwa_contours"] = np.stack(cs)
        else:
            lwa = O.cal_local_wave_activity(q, Q, dA, coord, increase, part)
        out["lwa_" + part] = lwa
    lwa2, cs2, ms2 = O.cal_local_wave_activity(q, Q, dA, coord, increase, "all", mask_idx=idx, variant=2)
    out["lwa2_all"] = lwa2
    out["lwa2_masks"] = np.stack(ms2).astype(np.int8)
    if not lead:                                            # the 2-D case carries no slice axis
        for k in out:
            if k in ("lwa_masks", "lwa2_masks", "lwa_contours"):
                out[k] = out[k][:, 0]
            elif k not in ("contour_coord", "table", "table_coord", "table_strict"):
                out[k] = out[k][0]
    return out


def combos_of(fx):
    return sorted(set(k.split("/")[1] for k in fx if k.startswith("out/")) - {"golden_pv"})


def test_reference_run_reproduces_the_published_notebook_levels():
    """The reference's own cal_contours, run here (NumPy 2.x) on the printed min/max of
    notebooks/1.Keff_atmos.ipynb:102-119, returns the printed fp32 levels digit for
    digit -- and so does the oracle.  (np.vectorize hands `levels` over as np.int64, so
    `1.0/divisor` is a float64 and the steps are fp64 under either NumPy regime.)"""
    import json
    g = json.load(open(os.path.join(GOLDEN, "contours_pv.json")))
    rows = np.array([[np.float32(x) for x in r] for r in g["printed"]])
    fx = G.load("ref_vort32")
    got = fx["out/golden_pv/ctr"]
    assert got.dtype == np.float32 and got.shape == (len(rows), g["levels_N"])
    assert np.array_equal(got[:, g["columns"]], rows)
    assert np.array_equal(O.cal_contours(fx["in/golden_pv_q"], g["levels_N"], True), got)


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_reproduces_the_reference_run(case):
    eq, lead, with_grd = CASES[case]
    fx = G.load(case)
    n_checked = 0
    for t in combos_of(fx):
        increase, lt = t.startswith("inc"), t.endswith("_lt")
        got = oracle_chain(fx, eq, lead, with_grd, increase, lt)
        for key in sorted(k for k in fx if k.startswith("out/%s/" % t)):
            name = key.split("/")[2]
            want, have = fx[key], np.asarray(got[name])
            assert have.shape == want.shape, (case, t, name, have.shape, want.shape)
            if name in ("table_strict", "area_strict") and want.dtype == np.float32:
                # fp32 sums of the broadcast path: order of summation only
                scale = np.nanmax(np.abs(want))
                assert np.nanmax(np.abs(have.astype(np.float64) - want)) <= 3e-5 * scale, (case, t, name)
            elif name in ("lwm_strict", "cm_strict", "ctr_at_strict"):
                # quotients of DIFFERENCES of those fp32 sums: structure check only
                f = np.isfinite(want) & np.isfinite(have)
                assert f.mean() > 0.8 and STRICT_BAR >= np.max(np.abs(have[f] - want[f])) / np.max(np.abs(want[f])), \
                    (case, t, name, np.max(np.abs(have[f] - want[f])) / np.max(np.abs(want[f])))
            else:
                assert have.dtype == want.dtype or name in ("lwa_contours",), (case, t, name, have.dtype, want.dtype)
                assert np.array_equal(have, want, equal_nan=want.dtype.kind == "f"), (case, t, name)
            n_checked += 1
    assert n_checked >= 30


def test_numpy1_and_numpy2_rules_differ_only_in_the_last_edge_nudge():
    """The one place where the NumPy regime reaches the results (core.py:1277-1278):
    under the 1.x rules the per-'time' edges are fp64 and xhistogram's +1e-8 closes the
    last bin over the maximum cell; under NEP 50 they stay fp32, the nudge is a no-op at
    |q| ~ 1 and the maximum cell is dropped.  Everything below the last level agrees."""
    fx = G.load("ref_time3")
    q, dA, N = fx["in/q"], fx["in/dA"], int(fx["in/N"])
    ctr = O.cal_contours(q, N, True)
    a1 = O.cal_integral_within_contours_hist(q, ctr, dA, True, scalar_rules="numpy1")
    a2 = O.cal_integral_within_contours_hist(q, ctr, dA, True, scalar_rules="numpy2")
    assert np.array_equal(a1[:, :-1], a2[:, :-1])
    for s in range(q.shape[0]):
        top = np.nansum(np.where(q[s] == np.nanmax(q[s]), dA, 0.0).astype(np.float64))
        assert a1[s, -1] - a2[s, -1] == pytest.approx(top, rel=1e-12)
    assert np.array_equal(a2, fx["out/inc_lt/area"])        # what the reference computes under NumPy 2


@pytest.mark.skipif(not os.path.isdir("/root/reference/xcontour"), reason="reference checkout not present")
def test_committed_fixtures_are_what_the_reference_produces():
    """Re-runs the reference on the stand-ins and compares with the committed files."""
    r = subprocess.run([sys.executable, "-B", os.path.join(GOLDEN, "make_reference_golden.py"), "--check"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_refshim_labelled_array_semantics():
    """The few xarray behaviours the fixtures rest on (oracle/refshim/xarray_shim.py)."""
    from oracle.refshim import xarray_shim as xr
    a = xr.DataArray(np.arange(6, dtype=np.float32).reshape(2, 3), dims=("y", "x"), coords={"y": [10., 20.]}, name="a")
    c = xr.DataArray(np.array([1.5, 3.5]), dims="contour")
    m = a.where(a < c)                                      # broadcast by name, NaN fill, float32 kept
    assert m.dims == ("y", "x", "contour") and m.dtype == np.float32 and np.isnan(m.values[1, 2, 1])
    assert (m * a).sum(["y", "x"]).values.tolist() == [1.0, 14.0]       # skipna
    assert xr.where(a > 2, -1, 0).dtype == np.int64
    assert xr.where(a > 2, -1, 0).where(a > 3).dtype == np.float64      # ints widen to hold NaN
    assert a.isel({"y": -1}).dims == ("x",) and float(a["y"][-1]) == 20.0
    assert (a.isel(y=1) - a).dims == ("x", "y") and (a - a.isel(y=1)).dims == ("y", "x")
    g = xr.DataArray(np.array([0., 1., 4., 9.]), dims="contour", coords={"contour": np.arange(4, dtype=np.float32)})
    assert g.differentiate("contour").values.tolist() == [1.0, 2.0, 4.0, 5.0]
    s = xr.concat([a.isel(y=0), a.isel(y=1)], "t")
    assert s.dims == ("t", "x")

    def f(x, n):
        assert isinstance(x, np.floating) and isinstance(n, np.integer)      # np.vectorize hands over NumPy scalars
        return x + np.arange(n)
    r = xr.apply_ufunc(f, a.min(dim=["x"]), 3, input_core_dims=[[], []], output_core_dims=[["k"]],
                       vectorize=True, output_dtypes=[np.float32])
    assert r.dims == ("y", "k") and r.dtype == np.float32 and r.values[1].tolist() == [3.0, 4.0, 5.0]


# ---------------------------------------------------------------------------
# GPU twin: the drop-in Contour2D (CUDA through the C ABI) on the same inputs,
# against what the reference's own code produced.
# ---------------------------------------------------------------------------
def _relmax(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.isinf(a), np.isinf(b))
    f = np.isfinite(b)
    if not f.any():
        return 0.0
    return np.max(np.abs(a[f] - b[f])) / max(np.max(np.abs(b[f])), 1e-300)


def _close(a, b, rtol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.isinf(a), np.isinf(b))
    f = np.isfinite(b)
    assert np.allclose(a[f], b[f], rtol=rtol, atol=0), np.abs(a[f] - b[f]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_product_matches_reference_fixtures(case, monkeypatch):
    """Bars: levels and integer masks bit-exact; integrals <= 1e-12 and LWA / LAPE
    fields <= 1e-10 of the array maximum (BASELINE.json north_star); contour-space
    derivatives to 1e-8 per element (they divide differences of the integrals).  The
    reference's strict path sums in fp32 when dA is fp32; the kernels sum in fp64."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import xcontour_b200 as xb
    from xcontour_b200 import utils as xutils
    eq, lead, with_grd = CASES[case]
    fx = G.load(case)
    monkeypatch.setattr(xutils, "NUMPY_SCALAR_RULES", str(fx["meta/scalar_rules"]))
    dims2 = (eq, "X")
    coords = {eq: fx["in/" + eq], "X": fx["in/X"]}
    dims_q, coords_q = (("time",) + dims2, dict(coords, time=fx["in/time"])) if lead else (dims2, coords)
    tr = xb.DataArray(fx["in/q"], dims=dims_q, coords=coords_q, name="trc")
    dAx = xb.DataArray(fx["in/dA"], dims=dims2, coords=coords, name="dA")
    mk = xb.DataArray(fx["in/mask"] if "in/mask" in fx else np.ones(fx["in/dA"].shape, np.float32),
                      dims=dims2, coords=coords, name="mask")
    idx = [int(i) for i in fx["in/mask_idx"]]
    fp32_sums = fx["in/dA"].dtype == np.float32
    for t in combos_of(fx):
        increase, lt = t.startswith("inc"), t.endswith("_lt")
        want = {k.split("/")[2]: fx[k] for k in fx if k.startswith("out/%s/" % t)}
        an = xb.Contour2D(tr, dAx, dims={"X": "X", eq: eq}, dimEq={eq: eq}, increase=increase, lt=lt)
        ctr = an.cal_contours(int(fx["in/N"]))
        assert ctr.dtype == np.float32 and np.array_equal(ctr.values, want["ctr"])
        assert np.array_equal(ctr["contour"].values, want["contour_coord"])
        table = an.cal_area_eqCoord_table_hist(mk)
        area = an.cal_integral_within_contours_hist(ctr).rename("intArea")
        eqc = table.lookup_coordinates(area).rename("eqCoord")
        assert np.array_equal(np.asarray(table._coord.values, np.float64), want["table_coord"].astype(np.float64))
        assert _relmax(table._table.values, want["table"]) <= 1e-12
        assert _relmax(area.values, want["area"]) <= 1e-12
        assert _relmax(eqc.values, want["eqCoord"]) <= 1e-11
        dq = an.cal_gradient_wrt_area(ctr, area)
        _close(dq.values, want["dqdA"], 1e-8)
        if with_grd:
            gx = xb.DataArray(fx["in/grdS"], dims=dims_q, coords=coords_q, name="grdS")
            intg = an.cal_integral_within_contours_hist(ctr, integrand=gx).rename("intgrdS")
            Lmin = xb.latitude_lengths_at(eqc).rename("Lmin")
            dint = an.cal_gradient_wrt_area(intg, area)
            Leq2 = an.cal_sqared_equivalent_length(dint, dq)
            nk = an.cal_normalized_Keff(Leq2, Lmin)
            assert _relmax(intg.values, want["intgrdS"]) <= 1e-12
            assert _relmax(Lmin.values, want["Lmin"]) <= 1e-11
            _close(dint.values, want["dintSdA"], 1e-8)
            _close(Leq2.values, want["Leq2"], 1e-8)
            _close(nk.values, want["nkeff"], 1e-8)
            if "lwm_hist" in want:
                _close(an.cal_contour_weigh_mean_hist(ctr, gx).values, want["lwm_hist"], 1e-8)
                _close(an.cal_contour_mean_hist(ctr, tr, gx).values, want["cm_hist"], 1e-8)
                at = an.cal_contours_at_hist(want["ctr_at_predef"], table)
                assert _relmax(at.values, want["ctr_at_hist"]) <= 1e-11
        if "area_strict" in want:
            bar = 3e-5 if fp32_sums else 1e-12
            assert _relmax(an.cal_integral_within_contours(ctr).values, want["area_strict"]) <= bar
            assert _relmax(an.cal_area_eqCoord_table(mk)._table.values, want["table_strict"]) <= bar
        pre = tr[eq].astype(np.float32)
        ds = an.interp_to_dataset(pre, eqc, xb.merge([ctr, area, eqc]))
        Q = ds["trc"]
        assert _relmax(Q.values, want["Q"]) <= 1e-11
        for part in ("all", "upper", "lower"):
            if "lwa_" + part not in want:
                continue
            if part == "all":
                lwa, cs, ms = an.cal_local_wave_activity(tr, Q, mask_idx=idx, part=part)
                assert np.array_equal(np.stack([np.asarray(m.values) for m in ms]).astype(np.int8), want["lwa_masks"])
            else:
                lwa = an.cal_local_wave_activity(tr, Q, part=part)
            assert lwa.dims == tr.dims and _relmax(lwa.values, want["lwa_" + part]) <= 1e-10
        lwa2, cs2, ms2 = an.cal_local_wave_activity2(tr, Q, mask_idx=idx, part="all")
        assert _relmax(lwa2.values, want["lwa2_all"]) <= 1e-10
        assert np.array_equal(np.stack([np.asarray(m.values) for m in ms2]).astype(np.int8), want["lwa2_masks"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/xcontour"), reason="reference checkout not present")
@pytest.mark.parametrize("increase,lt", [(True, True), (False, False)])
def test_oracle_equals_reference_on_the_full_vorticity_field(vort, increase, lt):
    """Not a stored fixture (the fields are too large to commit): the reference's own code,
    run here on the full 256x512 Data/barotropic_vorticity.nc field with N = 121 (the setup of
    tests/test_LWA.py), against the oracle -- bit for bit, Keff chain and LWA."""
    code = r'''
import sys, numpy as np
sys.dont_write_bytecode = True
sys.path.insert(0, %r)
from oracle import refshim, xcontour_oracle as O
ref = refshim.load_reference("/root/reference")
import xarray as xr
d = np.load(%r)
lat, lon, q = d["latitude"], d["longitude"], d["absolute_vorticity"]
increase, lt, N = %r, %r, 121
dA = O.latlon_cell_area(lat, lon).astype(np.float32)
grd = O.squared_gradient_latlon(q, lat, lon).astype(np.float32)
co = {"Y": lat, "X": lon}
tr = xr.DataArray(q, dims=("Y", "X"), coords=co, name="vor")
gx = xr.DataArray(grd, dims=("Y", "X"), coords=co, name="grdS")
an = ref.Contour2D(tr, xr.DataArray(dA, dims=("Y", "X"), coords=co), dims={"X": "X", "Y": "Y"},
                   dimEq={"Y": "Y"}, increase=increase, lt=lt)
ctr = an.cal_contours(N)
table = an.cal_area_eqCoord_table_hist(xr.DataArray(np.ones_like(q), dims=("Y", "X"), coords=co, name="m"))
area = an.cal_integral_within_contours_hist(ctr).rename("intArea")
intg = an.cal_integral_within_contours_hist(ctr, integrand=gx).rename("intgrdS")
latEq = table.lookup_coordinates(area).rename("latEq")
nk = an.cal_normalized_Keff(an.cal_sqared_equivalent_length(an.cal_gradient_wrt_area(intg, area),
                                                            an.cal_gradient_wrt_area(ctr, area)),
                            ref.latitude_lengths_at(latEq))
Q = an.interp_to_dataset(tr["Y"].astype(np.float32), latEq, [ctr, area, latEq])["vor"]
lwa = an.cal_local_wave_activity(tr, Q)
q3 = q[None]
o_ctr = O.cal_contours(q3, N, increase)
tbl, c = O.cal_area_eqCoord_table_hist(lat, np.ones_like(q), dA, 0, increase, lt)
o_area = O.cal_integral_within_contours_hist(q3, o_ctr[0], dA, lt, time_branch=False)
o_intg = O.cal_integral_within_contours_hist(q3, o_ctr[0], dA, lt, integrand=grd[None], time_branch=False)
o_latEq = O.table_lookup_coordinates(o_area, tbl, c)
with np.errstate(all="ignore"):
    o_nk = O.cal_normalized_Keff(O.cal_sqared_equivalent_length(O.cal_gradient_wrt_area(o_intg, o_area),
                                                                O.cal_gradient_wrt_area(o_ctr, o_area)),
                                 O.latitude_lengths_at(o_latEq))
o_Q = O.interp_to_coords(lat, o_latEq, o_ctr)
o_lwa = O.cal_local_wave_activity(q3, o_Q, dA, lat, increase)
for name, a, b in (("ctr", ctr.values, o_ctr[0]), ("area", area.values, o_area[0]), ("intgrdS", intg.values, o_intg[0]),
                   ("latEq", latEq.values, o_latEq[0]), ("nkeff", nk.values, o_nk[0]), ("Q", Q.values, o_Q[0]),
                   ("lwa", lwa.values, o_lwa[0])):
    assert a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True), name
print("identical")
''' % (ROOT, os.path.join(GOLDEN, "barotropic_vorticity.npz"), increase, lt)
    r = subprocess.run([sys.executable, "-B", "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "identical" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
