"""
Golden vectors from the reference's OWN code.

Runs the unmodified ``/root/reference/xcontour/core.py`` (imported in place, never
copied) on top of ``oracle/refshim`` -- a minimal stand-in for xarray / xhistogram,
which are absent from this image -- and stores inputs and outputs as one small
``.npz`` per case under tests/golden/.  Build container only; the tests read the
committed files (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_reference_golden.py          # rewrite the fixtures
    python tests/golden/make_reference_golden.py --check  # compare with the committed ones

What the fixtures pin: every line of the reference's hot path that is the reference's
own (levels, histogram edges, flips, CDF direction, table end points, np.interp
direction, d/dA, Keff, the LWA / LAPE mask algebra and j-loop, variant 2), executed
verbatim.  What they cannot pin: the internals of xarray and xhistogram, restated in
the stand-in from their documented behaviour (oracle/refshim/*.py headers).

NumPy regime: the fixtures are generated under the installed NumPy (2.x, NEP 50
scalar promotion).  The reference was written for NumPy 1.x; the two regimes differ
in one place on this path -- ``step = (c_last - c_first)/(len-1)`` of
``_histogram`` (core.py:1277, 1300) is fp32 under NEP 50 and fp64 under the 1.x
rules, which decides the dtype of the per-'time' edge array (core.py:1278) and hence
whether xhistogram's ``+1e-8`` nudge of the last edge is a no-op.  The oracle restates
both (``scalar_rules="numpy1"`` -- its default -- and ``"numpy2"``), the CUDA path implements both and follows the
installed NumPy by default; the fixture tests select the regime recorded in each file.

Cases
-----
ref_vort32.npz   Data/barotropic_vorticity.nc sub-sampled to 32x32, dims (Y, X): the call
                 order of tests/test_Keff_atmos.py:76-92 + tests/test_LWA.py:57-81 for all
                 four (increase, lt); static histogram bins.
ref_time3.npz    synthetic (time=3, Y=24, X=32) with NaN cells and repeated values:
                 contours vary along 'time' -> the per-'time' loop of _histogram.
ref_lape.npz     synthetic X-Z plane (time=2, Z=20, X=28), decreasing Z coordinate,
                 topography (NaN), fp64 cell areas, increase=False, lt=False: the setup of
                 tests/test_LAPE.py:56-100 (cal_local_APE, cal_local_wave_activity2).

Dimension names are the dict keys ('X', 'Y', 'Z'): the reference's strict and LWA paths
sum over ``dims.keys()`` (core.py:130, 404, 789), so with other names they raise
(SURVEY.md §8a H1); with key == value every path runs unmodified.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

COMBOS = [(True, True), (True, False), (False, True), (False, False)]


def tag(increase, lt):
    return ("inc" if increase else "dec") + ("_lt" if lt else "_gt")


# ------------------------------------------------------------------ inputs
def inputs_vort32():
    from oracle import xcontour_oracle as O
    d = np.load(os.path.join(HERE, "barotropic_vorticity.npz"))
    lat, lon = d["latitude"][::8].copy(), d["longitude"][::16].copy()
    q = d["absolute_vorticity"][::8, ::16].copy()
    dA = O.latlon_cell_area(lat, lon).astype(np.float32)
    grd = O.squared_gradient_latlon(q, lat, lon).astype(np.float32)
    return dict(Y=lat, X=lon, q=q, dA=dA, grdS=grd, N=np.int64(17), mask_idx=np.array([3, 20]))


def inputs_time3():
    rng = np.random.default_rng(20260317)
    ny, nx, S = 24, 32, 3
    Y = np.linspace(-57.5, 57.5, ny).astype(np.float32)
    X = (np.arange(nx) * (360.0 / nx)).astype(np.float32)
    phi, lam = np.deg2rad(Y.astype(np.float64))[:, None], np.deg2rad(X.astype(np.float64))[None, :]
    q = np.empty((S, ny, nx), np.float32)
    for s in range(S):
        q[s] = (np.sin(phi) + 0.25 * np.cos(phi) ** 2 * np.sin(3 * lam + 2 * phi + s)
                + 0.03 * rng.standard_normal((ny, nx))).astype(np.float32)
    q = np.round(q * 64) / np.float32(64)              # repeated values: ties with the levels and with Q
    q[1, 5:8, 10:14] = np.nan                          # undefined cells in one slice
    dA = (np.cos(phi) * np.ones_like(lam) * 1.0e9).astype(np.float32)
    grd = (rng.random((S, ny, nx)) * 1e-9).astype(np.float32)
    return dict(Y=Y, X=X, time=np.arange(S, dtype=np.int64), q=q, dA=dA, grdS=grd, N=np.int64(9),
                mask_idx=np.array([2, 17]))


def inputs_lape():
    rng = np.random.default_rng(7)
    nz, nx, S = 20, 28, 2
    Z = -(5.0 + 10.0 * np.arange(nz)).astype(np.float32)               # decreasing coordinate
    X = (50.0 * np.arange(nx)).astype(np.float32)
    b = np.empty((S, nz, nx), np.float64)
    for s in range(S):
        zz = Z.astype(np.float64)[:, None] / 200.0
        b[s] = 0.02 * zz + 0.004 * np.sin(2 * np.pi * (X[None, :] / X[-1]) + s) * np.exp(zz) \
            + 0.0005 * rng.standard_normal((nz, nx))
    b = b.astype(np.float32)
    maskC = np.ones((nz, nx), np.float32)
    for i in range(nx):                                                # a sloping bottom on the right
        depth = nz - max(0, (i - 16) // 2)
        maskC[depth:, i] = 0
    b[:, maskC == 0] = np.nan
    dz = np.full(nz, 10.0); dz[-5:] = 14.0
    dA = (dz[:, None] * np.full(nx, 50.0)[None, :]).astype(np.float64)
    return dict(Z=Z, X=X, time=np.arange(S, dtype=np.int64), q=b, dA=dA, mask=maskC, N=np.int64(13),
                mask_idx=np.array([4, 15]))


# --------------------------------------------------------- running the reference
def _run_chain(ref, xr, inp, eq, lead, increase, lt, out, with_grd=True, parts=("all",), variant2=True,
               strict=True, mask=None):
    """The call order of the reference's own scripts for one (increase, lt)."""
    t = tag(increase, lt)
    dims2 = (eq, "X")
    coords = {eq: inp[eq], "X": inp["X"]}
    if lead:
        coords_q = dict(coords, time=inp["time"])
        dims_q = ("time",) + dims2
    else:
        coords_q, dims_q = coords, dims2
    tr = xr.DataArray(inp["q"], dims=dims_q, coords=coords_q, name="trc")
    dAx = xr.DataArray(inp["dA"], dims=dims2, coords=coords, name="dA")
    an = ref.Contour2D(tr, dAx, dims={"X": "X", eq: eq}, dimEq={eq: eq}, increase=increase, lt=lt)
    ctr = an.cal_contours(int(inp["N"]))
    if mask is None:
        mk = xr.DataArray(np.ones(inp["dA"].shape, np.float32), dims=dims2, coords=coords, name="mask")
    else:
        mk = xr.DataArray(mask, dims=dims2, coords=coords, name="mask")
    table = an.cal_area_eqCoord_table_hist(mk)
    area = an.cal_integral_within_contours_hist(ctr).rename("intArea")
    eqc = table.lookup_coordinates(area).rename("eqCoord")
    out[t + "/ctr"] = ctr.values
    out[t + "/contour_coord"] = ctr["contour"].values
    out[t + "/table"] = table._table.values
    out[t + "/table_coord"] = table._coord.values
    out[t + "/area"] = area.values
    out[t + "/eqCoord"] = eqc.values
    dq = an.cal_gradient_wrt_area(ctr, area)
    out[t + "/dqdA"] = dq.values
    if with_grd:
        gx = xr.DataArray(inp["grdS"], dims=dims_q, coords=coords_q, name="grdS")
        intg = an.cal_integral_within_contours_hist(ctr, integrand=gx).rename("intgrdS")
        Lmin = ref.latitude_lengths_at(eqc).rename("Lmin")
        dint = an.cal_gradient_wrt_area(intg, area)
        Leq2 = an.cal_sqared_equivalent_length(dint, dq)
        nk = an.cal_normalized_Keff(Leq2, Lmin)
        for k, v in (("intgrdS", intg), ("Lmin", Lmin), ("dintSdA", dint), ("Leq2", Leq2), ("nkeff", nk)):
            out[t + "/" + k] = v.values
        if not lead:
            # along-contour means (core.py:491-616) and levels at
            # prescribed equivalent coordinates (core.py:316-360)
            out[t + "/lwm_hist"] = an.cal_contour_weigh_mean_hist(ctr, gx).values
            out[t + "/cm_hist"] = an.cal_contour_mean_hist(ctr, tr, gx).values
            predef = np.linspace(float(inp[eq][2]), float(inp[eq][-3]), 11).astype(np.float32)
            out[t + "/ctr_at_hist"] = an.cal_contours_at_hist(predef, table).values
            out[t + "/ctr_at_predef"] = predef
            # the conditional-integration twins (fp32 sums in the reference)
            out[t + "/lwm_strict"] = an.cal_contour_weigh_mean(ctr, gx).values
            out[t + "/cm_strict"] = an.cal_contour_mean(ctr, tr, gx).values
            out[t + "/ctr_at_strict"] = an.cal_contours_at(predef, an.cal_area_eqCoord_table(mk)).values
    if strict:
        out[t + "/table_strict"] = an.cal_area_eqCoord_table(mk)._table.values
        out[t + "/area_strict"] = an.cal_integral_within_contours(ctr).values
    pre = tr[eq].astype(np.float32)
    ds = an.interp_to_dataset(pre, eqc, xr.merge([ctr, area, eqc]))
    Q = ds["trc"]
    out[t + "/Q"] = Q.values
    idx = [int(i) for i in inp["mask_idx"]]
    for part in parts:
        lwa, cs, ms = an.cal_local_wave_activity(tr, Q, mask_idx=idx, part=part)
        out[t + "/lwa_" + part] = lwa.values
        if part == "all":
            out[t + "/lwa_masks"] = np.stack([m.transpose(*tr.dims).values for m in ms]).astype(np.int8)
            out[t + "/lwa_contours"] = np.stack([np.asarray(c.values, np.float64) for c in cs])
    if variant2:
        lwa2, cs2, ms2 = an.cal_local_wave_activity2(tr, Q, mask_idx=idx, part="all")
        out[t + "/lwa2_all"] = lwa2.values
        out[t + "/lwa2_masks"] = np.stack([m.transpose(*tr.dims).values for m in ms2]).astype(np.int8)


def _committed_inputs(name):
    """Inputs of a committed fixture (--check re-runs the reference on exactly these, so the
    comparison does not depend on the random generator of the installed NumPy)."""
    return {k[3:]: v for k, v in load(name).items() if k.startswith("in/")}


def generate(from_committed=False):
    from oracle import refshim
    ref = refshim.load_reference(REF)
    import xarray as xr                                 # the stand-in registered by load_reference
    regime = "numpy2" if int(np.__version__.split(".")[0]) >= 2 else "numpy1"
    cases = {}

    inp = _committed_inputs("ref_vort32") if from_committed else inputs_vort32()
    inp.pop("golden_pv_q", None)
    out = {}
    for inc, lt in COMBOS:
        _run_chain(ref, xr, inp, "Y", False, inc, lt, out, parts=("all", "upper", "lower") if inc == lt else ("all",))
    # the one published vector (notebooks/1.Keff_atmos.ipynb:102-119, contours_pv.json): a
    # two-cell plane per isentropic level holding the printed min and max reproduces the
    # reference's cal_contours(121) rows
    import json
    g = json.load(open(os.path.join(HERE, "contours_pv.json")))
    rows = np.array([[np.float32(x) for x in r] for r in g["printed"]])
    qpv = np.stack([rows[:, 0], rows[:, -1]], axis=1).astype(np.float32).reshape(-1, 1, 2)
    trpv = xr.DataArray(qpv, dims=("level", "Y", "X"), name="pv")
    anpv = ref.Contour2D(trpv, xr.DataArray(np.ones((1, 2), np.float32), dims=("Y", "X")),
                         dims={"X": "X", "Y": "Y"}, dimEq={"Y": "Y"}, increase=True, lt=True)
    inp["golden_pv_q"] = qpv
    out["golden_pv/ctr"] = anpv.cal_contours(int(g["levels_N"])).values
    cases["ref_vort32"] = (inp, out)

    inp = _committed_inputs("ref_time3") if from_committed else inputs_time3()
    out = {}
    for inc, lt in [(True, True), (False, False), (True, False)]:
        _run_chain(ref, xr, inp, "Y", True, inc, lt, out, parts=("all", "lower"), strict=(inc == lt))
    cases["ref_time3"] = (inp, out)

    inp = _committed_inputs("ref_lape") if from_committed else inputs_lape()
    out = {}
    for inc, lt in [(False, False), (False, True)]:
        _run_chain(ref, xr, inp, "Z", True, inc, lt, out, with_grd=False, parts=("all", "upper"), mask=inp["mask"])
    cases["ref_lape"] = (inp, out)

    packed = {}
    for name, (inp, out) in cases.items():
        d = {"in/" + k: np.asarray(v) for k, v in inp.items()}
        d.update({"out/" + k: np.asarray(v) for k, v in out.items()})
        d["meta/numpy"] = np.array(np.__version__)
        d["meta/scalar_rules"] = np.array(regime)
        d["meta/array_backend"] = np.array(refshim.BACKEND)
        packed[name] = d
    return packed


def load(name):
    """{key: array} of a committed fixture (used by the tests)."""
    with np.load(os.path.join(HERE, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def main():
    check = "--check" in sys.argv
    packed = generate(from_committed=check)
    bad = 0
    for name, d in packed.items():
        path = os.path.join(HERE, name + ".npz")
        if check:
            old = load(name)
            for k in sorted(set(d) | set(old)):
                if k.startswith("meta/"):
                    continue
                same = k in d and k in old and d[k].dtype == old[k].dtype and np.array_equal(d[k], old[k], equal_nan=d[k].dtype.kind == "f")
                if not same:
                    bad += 1
                    print("DIFF", name, k)
        else:
            np.savez_compressed(path, **d)
            print("wrote %s (%d arrays, %.1f KB)" % (path, len(d), os.path.getsize(path) / 1024.0))
    if check:
        print("fixtures %s" % ("differ" if bad else "are current"))
        sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
