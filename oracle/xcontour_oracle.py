"""
CPU oracle for the contour-coordinate hot path of miniufo/xcontour.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  The product (``xcontour_b200``) never does: it
calls hand-written CUDA through ``libxcb200.so`` and fails loudly without it.

What it is: a plain-NumPy restatement of the reference's arithmetic, function by
function, on bare ndarrays.  Every function cites the reference lines it follows
(paths relative to ``/root/reference``).  The reference cannot be imported as is in
this image (xarray, xhistogram, dask, xgcm, skimage are absent and there is no
network), but its own unmodified ``xcontour/core.py`` does run on the minimal
stand-ins of ``oracle/refshim`` -- that is how the golden vectors were made.

Parity status
-------------
* PINNED to the reference's own code: ``tests/golden/ref_*.npz`` hold inputs and
  outputs of the unmodified reference (levels, histogram/CDF incl. the per-'time'
  loop, both tables, lookups, d/dA, Leq2, nkeff, along-contour means, levels at
  prescribed coordinates, np.interp to the grid, LWA / LAPE for every part, variant 2
  and the integer masks; four (increase, lt) combinations; NaN cells, repeated values,
  a decreasing X-Z coordinate with topography) executed here through
  ``oracle/refshim`` by ``tests/golden/make_reference_golden.py``.  This module
  reproduces them BIT FOR BIT except for the strict broadcast path, which sums in fp32
  and agrees to summation-order level (``tests/test_reference_golden.py``).
* PINNED to the published numbers: ``cal_contours`` reproduces the 36 printed fp32
  values of ``notebooks/1.Keff_atmos.ipynb:102-119`` (tests/golden/contours_pv.json),
  and so does the reference run above.
* NOT pinned (restated from published behaviour, no source in /root/reference): the
  internals of xarray and of xhistogram 0.3.0 (digitize with right=False, last edge
  + 1e-8 in the edge dtype, out-of-range and NaN discarded, bincount weights in fp64;
  call sites ``core.py:1284``, ``1307``) as written down in ``oracle/refshim``; and
  ``squared_gradient_latlon``, which has no counterpart in the reference.
* NumPy regime: the reference was written for NumPy 1.x.  Its scalar-promotion rules
  differ from NEP 50 (NumPy >= 2) in one place on this path, the fp64 ``step`` of
  ``_histogram`` (see ``hist_edges``); ``scalar_rules="numpy1"`` (the default here) and
  ``"numpy2"`` (the regime the fixtures were generated in) restate both.  The CUDA path
  implements both and follows the installed NumPy unless ``XCB200_NUMPY_RULES`` says otherwise.

Array conventions: a tracer is ``q[S, n0, n1]`` (S independent slices, the 2-D
plane last); contour-space arrays are ``[S, N]``; ``dA`` is ``[n0, n1]``.
"""
import numpy as np

Rearth = 6371200.0  # xcontour/utils.py:19


# --------------------------------------------------------------------------
# contour levels                                   xcontour/core.py:205-266
# --------------------------------------------------------------------------
def cal_contours(q, levels, increase=True, dtype=np.float32):
    """Equally spaced levels between the per-slice min and max.

    core.py:222-249.  ``mmin/mmax`` are NaN-skipping reductions over the plane
    (xarray ``.min(dim=...)``).  ``mylinspace`` (core.py:228-232) is evaluated per
    slice through ``np.vectorize`` on NumPy *scalars* -- ``levels`` arrives as an
    np.int64, so ``1.0/divisor`` is a float64 under NumPy 1.x and 2.x alike:
        steps = (1.0/divisor) * (stop - start)     # f32 - f32 -> f32, then f64 * f32 -> f64
        steps * arange(levels) + start             # f64
    and cast to ``self.dtype`` by ``output_dtypes`` (core.py:246).  Pinned by the
    notebook golden vector and by the reference run of tests/golden/ref_vort32.npz.
    An array ``levels`` is broadcast verbatim (core.py:253-264).
    """
    q = np.asarray(q)
    S = q.shape[0]
    if isinstance(levels, (int, np.integer)):
        N = int(levels)
        with np.errstate(invalid="ignore"):
            mmin = np.nanmin(q.reshape(S, -1), axis=1)
            mmax = np.nanmax(q.reshape(S, -1), axis=1)
        start, end = (mmin, mmax) if increase else (mmax, mmin)
        diff = (end - start)                       # in the tracer dtype
        steps = np.float64(1.0 / (N - 1)) * diff.astype(np.float64)
        ctr = steps[:, None] * np.arange(N, dtype=np.float64)[None, :] \
            + start.astype(np.float64)[:, None]
        return ctr.astype(dtype)
    levs = np.asarray(levels)
    mn = np.nanmin(q.reshape(S, -1), axis=1)
    return ((mn[:, None] - mn[:, None]) + levs[None, :]).astype(dtype)


def contour_coord(N, dtype=np.float32):
    """core.py:248-249 -- the float coordinate 0..N-1 of the 'contour' dim."""
    return np.linspace(0.0, N - 1.0, N, dtype=dtype)


# --------------------------------------------------------------------------
# histogram / CDF                       xcontour/core.py:412-460, 1202-1325
# --------------------------------------------------------------------------
def hist_edges(ctr, time_branch=True, scalar_rules="numpy1"):
    """Bin edges the reference hands to xhistogram for one slice.

    One extra bin below the smallest level so the result has N entries;
    decreasing contours are reversed.  Two branches of ``_histogram``:

    * static bins (core.py:1296-1304): ``np.insert(bvalues, 0, bvalues[0]-step)``
      keeps the array dtype of the contours (fp32 by default);
    * bins varying along 'time' (core.py:1273-1281):
      ``np.concatenate([[ctr[0]-step], ctr])`` -- the list holds a float64
      scalar, so the whole edge array is promoted to fp64.

    In both, ``step = (c_last - c_first) / (len - 1)`` is a NumPy scalar of the
    contour dtype divided by a Python int, i.e. fp64 under the NumPy-1.x scalar
    rules the reference was written for (README.md:26 states numpy 1.15.4);
    restated explicitly here.
    ``scalar_rules="numpy2"`` restates the same lines under NEP 50 (NumPy >= 2: the
    step stays in the contour dtype, so both branches keep the contour dtype) --
    the regime in which tests/golden/ref_*.npz were generated from the reference's
    own code (tests/golden/make_reference_golden.py).
    Returns (edges[N+1] ascending, bincrease).
    """
    ctr = np.asarray(ctr)
    N = ctr.shape[0]
    bincrease = bool(ctr[0] < ctr[-1])
    first, last = (ctr[0], ctr[-1]) if bincrease else (ctr[-1], ctr[0])
    body = ctr if bincrease else ctr[::-1]
    if scalar_rules == "numpy2":
        step = (last - first) / ctr.dtype.type(N - 1)  # scalar of ctr.dtype / python int
        return np.concatenate([np.array([first - step], dtype=ctr.dtype), body]), bincrease
    assert scalar_rules == "numpy1", scalar_rules
    step = np.float64(last - first) / (N - 1)          # difference in ctr.dtype
    e0 = np.float64(first) - step
    if time_branch:
        edges = np.concatenate([[e0], body.astype(np.float64)])
    else:
        edges = np.concatenate([np.array([e0]).astype(ctr.dtype), body]).astype(ctr.dtype)
    return edges, bincrease


def xhistogram_1d(x, edges, weights):
    """xhistogram 0.3.0 ``_bincount_2d_vectorized`` for one variable, restated.

    Third-party, NOT under /root/reference (setup.py:40-45 lists bare
    'xhistogram'; README.md:26 states dev version 0.3.0).  Call sites:
    core.py:1284, 1307.  Published algorithm: last edge nudged by ``+1e-8`` *in
    the edge array's dtype*, ``np.digitize(x, edges)`` (right=False),
    ``np.bincount(idx, weights, minlength=len(edges)+1)`` in fp64, then the
    under/overflow slots -- which also collect NaN, because digitize sends NaN to
    ``len(edges)`` -- are dropped.
    """
    edges = np.asarray(edges)
    e = np.concatenate((edges[:-1], edges[-1:] + 1e-8)).astype(edges.dtype)
    idx = np.digitize(np.asarray(x).ravel(), e)
    cnt = np.bincount(idx, weights=np.asarray(weights).ravel(),
                      minlength=len(e) + 1)
    return cnt[1:-1]


def digitize_bins(x, edges):
    """Bin index (0..N-1, or -1 when discarded) under the same rule as
    ``xhistogram_1d`` -- used for the bit-exact bin-assignment parity test."""
    edges = np.asarray(edges)
    e = np.concatenate((edges[:-1], edges[-1:] + 1e-8)).astype(edges.dtype)
    idx = np.digitize(np.asarray(x).ravel(), e)
    out = idx.astype(np.int64) - 1
    out[(idx == 0) | (idx == len(e))] = -1
    return out.reshape(np.shape(x))


def histogram_cdf(x, ctr, weights, lt, time_branch=True, scalar_rules="numpy1"):
    """``_histogram`` for one slice (core.py:1262-1325), in *storage* order
    (ascending bin values), plus ``bincrease``."""
    edges, bincrease = hist_edges(ctr, time_branch, scalar_rules)
    pdf = xhistogram_1d(x, edges, weights)
    cdf = np.cumsum(pdf)                               # core.py:1320
    if not lt:
        cdf = cdf[-1] - cdf                            # core.py:1322-1323
    return cdf, bincrease


def cal_integral_within_contours_hist(q, ctr, dA, lt, integrand=None, time_branch=None,
                                      scalar_rules="numpy1"):
    """core.py:412-460.  ``wei = integrand*dA`` rounded in the operands' common
    dtype (core.py:444), ``fillna(0)`` (449), per-slice histogram CDF, then flip
    so that the contour index ascends (454-455).  The reference loops only over a
    dim named 'time' (core.py:1262-1287); every leading index is treated that way
    here.  Returns fp64 [S, N]."""
    q = np.asarray(q)
    ctr = np.asarray(ctr)
    S = q.shape[0]
    if time_branch is None:          # contours that vary per slice take the 'time' loop
        time_branch = ctr.ndim == 2
    out = np.empty((S, ctr.shape[-1]), dtype=np.float64)
    for s in range(S):
        c = ctr[s] if ctr.ndim == 2 else ctr
        if integrand is not None:
            wei = np.asarray(integrand)[s] * dA
        else:
            wei = np.asarray(dA)
        wei = np.where(np.isnan(wei), 0.0, wei).astype(wei.dtype)
        cdf, binc = histogram_cdf(q[s], c, wei, lt, time_branch, scalar_rules)
        out[s] = cdf if binc else cdf[::-1]
    return out


def cal_integral_within_contours(q, ctr, dA, lt, integrand=None, chunk=16):
    """core.py:363-409 -- strict conditional integration by broadcasting
    (``tracer < contour`` / ``tracer > contour``), NaN terms skipped by ``sum``
    (core.py:1376).  The result dtype is the promoted dtype of integrand*dA."""
    q = np.asarray(q)
    ctr = np.asarray(ctr)
    S, N = q.shape[0], ctr.shape[-1]
    dA = np.asarray(dA)
    res_dtype = np.result_type(q.dtype if integrand is None else
                               np.asarray(integrand).dtype, dA.dtype)
    out = np.empty((S, N), dtype=res_dtype)
    for s in range(S):
        c = ctr[s] if ctr.ndim == 2 else ctr
        if integrand is None:
            g = q[s] - q[s] + 1                       # core.py:396
        else:
            g = np.asarray(integrand)[s]
        for k0 in range(0, N, chunk):
            cc = c[k0:k0 + chunk][:, None, None]
            with np.errstate(invalid="ignore"):
                cond = (q[s][None] < cc) if lt else (q[s][None] > cc)
            msk = np.where(cond, g[None], np.nan)
            out[s, k0:k0 + chunk] = np.nansum(msk * dA[None], axis=(1, 2))
    return out


# --------------------------------------------------------------------------
# A(Yeq) table and lookups              xcontour/core.py:150-203, 1103-1174
# --------------------------------------------------------------------------
def cal_area_eqCoord_table_hist(coord, mask, dA, eq_axis, increase, lt, scalar_rules="numpy1"):
    """core.py:150-203.  Histogram of the eq-coordinate field (NaN where
    ``mask != 1``, core.py:178) against bins = the coordinate vector with weights
    ``dA``; ``ylt = lt if increase == yIncre else not lt`` (180-188).  The table
    is returned in ascending-coordinate storage order together with that
    ascending coordinate (core.py:195-198)."""
    coord = np.asarray(coord)
    mask = np.asarray(mask)
    shp = [1, 1]
    shp[eq_axis] = coord.shape[0]
    ctrVar = np.broadcast_to(coord.reshape(shp), mask.shape).astype(coord.dtype)
    ctrVar = np.where(mask == 1, ctrVar, np.nan)
    yIncre = not (coord[-1] < coord[0])
    ylt = lt if (increase == yIncre) else (not lt)
    cdf, _ = histogram_cdf(ctrVar, coord, dA, ylt, time_branch=False, scalar_rules=scalar_rules)
    coord_asc = coord if yIncre else coord[::-1]
    return cdf, coord_asc.copy()


def cal_area_eqCoord_table(coord, mask, dA, eq_axis, increase, lt):
    """core.py:73-147 -- strict-comparison table with the endpoint replaced by
    the total masked area (133-140).  Stored in the coordinate's own order."""
    coord = np.asarray(coord)
    mask = np.asarray(mask)
    shp = [1, 1]
    shp[eq_axis] = coord.shape[0]
    ctrVar = np.broadcast_to(coord.reshape(shp), mask.shape)
    eqDimIncre = bool(coord[-1] > coord[0])
    use_lt = (eqDimIncre == increase) if lt else (eqDimIncre != increase)
    tbl = np.empty(coord.shape[0], dtype=np.result_type(mask.dtype, dA.dtype))
    for k, c in enumerate(coord):
        cond = (ctrVar < c) if use_lt else (ctrVar > c)
        tbl[k] = abs(np.nansum(np.where(cond, mask, np.nan) * dA))
    maxArea = abs(np.nansum(mask * dA))
    if tbl[-1] > tbl[0]:
        tbl[-1] = maxArea
    else:
        tbl[0] = maxArea
    return tbl, coord.copy()


def interp1d(x, xf, yf, inc=True):
    """core.py:1405-1434 -- np.interp, reversing the table when decreasing."""
    if inc:
        return np.interp(x, xf, yf)
    return np.interp(x, xf[::-1], yf[::-1])


def table_lookup_coordinates(values, table, coord):
    """Table.lookup_coordinates (core.py:1136-1174): x given y = F(x).
    ``areaInc = table[-1] > table[0]`` (core.py:1122-1126)."""
    areaInc = bool(table[-1] > table[0])
    v = np.asarray(values)
    out = np.empty(v.shape, dtype=np.float64)
    for idx in np.ndindex(v.shape[:-1]):
        out[idx] = interp1d(v[idx], table, coord, inc=areaInc)
    return out


def interp_to_coords(predef, eqCoords, var):
    """core.py:1050-1100: per slice ``np.interp(predef, eqCoords[s], var[s])``;
    the direction is detected once from slice [0, 0, ...] (core.py:1080-1088)."""
    eq = np.asarray(eqCoords)
    vv = np.asarray(var)
    vals = eq
    while vals.ndim > 1:
        vals = vals[0]
    increasing = bool(vals[0] < vals[-1])
    out = np.empty(eq.shape[:-1] + (len(predef),), dtype=np.float64)
    for idx in np.ndindex(eq.shape[:-1]):
        out[idx] = interp1d(predef, eq[idx], vv[idx], inc=increasing)
    return out


def cal_contours_equal_area(q, dA, levels, increase=True, lt=True, dtype=np.float32, refine=8):
    """Equal-area levels by a weighted-quantile histogram (NOT in the reference --
    north_star kernel (1); the oracle of Contour2D.cal_contours_equal_area):
    fine equally spaced levels -> area CDF (histogram path) -> np.interp of the
    inverse relation at `levels` equally spaced areas."""
    N = int(levels)
    fine = cal_contours(q, (N - 1) * int(refine) + 1, increase, dtype)
    per = fine if q.shape[0] > 1 else fine[0]
    area = cal_integral_within_contours_hist(q, per, dA, lt)
    w = np.linspace(0.0, 1.0, N)
    tgt = area[:, :1] + (area[:, -1:] - area[:, :1]) * w[None, :]
    inc = bool(area[0, 0] < area[0, -1])
    out = np.stack([interp1d(tgt[s], area[s], fine[s].astype(np.float64), inc) for s in range(q.shape[0])])
    return out.astype(dtype)


def weighted_quantile_levels(q2d, dA, fractions):
    """Exact weighted quantiles of one slice by sorting (the 'segmented radix sort'
    alternative of north_star kernel (1)); used to bound the histogram method."""
    v = np.asarray(q2d, dtype=np.float64).ravel()
    w = np.asarray(dA, dtype=np.float64).ravel()
    ok = ~np.isnan(v)
    order = np.argsort(v[ok], kind="stable")
    vs, cw = v[ok][order], np.cumsum(w[ok][order])
    idx = np.searchsorted(cw, np.asarray(fractions) * cw[-1], side="left")
    return vs[np.minimum(idx, len(vs) - 1)]


# --------------------------------------------------------------------------
# d/dA, Leq2, Keff            xcontour/core.py:463-488, 619-637, 945-966
# --------------------------------------------------------------------------
def cal_gradient_wrt_area(var, area, dtype=np.float32, var_coord=None, area_coord=None):
    """core.py:480-483: ``differentiate('contour')`` is np.gradient (edge_order=1) against each array's own
    'contour' coordinate -- the float coordinate 0..N-1 after cal_contours(int) (default here), the level values
    themselves after cal_contours(array) (core.py:253-264) -- each in its own dtype."""
    var = np.asarray(var)
    area = np.asarray(area)
    coord = contour_coord(var.shape[-1], dtype)
    vc = coord if var_coord is None else np.asarray(var_coord)
    ac = coord if area_coord is None else np.asarray(area_coord)
    return np.gradient(var, vc, axis=-1) / np.gradient(area, ac, axis=-1)


def cal_sqared_equivalent_length(dgrdSdA, dqdA):
    """core.py:635."""
    return dgrdSdA / dqdA ** 2


def cal_normalized_Keff(Leq2, Lmin, mask=1e5):
    """core.py:963-964."""
    with np.errstate(invalid="ignore", divide="ignore"):
        nkeff = Leq2 / Lmin / Lmin
        return np.where(nkeff < mask, nkeff, np.nan)


def latitude_lengths_at(lats):
    """utils.py:518-534."""
    lats = np.asarray(lats)
    return (2.0 * np.pi * Rearth * np.cos(np.deg2rad(lats))).astype(lats.dtype)


def equivalent_latitudes(areas):
    """utils.py:491-515."""
    areas = np.asarray(areas)
    ratio = areas / 2.0 / np.pi / Rearth / Rearth - 1.0
    ratio = np.where(ratio < -1, -1.0, ratio)
    ratio = np.where(ratio > 1, 1.0, ratio)
    return np.rad2deg(np.arcsin(ratio)).astype(areas.dtype)


# --------------------------------------------------------------------------
# local wave activity / local APE       xcontour/core.py:696-799, 802-942
# --------------------------------------------------------------------------
def _lwa_masks(qe, m, increase):
    """core.py:759-766 (variant 1).  ``m`` broadcasts along the eq axis."""
    with np.errstate(invalid="ignore"):
        if increase:
            mask1 = np.where(qe > 0, -1, 0)
            mask2 = np.where(m, 0, mask1)
            mask3 = np.where(np.logical_and(qe < 0, m), 1, mask2)
        else:
            mask1 = np.where(qe < 0, -1, 0)
            mask2 = np.where(m, 0, mask1)
            mask3 = np.where(np.logical_and(qe > 0, m), 1, mask2)
    return mask3


def _select_part(mask3, part, increase):
    """core.py:773-784 -- NaN where the part is not selected."""
    if part == "all":
        return mask3.astype(np.float64)
    keep_pos = (part == "upper") == bool(increase)
    if keep_pos:
        return np.where(mask3 > 0, mask3, np.nan).astype(np.float64)
    return np.where(mask3 < 0, mask3, np.nan).astype(np.float64)


def cal_local_wave_activity(q, Q, dA, coord, increase, part="all",
                            mask_idx=None, rows=None, variant=1):
    """Brute-force loop over every index of the equivalent dimension, as the
    reference does (core.py:752-794; variant 2: core.py:858-900).

    ``q[S, ny, nx]`` with the equivalent dimension on axis 1, ``Q[S, ny]``,
    ``dA[ny, nx]``, ``coord[ny]``.  Hazards settled as in SURVEY.md §8a: the sum
    runs along the equivalent dimension only (H1: intended meaning of
    ``_integrate(..., self.dimEqN)``, core.py:789), and the weight is the literal
    current code ``qe * mask * (dA/dA.max()) * dA`` (H2).  ``rows`` restricts the
    j-loop (used for bounded CPU-baseline timing).
    Returns LWA[S, ny, nx] (fp64), and with ``mask_idx`` also (contours, masks).
    """
    part = part.lower()
    if part not in ("all", "upper", "lower"):
        raise Exception("invalid part, should be in ['all', 'upper', 'lower']")
    q = np.asarray(q)
    Q = np.asarray(Q)
    dA = np.asarray(dA)
    coord = np.asarray(coord)
    S, ny, nx = q.shape
    wei = dA / np.nanmax(dA)                             # core.py:723-724 (xarray .max() skips NaN)
    coord_incre = not (coord[-1] < coord[0])             # core.py:736-738
    if mask_idx is not None and max(mask_idx) >= ny:
        raise Exception("indices in mask_idx out of boundary")
    jj = range(ny) if rows is None else rows
    out = np.zeros((S, ny, nx), dtype=np.float64)
    contours, masks = [], []
    for j in jj:
        m = (coord >= coord[j]) if coord_incre else (coord <= coord[j])
        if variant == 1:
            qe = q - Q[:, j][:, None, None]              # core.py:754, dims (S, eq, x)
            mask3 = _lwa_masks(qe, m[None, :, None], increase)
            mf = _select_part(mask3, part, increase)
            lwa = -np.nansum(qe * mf * wei[None] * dA[None], axis=1)
        else:
            # core.py:860: q.isel(eq=j) - Q broadcasts to dims (S, x, eq) -- the eq axis
            # comes LAST, which is the order the terms are summed in (core.py:865-872, 896)
            qe = q[:, j, :][:, :, None] - Q[:, None, :]
            mask3 = _lwa_masks(qe, m[None, None, :], not increase)
            mf = _select_part(mask3, part, increase)
            lwa = -np.nansum(qe * mf * wei.T[None] * dA.T[None], axis=2)
            mask3 = mask3.transpose(0, 2, 1)
        if mask_idx is not None and j in mask_idx:
            contours.append(Q[:, j].copy())
            masks.append(mask3.copy())
        out[:, j, :] = lwa
    if mask_idx is not None:
        return out, contours, masks
    return out


def cal_local_wave_activity_fast(q, Q, dA, coord, increase, part="all"):
    """Reformulated LWA (variant 1) for monotone ``Q``: each cell (j', i) adds
    ``sign*(v - Q_j)*w`` to one contiguous j-range found by two binary searches,
    so a column costs O(ny log ny) instead of O(ny^2).  This is the algorithm
    the CUDA kernel implements; it is cross-checked here against the brute-force
    loop (no counterpart in the reference)."""
    q = np.asarray(q, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    dA = np.asarray(dA)
    S, ny, nx = q.shape
    # note: the reference's ``m`` (core.py:757) is "j' >= j" in INDEX space for
    # either direction of a strictly monotone coordinate, so ``coord`` drops out.
    # wei = dA/dA.max() is rounded in dA's own dtype (core.py:723-724)
    ww = (dA / np.nanmax(dA)).astype(np.float64) * dA.astype(np.float64)
    sgn = 1.0 if increase else -1.0
    keep_pos = (part == "upper") == bool(increase)
    out = np.zeros((S, ny, nx))
    jidx = np.broadcast_to(np.arange(ny)[:, None], (ny, nx))
    cols = np.broadcast_to(np.arange(nx)[None, :], (ny, nx))
    for s in range(S):
        Qs = sgn * Q[s]            # nondecreasing when the sort is consistent
        v = sgn * q[s]
        lo = np.searchsorted(Qs, v.ravel(), side="left").reshape(ny, nx)
        hi = np.searchsorted(Qs, v.ravel(), side="right").reshape(ny, nx)
        dS = np.zeros((ny + 2, nx))
        dV = np.zeros((ny + 2, nx))
        valid = ~np.isnan(v) & ~np.isnan(ww)
        # mask -1 region: j' < j < lo  (cell value above Q_j, below row j)
        t1 = valid & (lo > jidx + 1)
        # mask +1 region: hi <= j <= j' (cell value below Q_j, at/above row j)
        t2 = valid & (hi <= jidx)
        if part != "all":
            if keep_pos:
                t1 = np.zeros_like(t1)
            else:
                t2 = np.zeros_like(t2)
        for t, a, b, sg in ((t1, jidx + 1, lo, 1.0), (t2, hi, jidx + 1, -1.0)):
            w = sg * ww[t]
            vw = w * v[t]
            np.add.at(dS, (a[t], cols[t]), w)
            np.add.at(dS, (b[t], cols[t]), -w)
            np.add.at(dV, (a[t], cols[t]), vw)
            np.add.at(dV, (b[t], cols[t]), -vw)
        Sx = np.cumsum(dS, axis=0)[:ny]
        Vx = np.cumsum(dV, axis=0)[:ny]
        out[s] = sgn * (Vx - Qs[:, None] * Sx)
    return out


# --------------------------------------------------------------------------
# |grad q|^2 provider -- NOT in the reference
# --------------------------------------------------------------------------
def squared_gradient_latlon(q, lat_deg, lon_deg):
    """Centred finite differences on a regular lat-lon grid, periodic in
    longitude, one-sided at the first/last latitude:
        dq/dx = (q[j,i+1] - q[j,i-1]) * cx[j],  cx = 1 / ((2 dlambda) * (R cos phi_j))
        dq/dy = (q[j+1,i] - q[j-1,i]) * cy[j],  cy = 1 / ((phi_{j+1} - phi_{j-1}) * R)
        |grad q|^2 = (dq/dx)^2 + (dq/dy)^2
    The reference obtains this field from external packages whose source is not
    under /root/reference (xinvert / GeoApps, tests/test_Keff_ocean.py:31-32),
    so its parity is UNPINNED; this is the definition the CUDA stencil follows
    (same operation order, fp64 throughout)."""
    q = np.asarray(q, dtype=np.float64)
    phi = np.deg2rad(np.asarray(lat_deg, dtype=np.float64))
    lam = np.deg2rad(np.asarray(lon_deg, dtype=np.float64))
    dlam = lam[1] - lam[0]
    ny = q.shape[-2]
    jm = np.maximum(np.arange(ny) - 1, 0)
    jp = np.minimum(np.arange(ny) + 1, ny - 1)
    with np.errstate(divide="ignore"):
        cx = 1.0 / ((2.0 * dlam) * (Rearth * np.cos(phi)))
        cy = 1.0 / ((phi[jp] - phi[jm]) * Rearth)
    dqdx = (np.roll(q, -1, axis=-1) - np.roll(q, 1, axis=-1)) * cx[:, None]
    dqdy = (q[..., jp, :] - q[..., jm, :]) * cy[:, None]
    return dqdx * dqdx + dqdy * dqdy


BOUNDARY_PAD = {"periodic": "wrap", "extend": "edge", "reflect": "reflect", "fill": "constant"}


def row_metrics_latlon(lat_deg, lon_deg):
    """(cx, cy) of squared_gradient_latlon, see there."""
    phi = np.deg2rad(np.asarray(lat_deg, dtype=np.float64))
    lam = np.deg2rad(np.asarray(lon_deg, dtype=np.float64))
    ny = phi.shape[0]
    jm = np.maximum(np.arange(ny) - 1, 0)
    jp = np.minimum(np.arange(ny) + 1, ny - 1)
    with np.errstate(divide="ignore"):
        return 1.0 / ((2.0 * (lam[1] - lam[0])) * (Rearth * np.cos(phi))), 1.0 / ((phi[jp] - phi[jm]) * Rearth)


def row_metrics_cartesian(y, x):
    """cx = 1/(2 dx) (uniform x), cy[j] = 1/(y[j+1]-y[j-1]), the interior spacing continued at both ends."""
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    d = np.empty(y.shape[0])
    d[1:-1] = y[2:] - y[:-2]
    d[0] = 2.0 * (y[1] - y[0])
    d[-1] = 2.0 * (y[-1] - y[-2])
    return np.full(y.shape[0], 1.0 / (2.0 * (x[1] - x[0]))), 1.0 / d


def squared_gradient(q, cx, cy, bcx="periodic", bcy="extend", fill=0.0):
    """|grad q|^2 by centred differences with one layer of ghost cells:
        dq/dx = (q[j,i+1] - q[j,i-1]) * cx[j],   dq/dy = (q[j+1,i] - q[j-1,i]) * cy[j]
    ghost cells by numpy.pad: 'periodic' -> wrap, 'extend' -> edge, 'reflect' -> reflect, 'fill' -> constant.
    The reference takes this field from xinvert.FiniteDiff(BCs=...) / GeoApps (tests/test_Keff_ocean.py:26-32,
    tests/test_clength.py:39-45), whose sources are not under /root/reference: parity UNPINNED, the definition
    is ours; squared_gradient_latlon is the (periodic, extend) case with lat-lon metrics."""
    q = np.asarray(q, dtype=np.float64)

    def pad(axis, mode):
        pw = [(0, 0)] * q.ndim
        pw[axis] = (1, 1)
        kw = {"constant_values": float(np.float32(fill))} if mode == "fill" else {}
        return np.pad(q, pw, mode=BOUNDARY_PAD[mode], **kw)
    qx, qy = pad(-1, bcx), pad(-2, bcy)
    dqdx = (qx[..., 2:] - qx[..., :-2]) * np.asarray(cx)[:, None]
    dqdy = (qy[..., 2:, :] - qy[..., :-2, :]) * np.asarray(cy)[:, None]
    return dqdx * dqdx + dqdy * dqdy


# --------------------------------------------------------------------------
# raw reader for the one data file that is present
# --------------------------------------------------------------------------
def read_barotropic_vorticity(path):
    """Data/barotropic_vorticity.nc is NetCDF-4/HDF5 with contiguous,
    unfiltered little-endian fp32 datasets at fixed byte offsets."""
    buf = open(path, "rb").read()
    lat = np.frombuffer(buf, "<f4", 256, 885).copy()
    lon = np.frombuffer(buf, "<f4", 512, 1909).copy()
    q = np.frombuffer(buf, "<f4", 256 * 512, 10101).reshape(256, 512).copy()
    return lat, lon, q


def latlon_cell_area(lat_deg, lon_deg):
    """Spherical cell areas for a lat-lon grid with cell edges midway between
    grid latitudes, clipped at the poles:
        dA[j] = R^2 (sin(phi_{j+1/2}) - sin(phi_{j-1/2})) dlambda."""
    lat = np.asarray(lat_deg, dtype=np.float64)
    lon = np.asarray(lon_deg, dtype=np.float64)
    asc = lat[-1] > lat[0]
    la = lat if asc else lat[::-1]
    edges = np.empty(len(la) + 1)
    edges[1:-1] = 0.5 * (la[1:] + la[:-1])
    edges[0] = max(-90.0, la[0] - 0.5 * (la[1] - la[0]))
    edges[-1] = min(90.0, la[-1] + 0.5 * (la[-1] - la[-2]))
    band = Rearth ** 2 * np.diff(np.sin(np.deg2rad(edges)))
    if not asc:
        band = band[::-1]
    dlam = np.deg2rad(abs(lon[1] - lon[0]))
    return np.repeat((band * dlam)[:, None], len(lon), axis=1)
