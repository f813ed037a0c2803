"""
Property tests (hypothesis, derandomized: the same examples on every run) of the oracle itself -- CPU only.  They guard the
restatement against slips that the fixed-size tests could miss: random shapes,
NaN cells, ties with the levels, both directions and both comparison senses.
"""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import xcontour_oracle as O


def _field(seed, ny, nx, nan_frac):
    rng = np.random.default_rng(seed)
    y = np.linspace(-1, 1, ny)[:, None]
    q = (y + 0.3 * np.sin(3 * np.linspace(0, 6.28, nx))[None, :] + 0.05 * rng.standard_normal((ny, nx))).astype(np.float32)
    if nan_frac:
        q[rng.random((ny, nx)) < nan_frac] = np.nan
    dA = (0.5 + rng.random((ny, nx))).astype(np.float32)
    return q, dA


@settings(max_examples=40, deadline=None, derandomize=True)
@given(seed=st.integers(0, 10**6), ny=st.integers(4, 24), nx=st.integers(3, 30), N=st.integers(2, 40),
       increase=st.booleans(), lt=st.booleans(), nan_frac=st.sampled_from([0.0, 0.1]))
def test_hist_cdf_invariants(seed, ny, nx, N, increase, lt, nan_frac):
    q, dA = _field(seed, ny, nx, nan_frac)
    ctr = O.cal_contours(q[None], N, increase)
    if not np.diff(ctr[0]).all():
        return                                              # degenerate: the reference raises
    a = O.cal_integral_within_contours_hist(q[None], ctr[0], dA, lt)[0]
    d = np.diff(a)
    # monotone in the contour index, direction fixed by (increase, lt)
    assert np.all(d >= 0) or np.all(d <= 0)
    assert (a[-1] >= a[0]) == (increase == lt) or a[-1] == a[0]
    # never more than the valid area, never negative
    tot = np.nansum(np.where(np.isnan(q), np.nan, dA).astype(np.float64))
    assert a.min() >= 0 and a.max() <= tot * (1 + 1e-12)
    # strict path differs only by cells sitting exactly on a level / the extreme cell
    s = O.cal_integral_within_contours(q[None], ctr[0], dA, lt)[0].astype(np.float64)
    on_level = np.isin(q, ctr[0])
    slack = np.nansum(np.where(on_level, dA, 0.0)) + 1e-4 * tot + dA.max()
    assert np.abs(a - s).max() <= slack


@settings(max_examples=25, deadline=None, derandomize=True)
@given(seed=st.integers(0, 10**6), ny=st.integers(5, 40), nx=st.integers(2, 12),
       increase=st.booleans(), part=st.sampled_from(["all", "upper", "lower"]), ties=st.booleans())
def test_lwa_reformulation_equals_reference_loop(seed, ny, nx, increase, part, ties):
    rng = np.random.default_rng(seed)
    q, dA = _field(seed, ny, nx, 0.05)
    Q = np.sort(rng.standard_normal(ny))[None]
    if ties:                                                # flat stretches and exact hits
        Q[0, ny // 3: ny // 3 + 3] = Q[0, ny // 3]
        q[ny // 2, :] = Q[0, ny // 3]
    if not increase:
        Q = Q[:, ::-1].copy()
    coord = np.arange(ny, dtype=np.float64)
    brute = O.cal_local_wave_activity(q[None], Q, dA, coord, increase, part)
    fast = O.cal_local_wave_activity_fast(q[None], Q, dA, coord, increase, part)
    scale = max(np.abs(brute).max(), 1e-300)
    # the reformulation evaluates V_j - Q_j S_j: where a part is (almost) empty the field is ~0 and what is
    # left is the rounding of the two sums, bounded by the size of their terms, not by the field maximum
    ww = (dA / np.nanmax(dA)).astype(np.float64) * dA
    floor = 8 * ny * np.finfo(np.float64).eps * np.nanmax(ww) * max(np.nanmax(np.abs(q)), np.abs(Q).max())
    assert np.abs(brute - fast).max() <= 1e-11 * scale + floor
    if part == "all":
        assert (brute >= -1e-12 * scale).all() if increase else (brute <= 1e-12 * scale).all()


@settings(max_examples=30, deadline=None, derandomize=True)
@given(seed=st.integers(0, 10**6), n=st.integers(2, 30), dtype=st.sampled_from([np.float32, np.float64]),
       time_branch=st.booleans(), decreasing=st.booleans())
def test_hist_edges_structure(seed, n, dtype, time_branch, decreasing):
    rng = np.random.default_rng(seed)
    lo, hi = np.sort(rng.standard_normal(2) * 10.0 ** rng.integers(-4, 3))
    if lo == hi:
        return
    ctr = np.linspace(lo, hi, n).astype(dtype)
    if not np.diff(ctr).all():
        return
    if decreasing:
        ctr = ctr[::-1].copy()
    e, binc = O.hist_edges(ctr, time_branch)
    assert binc == (not decreasing) and e.shape == (n + 1,)
    assert e.dtype == (np.float64 if time_branch else dtype)
    assert np.all(np.diff(e) > 0)                                       # ascending
    assert np.array_equal(e[1:], np.sort(ctr).astype(e.dtype))          # the levels themselves
    assert e[0] < e[1]                                                  # one extra bin below
