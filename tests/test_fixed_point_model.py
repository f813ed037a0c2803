"""
CPU models (Python integers) of the two exact-accumulation schemes the CUDA kernels
use, checked against the oracle -- so that the arithmetic design is covered by the
`-m "not gpu"` suite as well:

* k_lwa_fx (xcontour_b200/csrc/lwa.cu): per-column difference arrays of 64-bit
  fixed-point terms X_S = rn(w 2^kS), X_V = rn(w (v - c) 2^kV), one -X deposit at the
  far end of each cell's range, the +X at slot j'+1 re-derived in the prefix walk;
* k_hist_keff<true> (hist_keff.cu): exponent-windowed 96-bit accumulators.
"""
import math
import numpy as np
import pytest

from oracle import xcontour_oracle as O


def lwa_fixed_point_model(q3, Q, dA, increase, part, own_in_planes=False):
    """own_in_planes=True models the -DXC_FX_OWN=1 build: the +X deposit goes to slot j'+1 of
    the planes in the scatter phase (inactive cells deposit nothing) and the walk is a plain
    inclusive prefix."""
    S, ny, nx = q3.shape
    sg = 1.0 if increase else -1.0
    dAmax = np.nanmax(dA)
    ww = (dA / dAmax).astype(np.float64) * dA.astype(np.float64)
    out = np.zeros((S, ny, nx))
    keep_pos = (part == "upper") == increase
    use_t1, use_t2 = part == "all" or not keep_pos, part == "all" or keep_pos
    wm = float(np.nanmax(np.abs(ww)))
    for s in range(S):
        Qs = sg * Q[s]
        a, b = sg * float(np.nanmin(q3[s])), sg * float(np.nanmax(q3[s]))     # fp64 like the kernel
        vlo, vhi = min(a, b), max(a, b)
        c = 0.5 * vlo + 0.5 * vhi
        MV, MS = wm * max(vhi - c, c - vlo) * 1.0000001, wm
        hb = 1
        while (1 << hb) < ny + 1:
            hb += 1
        kb = min(51, 62 - hb) - 1
        kS = kb - math.frexp(MS)[1] + 1 if MS > 0 else 0           # kb - ilogb(M)
        kV = kb - math.frexp(MV)[1] + 1 if MV > 0 else 0
        for i in range(nx):
            far_S, far_V = [0] * (ny + 2), [0] * (ny + 2)
            own = []
            for jp in range(ny):
                v, w = sg * float(q3[s, jp, i]), float(ww[jp, i])
                if v != v or w != w:
                    own.append((0, 0))
                    continue
                XS, XV = int(np.rint(w * 2.0 ** kS)), int(np.rint((w * (v - c)) * 2.0 ** kV))
                assert abs(XS) <= 1 << 51 and abs(XV) <= 1 << 51      # range of the magic-number rounding
                lo, hi = int(np.searchsorted(Qs, v, "left")), int(np.searchsorted(Qs, v, "right"))
                t = jp + 1                                            # inactive: cancels the own deposit
                if lo > jp + 1:
                    t = lo if use_t1 else t
                elif hi <= jp and use_t2:
                    t = hi
                if own_in_planes:
                    own.append((0, 0))
                    if t == jp + 1:
                        continue
                    far_S[jp + 1] += XS
                    far_V[jp + 1] += XV
                else:
                    own.append((XS, XV))
                far_S[t] -= XS
                far_V[t] -= XV
            RS = RV = 0
            for j in range(ny):
                RS += far_S[j]
                RV += far_V[j]
                assert abs(RS) < 1 << 63 and abs(RV) < 1 << 63
                out[s, j, i] = sg * (RV * 2.0 ** -kV - (Qs[j] - c) * (RS * 2.0 ** -kS))
                RS += own[j][0]
                RV += own[j][1]
    return out


@pytest.mark.parametrize("increase", [True, False])
@pytest.mark.parametrize("part", ["all", "upper", "lower"])
def test_lwa_fixed_point_model_matches_reference_loop(increase, part):
    rng = np.random.default_rng(5)
    ny, nx = 61, 9
    y = np.linspace(-1, 1, ny)[:, None]
    q = y + 0.4 * np.sin(np.linspace(0, 12.56, nx))[None, :] * (1 - y ** 2) + 0.03 * rng.standard_normal((ny, nx))
    q = (np.round(q * 16) / 16).astype(np.float32)                   # homogenised patches, exact hits
    q3 = np.stack([q, q[::-1] + 300.0])                              # second slice: large offset (cancellation)
    dA = (0.5 + rng.random((ny, nx))).astype(np.float64)
    dA[11, 7] = np.nan
    Q = np.stack([np.sort(rng.choice(np.unique(q3[k]), size=ny).astype(np.float64)) for k in range(2)])
    if not increase:
        Q = Q[:, ::-1].copy()
    q3[1, 40:44, 2:5] = np.nan
    coord = np.arange(ny, dtype=np.float64)
    ref = O.cal_local_wave_activity(q3, Q, dA, coord, increase, part)
    out = lwa_fixed_point_model(q3, Q, dA, increase, part)
    for s in range(2):
        assert np.abs(out[s] - ref[s]).max() <= 1e-12 * np.abs(ref[s]).max()
    # the prepared variant (own-slot deposits in the planes) sums the same integers
    assert np.array_equal(lwa_fixed_point_model(q3, Q, dA, increase, part, own_in_planes=True), out)


# --------------------------------------------------------------------------- row-march binning kernel (bin_rows.cu)
def add96_words(G):
    """br_add96: the three 32-bit words of floor(G), 0 <= G < 2^84, from two round-to-zero magic additions."""
    assert 0.0 <= G < 2.0 ** 84
    hi = int(G // 2.0 ** 32)                       # exact: the quotient is < 2^52
    t = 2.0 ** 84 + hi * 2.0 ** 32                 # what __dadd_rz(G, 0x1p84) returns: its mantissa holds floor(G / 2^32)
    assert hi < 2 ** 52 and float(int(t)) == t      # representable: the round-to-zero sum IS this value
    r = G - (t - 2.0 ** 84)                        # exact, in [0, 2^32)
    assert 0.0 <= r < 2.0 ** 32
    return int(r) & 0xffffffff, hi & 0xffffffff, (hi >> 32) & 0xfffff


def bin_rows_sum_model(terms, hbits=17):
    """Python-integer model of one (bin, copy) accumulator of k_bin_rows: scale so that the bound of the terms
    stays below 2^(96-h), add floor(G) word by word with the carries the kernel derives from the values the
    atomics return, convert back.  Returns (sum, scale exponent)."""
    bound = max(terms) * 1.000001
    e = math.frexp(bound)[1]                       # bound < 2^e
    k = 96 - hbits - e
    k = (k if k >= 0 else k - 1) // 2 * 2          # even: folded into the squared row metrics
    w = [0, 0, 0]
    for x in terms:
        G = math.ldexp(x, k)
        assert G < 2.0 ** (96 - hbits)
        v0, v1, v2 = add96_words(G)
        o0 = w[0]; w[0] = (w[0] + v0) & 0xffffffff
        c0 = (o0 + v0) >> 32
        t1 = (v1 + c0) & 0xffffffff; k1 = (v1 + c0) >> 32
        o1 = w[1]; w[1] = (w[1] + t1) & 0xffffffff
        c1 = (o1 + t1) >> 32
        w[2] = (w[2] + v2 + k1 + c1) & 0xffffffff
    total = (w[2] << 64) | (w[1] << 32) | w[0]
    assert total == sum(int(math.ldexp(x, k)) for x in terms)       # the words ARE the exact integer sum
    return math.ldexp(float(total), -k), k


def test_bin_rows_accumulator_is_the_exact_sum_of_truncated_terms():
    rng = np.random.default_rng(3)
    for trial in range(20):
        n = int(rng.integers(1, 4000))
        x = (np.abs(rng.standard_normal(n)) * 2.0 ** rng.integers(-24, 1, n)).tolist()
        if trial % 4 == 0:
            for i in rng.integers(0, n, 5):
                x[i] = 0.0
        x.append(8.0)                                               # the bound the scale is derived from
        got, k = bin_rows_sum_model(x)
        exact = math.fsum(x)
        # each term loses less than 2^-k: 2^-78 of the bound with h = 17
        assert 0.0 <= exact - got <= len(x) * 2.0 ** -k + 4e-16 * exact
        assert abs(got - exact) <= 1e-15 * exact
    # gradients 2^-13 of the largest representable one (terms 2^-26 of the bound) keep full fp64 precision,
    # fp32-ulp sized differences (terms 2^-46 of the bound) 1e-9 of their own value
    for d, tol in ((26, 2e-16), (46, 1e-9)):
        x = [1.0] + [2.0 ** -d * (1 + i * 2.0 ** -30) for i in range(1000)]
        got, _ = bin_rows_sum_model(x)
        small = math.fsum(x[1:])
        assert abs((got - 1.0) - small) <= tol * small + 2.0 ** -52


def test_row_count_area_model():
    """k_bin_rows: on a grid whose dA is constant along a row the area of a bin is sum_rows dA[row] * (cells of
    that row in the bin); the counts of two rows share a 32-bit word (u16 halves selected by the row's parity,
    one shared-memory atomic per cell) and every product dA * count is exact in fp64."""
    rng = np.random.default_rng(11)
    ny, nx, N = 38, 1440, 361
    dA_row = (np.cos(np.deg2rad(np.linspace(-89.9, 89.9, ny))) * 7.7e8).astype(np.float32)
    bins = rng.integers(-1, N, size=(ny, nx))                        # -1: outside / NaN
    bins[5, :] = 17                                                  # a whole row in one bin: the largest count
    words = np.zeros((ny // 2, N), dtype=np.uint32)
    for j in range(ny):
        for b in bins[j][bins[j] >= 0]:
            words[j >> 1, b] += np.uint32(1) << np.uint32((j & 1) << 4)
    assert ((words & 0xffff) <= nx).all() and ((words >> 16) <= nx).all()      # no carry between the halves
    area = np.zeros(N)
    for n in range(N):
        t = 0.0
        for j in range(ny):
            cn = int(words[j >> 1, n] >> ((j & 1) << 4)) & 0xffff
            if cn:
                assert float(dA_row[j]) * cn == float(np.float64(dA_row[j]) * np.float64(cn))    # 24 + 11 bits: exact
                t += float(dA_row[j]) * cn
        area[n] = t
    ref = np.bincount(bins[bins >= 0], weights=np.broadcast_to(dA_row[:, None], bins.shape)[bins >= 0].astype(np.float64),
                      minlength=N)
    assert np.abs(area - ref).max() <= 1e-14 * ref.max()
