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


HKX_NW, HKX_WBITS, HKX_MARGIN = 6, 24, 12


def windowed_sum_model(terms):
    """Python-integer model of hkx_add / hkx_window for one bin of one CTA."""
    exps = [math.frexp(x)[1] + 1022 for x in terms[:64] if x > 0 and math.isfinite(x)]   # biased exponents of the sample
    e_base = (min(exps) if exps else 1023 - 72) - HKX_MARGIN
    acc, esc = [0] * HKX_NW, 0.0
    for x in terms:
        if x == 0:
            continue
        m, e = math.frexp(x)                                          # x = m 2^e, 0.5 <= m < 1
        ex = e + 1022
        rel = ex - e_base
        if x < 0 or not math.isfinite(x) or ex <= 0 or rel >= HKX_NW * HKX_WBITS:
            esc += x
            continue
        mi = int(m * (1 << 53))                                       # 53-bit mantissa, x = mi 2^(ex-1075)
        if rel < 0:
            mi, rel = (mi >> -rel if rel > -53 else 0), 0
        w, sh = rel // HKX_WBITS, rel % HKX_WBITS
        acc[w] += mi << sh
        assert acc[w] < 1 << 96
    tot = 0.0
    for w in range(HKX_NW):
        tot += math.ldexp(float(acc[w]), e_base + w * HKX_WBITS - 1075)
    return tot + esc


def test_windowed_accumulators_are_exact_over_120_binary_orders():
    rng = np.random.default_rng(3)
    for trial in range(20):
        n = int(rng.integers(1, 4000))
        x = np.abs(rng.standard_normal(n)) * 2.0 ** rng.integers(-20, 20, n)
        if trial % 3 == 0:                                            # polar rows: a few terms 2^100 above the rest
            x[rng.integers(0, n, 3)] *= 2.0 ** 100
        if trial % 4 == 0:
            x[rng.integers(0, n, 5)] = 0.0
        exact = math.fsum(x)
        got = windowed_sum_model([float(v) for v in x])
        assert abs(got - exact) <= 4e-16 * exact
    # terms far below the anchor lose only what lies under 2^-52 of the base scale
    x = [1.0] * 10 + [2.0 ** -30 * (1 + 2.0 ** -40)] * 1000
    assert abs(windowed_sum_model(x) - math.fsum(x)) <= 1e-15 * math.fsum(x)
    # negative, infinite and huge terms take the side table
    assert windowed_sum_model([1.0, -0.25, 2.0 ** 300]) == 1.0 - 0.25 + 2.0 ** 300
    assert math.isinf(windowed_sum_model([1.0, math.inf]))


def test_lean_term_decomposition_equals_the_default(tmp_path):
    """xcontour_b200/csrc/hkx_decompose.cuh compiled for the CPU: the funnel-shift decomposition of the
    prepared -DXC_HKX_LEAN=1 build returns the same (window, 96-bit value) as the default statement for
    20 million terms (random bit patterns, exponents around the window range, anchors over the whole
    exponent range)."""
    import os, shutil, subprocess
    if shutil.which("g++") is None:
        pytest.skip("no host compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "hkx_decompose_test")
    subprocess.check_call(["g++", "-O2", "-o", exe, os.path.join(root, "tests", "host", "hkx_decompose_test.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and " 0 mismatches" in r.stdout, r.stdout[-500:]


def test_row_count_area_model():
    """Model of the prepared -DXC_HKX_ROWCNT=1 build: on a grid whose dA is constant along a row the area of
    a bin is sum_rows dA[row] * (cells of that row in the bin); the counts are packed two per 32-bit word
    (u16 halves, one ATOMS.ADD per cell) and every product dA * count is exact in fp64."""
    rng = np.random.default_rng(11)
    ny, nx, N = 37, 1440, 361
    dA_row = (np.cos(np.deg2rad(np.linspace(-89.9, 89.9, ny))) * 7.7e8).astype(np.float32)
    bins = rng.integers(-1, N, size=(ny, nx))                        # -1: outside / NaN
    bins[5, :] = 17                                                  # a whole row in one bin: the largest count
    words = np.zeros((ny, (N + 1) >> 1), dtype=np.uint32)
    for j in range(ny):
        for b in bins[j][bins[j] >= 0]:
            words[j, b >> 1] += np.uint32(1) << np.uint32((b & 1) << 4)
    assert ((words & 0xffff) <= nx).all() and ((words >> 16) <= nx).all()      # no carry between the halves
    area = np.zeros(N)
    for n in range(N):
        t = 0.0
        for j in range(ny):
            cn = int(words[j, n >> 1] >> ((n & 1) << 4)) & 0xffff
            if cn:
                assert float(dA_row[j]) * cn == float(np.float64(dA_row[j]) * np.float64(cn))    # 24 + 11 bits: exact
                t += float(dA_row[j]) * cn
        area[n] = t
    ref = np.bincount(bins[bins >= 0], weights=np.broadcast_to(dA_row[:, None], bins.shape)[bins >= 0].astype(np.float64),
                      minlength=N)
    assert np.abs(area - ref).max() <= 1e-14 * ref.max()
