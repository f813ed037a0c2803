// Dedicated binning kernel of the fused Keff pass: fp32 tracer, fp32 cell areas,
// uniform (linspace) edges, two accumulators per bin -- {dA, |grad q|^2 * dA} with
// the gradient stencil evaluated in flight.  Same arithmetic, bin rule and
// warp-private fp64 histograms as the general kernel in hist.cu (the parity tests
// compare both against the oracle); everything that kernel decides at run time
// (dtype switches, optional integrands, masks, bin-index output, edge search mode)
// is fixed here, which halves the instruction count of the hot loop.
#include "common.cuh"
#include "internal.h"
#include "grad2.cuh"
#include "hkx_decompose.cuh"
#include <math_constants.h>

namespace xc {

constexpr int HK_WARPS = 16;

struct HistKeffParams {
    const float* q; int P; int per; long s0;
    const double* edges; int N;              // [S][N+1] ascending, uniform
    const float* dA;
    int ny, nx; const double* cx; const double* cy;
    double* part;                            // [S][C][2][N]
};

// bin p with e[p] <= v < e[p+1], -1 outside / NaN.  t = (v - e0)/h is evaluated in
// fp32; `delta` bounds |t_fp32 - t_exact| plus the deviation of the true edges from
// the linear model (computed per slice from the actual edge values), so whenever
// frac(t) is at least delta away from 0 and 1 the floor IS the bin and no edge has
// to be read.  Otherwise (about one cell in a thousand) the true fp64 edges in
// shared memory decide, exactly as in the general kernel.
__device__ __forceinline__ int hk_find_bin(float vf, const double* e, int N, float basef, float invf,
                                           float delta)
{
    if (vf != vf) return -1;
    const float t = (vf - basef) * invf;
    const float fl = floorf(t);
    if (t - fl >= delta && (fl + 1.0f) - t >= delta && fl >= 0.0f && fl < (float)N)
        return (int)fl;
    const double v = (double)vf;
    int p = __float2int_rd(fminf(fmaxf(t, 0.0f), (float)(N - 1)));
    for (;;) {
        const double lo = e[p], hi = e[p + 1];
        const int d = (v >= hi) - (v < lo);
        if (d == 0) return p;
        p += d;
        if ((unsigned)p >= (unsigned)N) return -1;
    }
}

#ifndef XC_HK_DEDUP      /* 0: MATCH.ANY, peel or pointer jumping by duplicate count; 1: four MATCH.ANY up front; 2: byte tags; 4: peel only */
#define XC_HK_DEDUP 0
#endif

// one scatter of (w0, w1) into H[bin] for the lanes with bin >= 0 (see hist.cu)
__device__ __forceinline__ void hk_peel(double2* H, unsigned pr, int bin, double w0, double w1, int lane)
{
    do {
        if (pr && (__ffs(pr) - 1) == lane) {
            double2 t = H[bin]; t.x += w0; t.y += w1; H[bin] = t;
        }
        pr &= pr - 1u;
        __syncwarp();
    } while (__any_sync(XC_FULL, pr != 0u));
}
__device__ __forceinline__ unsigned hk_match(int bin, int lane)
{
    const bool a = bin >= 0;
    unsigned pr = __match_any_sync(XC_FULL, a ? (unsigned)bin : (0x80000000u | (unsigned)lane));
    return a ? pr : 0u;
}
// Heavy duplication (smooth fields: many lanes of a warp step share a bin) would
// make the peel take one round per duplicate.  hk_jump combines the lanes that share
// a bin BEFORE touching shared memory: the peers of a bin form a linked list in lane
// order; pointer jumping (each lane adds the value of its successor and takes over
// the successor's successor) leaves the group total in the lowest lane after
// ceil(log2(group size)) shuffle steps, and that lane alone does one conflict-free
// read-modify-write.  hk_scatter picks per warp step: peel for <= 3 duplicates,
// pointer jumping beyond.
__device__ __forceinline__ void hk_scatter(double2* H, int bin, double w0, double w1, int lane)
{
    const bool a = bin >= 0;
    unsigned pr = hk_match(bin, lane);
    const int cnt = __popc(pr);
    if (__reduce_max_sync(XC_FULL, cnt) <= 3) { hk_peel(H, pr, bin, w0, w1, lane); return; }
    const unsigned above = (lane == 31) ? 0u : (pr & (0xffffffffu << (lane + 1)));
    int nxt = above ? (__ffs(above) - 1) : -1;
    const bool leader = a && ((__ffs(pr) - 1) == lane);
    while (__any_sync(XC_FULL, nxt >= 0)) {
        const int src = nxt >= 0 ? nxt : lane;
        const double g0 = __shfl_sync(XC_FULL, w0, src), g1 = __shfl_sync(XC_FULL, w1, src);
        const int gn = __shfl_sync(XC_FULL, nxt, src);
        if (nxt >= 0) { w0 += g0; w1 += g1; nxt = gn; }
    }
    if (leader) { double2 t = H[bin]; t.x += w0; t.y += w1; H[bin] = t; }
    __syncwarp();
}
__device__ __forceinline__ void hk_tag(double2* H, uint8_t* tag, int bin, double w0, double w1, int lane)
{
    bool a = bin >= 0;
    unsigned pending = __ballot_sync(XC_FULL, a);
    while (pending) {
        if (a) tag_store(tag + bin, (unsigned)lane);
        __syncwarp();
        if (a && tag_load(tag + bin) == (unsigned)lane) {
            double2 t = H[bin]; t.x += w0; t.y += w1; H[bin] = t;
            a = false;
        }
        __syncwarp();
        pending = __ballot_sync(XC_FULL, a);
    }
}

#ifndef XC_HKX_MATCH     /* 0: always direct adds; 1: register pre-combination in crowded warp steps; 2: in every step */
#define XC_HKX_MATCH 1
#endif
#ifndef XC_HKX_CROWD     /* neighbouring-lane bin matches (of 31) that make a warp step `crowded` */
#define XC_HKX_CROWD 8
#endif

// ---------------------------------------------------------------------------
// Exact accumulation on native shared-memory integer atomics (default).
// Measured on B200 (scripts/micro/atoms_bench.cu): ATOMS.ADD.U32 costs ~3 cycles per
// warp instruction even with bank conflicts and colliding addresses, one
// conflict-free fp64-pair read-modify-write round 15-18 (and the lanes of a warp that
// share a bin need one round each).  A term x = m * 2^(ex-1075) (m: 53-bit mantissa)
// is added as the integer m << sh into a 96-bit accumulator of the window
// w = (ex - e_base) / 24, sh = (ex - e_base) % 24: inside a window every add is exact
// and order-independent, HKX_NW windows cover 144 binary orders of magnitude (the
// polar rows of a lat-lon grid put |grad q|^2 dA 2^108 above the mid-latitude
// values), and a CTA's <= 2^18 terms per bin cannot overflow 96 bits (77 + 18).
// e_base comes from the smallest term of the CTA's first 2048 cells; terms below it
// are truncated at 2^-52 of the base scale, terms above the top window, negative or
// non-finite ones take an fp64 CAS add into a side table (never on sane data).  All
// warps of a CTA share ONE set of accumulators, so there is no lane election at all;
// only in warp steps where neighbouring lanes sit in the same bin (smooth stretches,
// where the atomic unit would serialise the colliding lanes) the lanes of a bin are
// first combined in registers as in hk_scatter.
constexpr int HKX_NW = 6;
constexpr int HKX_WBITS = 24;
constexpr int HKX_MARGIN = 12;           // windows start this many bits below the smallest sampled term

#ifndef XC_HKX_EXP           /* timing-split build only (results wrong by construction): 1 = bins, stencil and
                                weights are computed but nothing is accumulated (scripts/lwa_split.sh) */
#define XC_HKX_EXP 0
#endif
#ifndef XC_HKX_ROWCNT        /* 1: the AREA of every cell whose dA equals the first dA of its row (all of them on a lat-lon
                                grid) is accumulated as an exact integer cell count per (row, bin) -- one ATOMS.ADD --
                                and multiplied by that dA once per row at the end; other cells keep the windowed add.
                                hkx_add is 55 % of the kernel's instructions and the area is half of it.  A/B switch,
                                not yet timed (scripts/ab_round2.sh). */
#define XC_HKX_ROWCNT 0
#endif
#ifndef XC_HKX_LEAN          /* 1: same sums from fewer instructions -- 32-bit funnel shifts instead of 64-bit variable
                                shifts (hkx_decompose.cuh, checked on the CPU against the default statement), the
                                below-anchor truncation out of line, carries from 64-bit adds, no division in the cell
                                loop.  The ncu source page of the default build puts 55 % of the kernel's instructions
                                in hkx_add, 8 % of them in the (never taken) truncation alone.  Candidate for round 2,
                                not yet timed (scripts/ab_round2.sh). */
#define XC_HKX_LEAN 0
#endif
#if XC_HKX_LEAN
__device__ __noinline__ unsigned long long hkx_shift_below(unsigned long long m, int rel)
{
    return rel > -53 ? (m >> (-rel)) : 0ull;
}
__device__ __forceinline__ void hkx_add(uint32_t* acc, double* esc, int N, int bin, int e_base, double x)
{
    const int hi = __double2hiint(x);
    const uint32_t lo = (uint32_t)__double2loint(x);
    if ((uint32_t)hi - 0x00100000u >= hkx_accept_span(e_base, HKX_NW, HKX_WBITS)) {   // one comparison, see hkx_decompose.cuh
        if (x != 0.0) atomicAdd(esc + bin, x);
        return;
    }
    int rel = (hi >> 20) - e_base;
    uint32_t mh = (uint32_t)((hi & 0xfffff) | 0x100000), ml = lo;
    if (rel < 0) {                                           // below the anchor: a real branch, never taken on sane data
        const unsigned long long r = hkx_shift_below(((unsigned long long)mh << 32) | ml, rel);
        mh = (uint32_t)(r >> 32); ml = (uint32_t)r; rel = 0;
    }
    const int w = (rel * 2731) >> 16;
    const int sh = rel - w * HKX_WBITS;
    const uint32_t v0 = ml << sh, v1 = __funnelshift_l(ml, mh, sh), v2 = __funnelshift_l(mh, 0u, sh);
    uint32_t* s = acc + ((size_t)w * N + bin) * 3;
    const uint32_t o0 = atomicAdd(s, v0);
    const unsigned long long t0 = (unsigned long long)o0 + v0;                   // bit 32: carry out of word 0
    const unsigned long long t1 = (unsigned long long)v1 + (uint32_t)(t0 >> 32);
    const uint32_t o1 = atomicAdd(s + 1, (uint32_t)t1);
    const unsigned long long u1 = (unsigned long long)o1 + (uint32_t)t1;
    atomicAdd(s + 2, v2 + (uint32_t)(t1 >> 32) + (uint32_t)(u1 >> 32));
}
#else
__device__ __forceinline__ void hkx_add(uint32_t* acc, double* esc, int N, int bin, int e_base, double x)
{
#if XC_HKX_EXP == 1
    if (x == 1.2345e-300) atomicAdd(esc + bin, x);               // keeps bin and x alive, never taken
    return;
#endif
    const int hi = __double2hiint(x);
    const uint32_t lo = (uint32_t)__double2loint(x);
    const int ex = (hi >> 20) & 0x7ff;
    int rel = ex - e_base;
    if (hi < 0 || ex == 0x7ff || ex == 0 || rel >= HKX_NW * HKX_WBITS) {
        if (x != 0.0) atomicAdd(esc + bin, x);               // negative, non-finite, denormal, above the top window
        return;
    }
    unsigned long long m = ((unsigned long long)(uint32_t)((hi & 0xfffff) | 0x100000) << 32) | lo;
    if (rel < 0) { m = rel > -53 ? (m >> (-rel)) : 0ull; rel = 0; }
    const int w = (rel * 2731) >> 16;                        // rel / 24 for rel < 8192
    const int sh = rel - w * HKX_WBITS;
    const unsigned long long v = m << sh;
    const uint32_t v0 = (uint32_t)v, v1 = (uint32_t)(v >> 32);
    const uint32_t v2 = sh > 11 ? (uint32_t)(m >> (64 - sh)) : 0u;
    uint32_t* s = acc + ((size_t)w * N + bin) * 3;
    const uint32_t o0 = atomicAdd(s, v0);
    const uint32_t c0 = (uint32_t)((o0 + v0) < v0);
    const uint32_t t1 = v1 + c0;
    const uint32_t o1 = atomicAdd(s + 1, t1);
    const uint32_t c1 = (uint32_t)(t1 < c0) + (uint32_t)((o1 + t1) < t1);
    atomicAdd(s + 2, v2 + c1);
}
#endif
// exponent field of a term that can anchor the windows (positive, finite, normal), else INT_MAX
__device__ __forceinline__ int hkx_exp(double x)
{
    const int hi = __double2hiint(x);
    const int ex = (hi >> 20) & 0x7ff;
    return (hi > 0 && ex != 0x7ff && ex != 0) ? ex : 0x7fffffff;
}
// value of a window's 96-bit accumulator times 2^(unit exponent)
__device__ __forceinline__ double hkx_window(const uint32_t* s, int unit_exp)
{
    const double d = fma((double)s[2], 18446744073709551616.0, fma((double)s[1], 4294967296.0, (double)s[0]));
    return ldexp(d, unit_exp);
}
// one scatter step of the fixed-point path: lanes that share a bin with more than 3
// others are combined in registers first (pointer jumping over the MATCH.ANY peer
// list, fixed lane order), everybody else adds directly
__device__ __forceinline__ void hkx_scatter(uint32_t* accA, uint32_t* accG, double* escA, double* escG, int N,
                                            int eA, int eG, int bin, double w0, double w1, int lane, bool crowded)
{
#if XC_HKX_MATCH
    if (XC_HKX_MATCH == 2 || crowded) {                          // warp-uniform
        const bool a = bin >= 0;
        const unsigned pr = hk_match(bin, lane);
        const unsigned above = (lane == 31) ? 0u : (pr & (0xffffffffu << (lane + 1)));
        int nxt = above ? (__ffs(above) - 1) : -1;
        const bool leader = a && ((__ffs(pr) - 1) == lane);
        while (__any_sync(XC_FULL, nxt >= 0)) {
            const int src = nxt >= 0 ? nxt : lane;
            const double g0 = __shfl_sync(XC_FULL, w0, src), g1 = __shfl_sync(XC_FULL, w1, src);
            const int gn = __shfl_sync(XC_FULL, nxt, src);
            if (nxt >= 0) { w0 += g0; w1 += g1; nxt = gn; }
        }
#if XC_HKX_ROWCNT
        if (leader) hkx_add(accG, escG, N, bin, eG, w1);         // the area went to the row counts
#else
        if (leader) { hkx_add(accA, escA, N, bin, eA, w0); hkx_add(accG, escG, N, bin, eG, w1); }
#endif
        return;
    }
#endif
#if XC_HKX_ROWCNT
    if (bin >= 0) hkx_add(accG, escG, N, bin, eG, w1);
#else
    if (bin >= 0) { hkx_add(accA, escA, N, bin, eA, w0); hkx_add(accG, escG, N, bin, eG, w1); }
#endif
}

// grid = (C, nslices), block = 16 warps, 2 CTAs per SM.
template <bool FX>
__global__ void __launch_bounds__(HK_WARPS * 32, 2)
k_hist_keff(const HistKeffParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int N = p.N;
    double*  e = reinterpret_cast<double*>(smem);
    double2* H = reinterpret_cast<double2*>(e + ((N + 2) & ~1));      // !FX: [warp][N]
    double*   esc = reinterpret_cast<double*>(H);                     //  FX: [2][N] side table, then
    uint32_t* acc = reinterpret_cast<uint32_t*>(esc + 2 * N);         //      [2][HKX_NW][N][3] window accumulators
    const size_t accn = (size_t)HKX_NW * N * 3;
#if XC_HKX_ROWCNT
    uint32_t* cnt = acc + 2 * accn;                                   //      [rows of this CTA][(N+1)/2] packed u16 cell counts
    const int CW = (N + 1) >> 1, cnt_rows = p.per / p.nx + 2;
#endif
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long s = p.s0 + blockIdx.y;
    const int c = blockIdx.x, C = gridDim.x;

    const double* eg = p.edges + s * (long)(N + 1);
    for (int k = tid; k <= N; k += blockDim.x) e[k] = eg[k];
    if (FX) {
        for (int i = tid; i < 2 * N; i += blockDim.x) esc[i] = 0.0;
        for (int i = tid; i < (int)(2 * accn); i += blockDim.x) acc[i] = 0u;
#if XC_HKX_ROWCNT
        for (int i = tid; i < cnt_rows * CW; i += blockDim.x) cnt[i] = 0u;
#endif
    } else {
        for (int i = tid; i < HK_WARPS * N; i += blockDim.x) H[i] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    const double span = e[N] - e[0];
    const float basef = (float)e[0];
    const float invf = span > 0.0 ? (float)((double)N / span) : 0.0f;
    // safety margin of the arithmetic bin guess, in bins (see hk_find_bin)
    float delta;
    {
        const double h = span / (double)N;
        double dev = 0.0;
        for (int k = tid; k <= N; k += blockDim.x) dev = fmax(dev, fabs(e[k] - (e[0] + (double)k * h)));
        dev = warp_max(dev);
        __shared__ double sdev[HK_WARPS];
        if (lane == 0) sdev[warp] = dev;
        __syncthreads();
        dev = 0.0;
        for (int w = 0; w < HK_WARPS; ++w) dev = fmax(dev, sdev[w]);
        const double errb = fabs(e[0] - (double)basef);
        const double d = (h > 0.0) ? (dev + errb) / h + 8.0 * (double)N * 5.9604644775390625e-8 + 1e-6 : 1.0;
        delta = (d < 0.5 && isfinite(d)) ? (float)d : 2.0f;       // 2.0 -> always take the exact path
    }

    const float* qs = p.q + s * (long)p.P;
    const int nx = p.nx, ny = p.ny;
    const int beg = c * p.per, end = min(p.P, beg + p.per);
    double2* Hw = H + (size_t)warp * N;
    uint8_t* tagw = reinterpret_cast<uint8_t*>(H + (size_t)HK_WARPS * N) + (size_t)warp * ((N + 15) & ~15);
    (void)tagw;

    __shared__ int sbase[2 * HK_WARPS];
    bool first = true; int eA = 0, eG = 0;
    // (FX) every warp of the CTA takes part in the first step -- `per` is a multiple of HK_WARPS*128 or the
    // loop bound below is padded so that the anchoring barrier is reached by all threads
    const int end_it = FX ? beg + ((end - beg + HK_WARPS * 128 - 1) / (HK_WARPS * 128)) * (HK_WARPS * 128) : end;
#if XC_HKX_LEAN
    const bool walk = nx >= HK_WARPS * 128;             // then a step crosses at most one row boundary
    int jw = (beg + warp * 128 + lane * 4) / nx, cw = (beg + warp * 128 + lane * 4) - jw * nx;
#endif
    for (int base = beg + warp * 128; base < end_it; base += HK_WARPS * 128) {
        const int i0 = base + lane * 4;
        const bool ok = i0 < end;                       // P, per and nx are multiples of 4: all-or-nothing
        int j = 0, col = 0;
        float4 qc = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F), a4 = qc, nn = qc, ss = qc;
        float wv = 0.f, ev = 0.f; double cx = 0.0, cy = 0.0;
        if (ok) {
#if XC_HKX_LEAN
            if (walk) { j = jw; col = cw; cw += HK_WARPS * 128; if (cw >= nx) { cw -= nx; ++jw; } }
            else { j = i0 / nx; col = i0 - j * nx; }
#else
            j = i0 / nx; col = i0 - j * nx;
#endif
            const float* row = qs + (long)j * nx;
            const int jm = j == 0 ? 0 : j - 1, jp = j == ny - 1 ? ny - 1 : j + 1;
            qc = __ldg(reinterpret_cast<const float4*>(row + col));
            a4 = __ldg(reinterpret_cast<const float4*>(p.dA + i0));
            nn = __ldg(reinterpret_cast<const float4*>(qs + (long)jp * nx + col));
            ss = __ldg(reinterpret_cast<const float4*>(qs + (long)jm * nx + col));
            wv = __ldg(row + (col == 0 ? nx - 1 : col - 1));
            ev = __ldg(row + (col + 4 == nx ? 0 : col + 4));
            cx = __ldg(p.cx + j); cy = __ldg(p.cy + j);
        }
        const double c0 = (double)qc.x, c1 = (double)qc.y, c2 = (double)qc.z, c3 = (double)qc.w;
        const double g0 = grad2_from(c1, (double)wv, (double)nn.x, (double)ss.x, cx, cy);
        const double g1 = grad2_from(c2, c0, (double)nn.y, (double)ss.y, cx, cy);
        const double g2 = grad2_from(c3, c1, (double)nn.z, (double)ss.z, cx, cy);
        const double g3 = grad2_from((double)ev, c2, (double)nn.w, (double)ss.w, cx, cy);
        const int b0 = ok ? hk_find_bin(qc.x, e, N, basef, invf, delta) : -1;
        const int b1 = ok ? hk_find_bin(qc.y, e, N, basef, invf, delta) : -1;
        const int b2 = ok ? hk_find_bin(qc.z, e, N, basef, invf, delta) : -1;
        const int b3 = ok ? hk_find_bin(qc.w, e, N, basef, invf, delta) : -1;
        // weights: dA (NaN -> 0) and |grad q|^2 * dA rounded in fp64 (NaN -> 0), core.py:444-449
        const double a0 = (a4.x == a4.x) ? (double)a4.x : 0.0, a1 = (a4.y == a4.y) ? (double)a4.y : 0.0;
        const double a2 = (a4.z == a4.z) ? (double)a4.z : 0.0, a3 = (a4.w == a4.w) ? (double)a4.w : 0.0;
        double p0 = __dmul_rn(g0, (double)a4.x), p1 = __dmul_rn(g1, (double)a4.y);
        double p2 = __dmul_rn(g2, (double)a4.z), p3 = __dmul_rn(g3, (double)a4.w);
        p0 = (p0 == p0) ? p0 : 0.0; p1 = (p1 == p1) ? p1 : 0.0;
        p2 = (p2 == p2) ? p2 : 0.0; p3 = (p3 == p3) ? p3 : 0.0;
        if (FX) {
            if (first) {                                    // anchor the windows on the CTA's first cells
                first = false;
                int mA = min(min(hkx_exp(b0 >= 0 ? a0 : 0.0), hkx_exp(b1 >= 0 ? a1 : 0.0)), min(hkx_exp(b2 >= 0 ? a2 : 0.0), hkx_exp(b3 >= 0 ? a3 : 0.0)));
                int mG = min(min(hkx_exp(b0 >= 0 ? p0 : 0.0), hkx_exp(b1 >= 0 ? p1 : 0.0)), min(hkx_exp(b2 >= 0 ? p2 : 0.0), hkx_exp(b3 >= 0 ? p3 : 0.0)));
                mA = __reduce_min_sync(XC_FULL, mA); mG = __reduce_min_sync(XC_FULL, mG);
                if (lane == 0) { sbase[warp] = mA; sbase[HK_WARPS + warp] = mG; }
                __syncthreads();
                mA = 0x7fffffff; mG = 0x7fffffff;
                for (int w = 0; w < HK_WARPS; ++w) { mA = min(mA, sbase[w]); mG = min(mG, sbase[HK_WARPS + w]); }
                eA = (mA == 0x7fffffff ? 1023 - 72 : mA) - HKX_MARGIN;
                eG = (mG == 0x7fffffff ? 1023 - 72 : mG) - HKX_MARGIN;
            }
#if XC_HKX_ROWCNT
            if (ok) {                                       // area: integer counts where dA is the row's own value
                const float aref = __ldg(p.dA + (long)j * nx);
                uint32_t* crow = cnt + (size_t)(j - beg / nx) * CW;
                if (b0 >= 0) { if (a4.x == aref) atomicAdd(crow + (b0 >> 1), 1u << ((b0 & 1) << 4)); else hkx_add(acc, esc, N, b0, eA, a0); }
                if (b1 >= 0) { if (a4.y == aref) atomicAdd(crow + (b1 >> 1), 1u << ((b1 & 1) << 4)); else hkx_add(acc, esc, N, b1, eA, a1); }
                if (b2 >= 0) { if (a4.z == aref) atomicAdd(crow + (b2 >> 1), 1u << ((b2 & 1) << 4)); else hkx_add(acc, esc, N, b2, eA, a2); }
                if (b3 >= 0) { if (a4.w == aref) atomicAdd(crow + (b3 >> 1), 1u << ((b3 & 1) << 4)); else hkx_add(acc, esc, N, b3, eA, a3); }
            }
#endif
            // smooth stretch of the field (neighbouring lanes in the same bin): combine in registers first
            const int nb = __shfl_down_sync(XC_FULL, b0, 1);
            const bool crowded = __popc(__ballot_sync(XC_FULL, b0 >= 0 && b0 == nb)) >= XC_HKX_CROWD;
            if (crowded) {                                  // the lane's own four cells first
                int c1 = b1, c2 = b2, c3 = b3;
                double x0 = a0, y0 = p0, x1 = a1, y1 = p1, x2 = a2, y2 = p2, x3 = a3, y3 = p3;
                if (c1 >= 0 && c1 == b0) { x0 += x1; y0 += y1; c1 = -1; }
                if (c2 >= 0) { if (c2 == b0) { x0 += x2; y0 += y2; c2 = -1; } else if (c2 == c1) { x1 += x2; y1 += y2; c2 = -1; } }
                if (c3 >= 0) { if (c3 == b0) { x0 += x3; y0 += y3; c3 = -1; } else if (c3 == c1) { x1 += x3; y1 += y3; c3 = -1; }
                               else if (c3 == c2) { x2 += x3; y2 += y3; c3 = -1; } }
                hkx_scatter(acc, acc + accn, esc, esc + N, N, eA, eG, b0, x0, y0, lane, true);
                if (__any_sync(XC_FULL, c1 >= 0)) hkx_scatter(acc, acc + accn, esc, esc + N, N, eA, eG, c1, x1, y1, lane, true);
                if (__any_sync(XC_FULL, c2 >= 0)) hkx_scatter(acc, acc + accn, esc, esc + N, N, eA, eG, c2, x2, y2, lane, true);
                if (__any_sync(XC_FULL, c3 >= 0)) hkx_scatter(acc, acc + accn, esc, esc + N, N, eA, eG, c3, x3, y3, lane, true);
            } else {
                hkx_scatter(acc, acc + accn, esc, esc + N, N, eA, eG, b0, a0, p0, lane, false);
                hkx_scatter(acc, acc + accn, esc, esc + N, N, eA, eG, b1, a1, p1, lane, false);
                hkx_scatter(acc, acc + accn, esc, esc + N, N, eA, eG, b2, a2, p2, lane, false);
                hkx_scatter(acc, acc + accn, esc, esc + N, N, eA, eG, b3, a3, p3, lane, false);
            }
            continue;
        }
#if XC_HK_DEDUP == 0
        hk_scatter(Hw, b0, a0, p0, lane);
        hk_scatter(Hw, b1, a1, p1, lane);
        hk_scatter(Hw, b2, a2, p2, lane);
        hk_scatter(Hw, b3, a3, p3, lane);
#elif XC_HK_DEDUP == 4
        hk_peel(Hw, hk_match(b0, lane), b0, a0, p0, lane);
        hk_peel(Hw, hk_match(b1, lane), b1, a1, p1, lane);
        hk_peel(Hw, hk_match(b2, lane), b2, a2, p2, lane);
        hk_peel(Hw, hk_match(b3, lane), b3, a3, p3, lane);
#elif XC_HK_DEDUP == 1
        const unsigned m0 = hk_match(b0, lane), m1 = hk_match(b1, lane), m2 = hk_match(b2, lane), m3 = hk_match(b3, lane);
        hk_peel(Hw, m0, b0, a0, p0, lane);
        hk_peel(Hw, m1, b1, a1, p1, lane);
        hk_peel(Hw, m2, b2, a2, p2, lane);
        hk_peel(Hw, m3, b3, a3, p3, lane);
#else
        hk_tag(Hw, tagw, b0, a0, p0, lane);
        hk_tag(Hw, tagw, b1, a1, p1, lane);
        hk_tag(Hw, tagw, b2, a2, p2, lane);
        hk_tag(Hw, tagw, b3, a3, p3, lane);
#endif
    }
    __syncthreads();
    double* out = p.part + ((size_t)(blockIdx.y + p.s0) * C + c) * 2 * N;
    if (FX) {
        for (int idx = tid; idx < 2 * N; idx += blockDim.x) {
            const int k = idx / N, n = idx - k * N;
            const int eb = k == 0 ? eA : eG;
            const uint32_t* a = acc + (size_t)k * accn + (size_t)n * 3;
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < HKX_NW; ++w)                     // smallest window first
                t += hkx_window(a + (size_t)w * N * 3, eb + w * HKX_WBITS - 1075);
#if XC_HKX_ROWCNT
            if (k == 0 && end > beg) {                           // + dA[row] * cells of that row in the bin, row by row (exact products)
                const int jf = beg / nx, nrows = (end - 1) / nx - jf + 1;
                for (int r = 0; r < nrows; ++r) {
                    const uint32_t cn = (cnt[(size_t)r * CW + (n >> 1)] >> ((n & 1) << 4)) & 0xffffu;
                    if (cn) t += (double)__ldg(p.dA + (long)(jf + r) * nx) * (double)cn;
                }
            }
#endif
            out[idx] = t + esc[idx];
        }
        return;
    }
    for (int idx = tid; idx < 2 * N; idx += blockDim.x) {
        const int k = idx / N, n = idx - k * N;
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < HK_WARPS; ++w) {
            const double2 t = H[(size_t)w * N + n];
            acc += k == 0 ? t.x : t.y;
        }
        out[idx] = acc;
    }
}

}  // namespace xc

using namespace xc;

// Returns 1 when the dedicated kernel does not apply (the caller then uses the
// general kernel), 0 when it was launched, 2 on error.
int xc::hist_keff_try(const void* q, int q_dtype, long S, long P, const double* edges, int N,
                      const void* dA, int dA_dtype, const StencilArgs* st, int C,
                      double* part, void* stream)
{
    if (q_dtype != XC_F32 || dA_dtype != XC_F32 || !st) return 1;
    if ((P & 3) || (st->nx & 3) || P >= (1L << 30) || st->nx < 8) return 1;
    if ((((uintptr_t)q) & 15) || (((uintptr_t)dA) & 15)) return 1;
    static const char* off = getenv("XCB200_NO_HIST_KEFF");
    if (off) return 1;
    // XCB200_HIST_FX=0: warp-private fp64 histograms with lane de-duplication instead of the
    // exact integer accumulators (A/B timing, cross-check in the tests)
    static const char* fxe = getenv("XCB200_HIST_FX");
    const size_t smem_fx = (size_t)((N + 2) & ~1) * 8 + (size_t)2 * N * 8 + (size_t)2 * HKX_NW * N * 3 * 4;
    const size_t smem_rmw = (size_t)((N + 2) & ~1) * 8 + (size_t)HK_WARPS * N * 16 + (size_t)HK_WARPS * ((N + 15) & ~15);
    const long per = (((P + C - 1) / C) + 3) & ~3L;
#if XC_HKX_ROWCNT
    const size_t smem_cnt = (size_t)(per / st->nx + 2) * ((N + 1) >> 1) * 4;
    const size_t smem_fx_all = smem_fx + smem_cnt;
    const bool fx = !(fxe && fxe[0] == '0') && smem_fx_all <= 113 * 1024 && per <= (1L << 18) && st->nx <= 65535;
#define smem_fx smem_fx_all
#else
    const bool fx = !(fxe && fxe[0] == '0') && smem_fx <= 113 * 1024 && per <= (1L << 18);
#endif   // <= 2^18 terms of < 2^77 per bin and CTA: 95 bits
    const size_t smem = fx ? smem_fx : smem_rmw;
    if (smem > 100 * 1024 && !fx) return 1;           // two CTAs per SM
    HistKeffParams p;
    p.q = (const float*)q; p.P = (int)P; p.per = (int)per;
    p.edges = edges; p.N = N; p.dA = (const float*)dA;
    p.ny = st->ny; p.nx = st->nx; p.cx = st->cx; p.cy = st->cy; p.part = part;
    if (cudaFuncSetAttribute(fx ? k_hist_keff<true> : k_hist_keff<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        set_error("k_hist_keff: cannot reserve %zu bytes of shared memory", smem); return 2;
    }
    for (long s0 = 0; s0 < S; s0 += 65535) {
        const long ns = S - s0 < 65535 ? S - s0 : 65535;
        p.s0 = s0;
        if (fx) k_hist_keff<true><<<dim3((unsigned)C, (unsigned)ns), HK_WARPS * 32, smem, (cudaStream_t)stream>>>(p);
        else    k_hist_keff<false><<<dim3((unsigned)C, (unsigned)ns), HK_WARPS * 32, smem, (cudaStream_t)stream>>>(p);
        count_launch();
        if (cudaGetLastError() != cudaSuccess) { set_error("k_hist_keff launch failed"); return 2; }
    }
    return 0;
}
