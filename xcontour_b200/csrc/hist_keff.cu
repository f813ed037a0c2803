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

// grid = (C, nslices), block = 16 warps, 2 CTAs per SM.
__global__ void __launch_bounds__(HK_WARPS * 32, 2)
k_hist_keff(const HistKeffParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int N = p.N;
    double*  e = reinterpret_cast<double*>(smem);
    double2* H = reinterpret_cast<double2*>(e + ((N + 2) & ~1));      // [warp][N]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long s = p.s0 + blockIdx.y;
    const int c = blockIdx.x, C = gridDim.x;

    const double* eg = p.edges + s * (long)(N + 1);
    for (int k = tid; k <= N; k += blockDim.x) e[k] = eg[k];
    for (int i = tid; i < HK_WARPS * N; i += blockDim.x) H[i] = make_double2(0.0, 0.0);
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

    for (int base = beg + warp * 128; base < end; base += HK_WARPS * 128) {
        const int i0 = base + lane * 4;
        const bool ok = i0 < end;                       // P, per and nx are multiples of 4: all-or-nothing
        int j = 0, col = 0;
        float4 qc = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F), a4 = qc, nn = qc, ss = qc;
        float wv = 0.f, ev = 0.f; double cx = 0.0, cy = 0.0;
        if (ok) {
            j = i0 / nx; col = i0 - j * nx;
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
    const size_t smem = (size_t)((N + 2) & ~1) * 8 + (size_t)HK_WARPS * N * 16 + (size_t)HK_WARPS * ((N + 15) & ~15);
    if (smem > 100 * 1024) return 1;                  // two CTAs per SM
    static const char* off = getenv("XCB200_NO_HIST_KEFF");
    if (off) return 1;
    HistKeffParams p;
    p.q = (const float*)q; p.P = (int)P; p.per = (int)((((P + C - 1) / C) + 3) & ~3L);
    p.edges = edges; p.N = N; p.dA = (const float*)dA;
    p.ny = st->ny; p.nx = st->nx; p.cx = st->cx; p.cy = st->cy; p.part = part;
    if (cudaFuncSetAttribute(k_hist_keff, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        set_error("k_hist_keff: cannot reserve %zu bytes of shared memory", smem); return 2;
    }
    for (long s0 = 0; s0 < S; s0 += 65535) {
        const long ns = S - s0 < 65535 ? S - s0 : 65535;
        p.s0 = s0;
        k_hist_keff<<<dim3((unsigned)C, (unsigned)ns), HK_WARPS * 32, smem, (cudaStream_t)stream>>>(p);
        count_launch();
        if (cudaGetLastError() != cudaSuccess) { set_error("k_hist_keff launch failed"); return 2; }
    }
    return 0;
}
