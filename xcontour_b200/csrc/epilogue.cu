// Fused contour-space epilogue of the Keff + LWA batch (one CTA per slice):
// reduces the per-CTA histogram partials in a fixed order, scans them into the
// CDFs (area, int |grad q|^2 dA), then -- with exactly the arithmetic of the
// stand-alone kernels in hist.cu / contour_ops.cu -- latEq = np.interp(area,
// table, coord), Lmin, d/dA, Leq2, nkeff, the sorted profile
// Q = np.interp(eq_coord, latEq, ctr) and its sortedness flag for the LWA kernel.
// Replaces nine tiny launches per pass by one (they were launch-latency bound).
#include "common.cuh"
#include "internal.h"
#include <math_constants.h>

namespace xc {

constexpr double kRe = 6371200.0;
constexpr double kPiE = 3.141592653589793;

__device__ __forceinline__ double grad_s(const double* f, int k, int N, bool as_f32)
{
    const int km = k == 0 ? 0 : k - 1, kp = k == N - 1 ? N - 1 : k + 1;
    const bool interior = (k > 0) && (k < N - 1);
    if (as_f32) {
        float d = __fsub_rn((float)f[kp], (float)f[km]);
        return (double)(interior ? __fdiv_rn(d, 2.0f) : d);
    }
    double d = __dsub_rn(f[kp], f[km]);
    return interior ? __ddiv_rn(d, 2.0) : d;
}

struct EpiParams {
    const double* part; int C; int N; int lt;
    const int32_t* decreasing;
    const double* ctr; int ctr_f32;
    const double* table; const double* table_coord; int n_table;
    const double* eq_coord; int ny;
    double keff_mask; int increase;
    double *area, *intg, *latEq, *Lmin, *dint, *dq, *Leq2, *nkeff, *Qref;
    int32_t* sorted; int32_t* any_unsorted;
};

__global__ void __launch_bounds__(1024) k_scan_epilogue(const EpiParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int N = p.N;
    double* sa = reinterpret_cast<double*>(smem);   // area   [N]
    double* sg = sa + N;                            // intgrdS[N]
    double* sc = sg + N;                            // ctr    [N]
    double* sl = sc + N;                            // latEq  [N]
    double* stab = sl + N;                          // A(Yeq) table       [n_table]
    double* scrd = stab + p.n_table;                // table coordinates  [n_table]
    const long s = blockIdx.x;
    const int tid = threadIdx.x;
    const bool rev = p.decreasing && p.decreasing[s] != 0;

    // fixed-order reduction of the C partials (8 loads in flight)
    for (int idx = tid; idx < 2 * N; idx += blockDim.x) {
        const int k = idx / N, n = idx - k * N;
        const double* pp = p.part + ((size_t)s * p.C * 2 + k) * N + n;
        const size_t cs = (size_t)2 * N;
        double acc = 0.0; int c = 0;
        for (; c + 8 <= p.C; c += 8) {
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = pp[(size_t)(c + u) * cs];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += t[u];
        }
        for (; c < p.C; ++c) acc += pp[(size_t)c * cs];
        (k == 0 ? sa : sg)[n] = acc;
    }
    for (int n = tid; n < N; n += blockDim.x) sc[n] = p.ctr[s * N + n];
    for (int n = tid; n < p.n_table; n += blockDim.x) { stab[n] = p.table[n]; scrd[n] = p.table_coord[n]; }
    __syncthreads();
    // sequential running sums, exactly np.cumsum's order (core.py:1320): empty bins
    // leave the CDFs bit-for-bit flat (so d/dA sees exact zeros where the reference
    // does); the two accumulators are scanned by two different warps at once
    if (tid == 0)  serial_cumsum(sa, N);
    if (tid == 32) serial_cumsum(sg, N);
    __syncthreads();
    // cdf[-1] - cdf for the 'greater than' case (core.py:1322-1323), then the flip
    // that makes the contour index ascend (core.py:454-455) -- both in place
    const double ta = sa[N - 1], tg = sg[N - 1];
    __syncthreads();
    if (!p.lt) {
        for (int n = tid; n < N; n += blockDim.x) { sa[n] = ta - sa[n]; sg[n] = tg - sg[n]; }
        __syncthreads();
    }
    if (rev) {
        for (int n = tid; n < N / 2; n += blockDim.x) {
            double t = sa[n]; sa[n] = sa[N - 1 - n]; sa[N - 1 - n] = t;
            t = sg[n]; sg[n] = sg[N - 1 - n]; sg[N - 1 - n] = t;
        }
        __syncthreads();
    }
    // latEq = Table.lookup_coordinates(area): direction from the table (core.py:1122-1126)
    const bool trev = !(stab[p.n_table - 1] > stab[0]);
    for (int n = tid; n < N; n += blockDim.x)
        sl[n] = np_interp(sa[n], stab, scrd, p.n_table, trev);
    __syncthreads();
    for (int n = tid; n < N; n += blockDim.x) {
        const size_t o = (size_t)s * N + n;
        const double area = sa[n], intg = sg[n], latEq = sl[n];
        const double rad = __dmul_rn(latEq, kPiE / 180.0);
        const double Lmin = __dmul_rn(__dmul_rn(__dmul_rn(2.0, kPiE), kRe), cos(rad));
        const double da = grad_s(sa, n, N, false);
        const double dint = __ddiv_rn(grad_s(sg, n, N, false), da);
        const double dq = __ddiv_rn(grad_s(sc, n, N, p.ctr_f32 != 0), da);
        const double Leq2 = __ddiv_rn(dint, __dmul_rn(dq, dq));
        const double nk = __ddiv_rn(__ddiv_rn(Leq2, Lmin), Lmin);
        if (p.area) p.area[o] = area;
        if (p.intg) p.intg[o] = intg;
        if (p.latEq) p.latEq[o] = latEq;
        if (p.Lmin) p.Lmin[o] = Lmin;
        if (p.dint) p.dint[o] = dint;
        if (p.dq) p.dq[o] = dq;
        if (p.Leq2) p.Leq2[o] = Leq2;
        if (p.nkeff) p.nkeff[o] = (nk < p.keff_mask) ? nk : CUDART_NAN;
    }
    // Q = interp_to_coords(eq_coord, latEq, ctr); direction as core.py:1085-1088
    const bool qrev = !(sl[0] < sl[N - 1]);
    const double sgn = p.increase ? 1.0 : -1.0;
    int bad = 0;
    double* Qs = p.Qref + (size_t)s * p.ny;
    for (int m = tid; m < p.ny; m += blockDim.x)
        Qs[m] = np_interp(p.eq_coord[m], sl, sc, N, qrev);
    __syncthreads();                           // global writes of this block are visible to it
    for (int m = tid; m < p.ny; m += blockDim.x) {
        const double a = sgn * Qs[m];
        if (isnan(a)) bad = 1;
        if (m + 1 < p.ny && !(a <= sgn * Qs[m + 1])) bad = 1;
    }
    bad = __syncthreads_or(bad);
    if (tid == 0) {
        p.sorted[s] = bad ? 0 : 1;
        if (bad) atomicOr(p.any_unsorted, 1);
    }
}

// fixed-order reduction of many per-CTA partials ([S][C][2][N] -> [S][1][2][N]) spread over the GPU: with hundreds
// of partials per slice (config 5: 296 x 2 x 2048 doubles) one CTA per slice would be latency-bound on it
__global__ void __launch_bounds__(256) k_reduce_partials(const double* __restrict__ part, int C, int N2, double* __restrict__ red)
{
    const long s = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N2) return;
    const double* pp = part + (size_t)s * C * N2 + idx;
    double acc = 0.0; int c = 0;
    for (; c + 8 <= C; c += 8) {
        double t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = __ldg(pp + (size_t)(c + u) * N2);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += t[u];
    }
    for (; c < C; ++c) acc += __ldg(pp + (size_t)c * N2);
    red[(size_t)s * N2 + idx] = acc;
}

}  // namespace xc

using namespace xc;

int xc::scan_epilogue(const double* part, int C, long S, int N, int lt, const int32_t* decreasing,
                      const double* ctr, int ctr_f32,
                      const double* table, const double* table_coord, int n_table,
                      const double* eq_coord, int ny, double keff_mask, int increase,
                      double* area, double* intg, double* latEq, double* Lmin, double* dint,
                      double* dq, double* Leq2, double* nkeff, double* Qref,
                      int32_t* sorted, int32_t* any_unsorted, void* stream, double* reduce_buf)
{
    if (reduce_buf && C > 16) {                       // same summation order (c ascending), many CTAs per slice
        XC_REQUIRE(S <= 65535, "xc_keff_lwa_batch: too many slices per pass");
        dim3 grid((unsigned)((2 * N + 255) / 256), (unsigned)S);
        k_reduce_partials<<<grid, 256, 0, (cudaStream_t)stream>>>(part, C, 2 * N, reduce_buf);
        XC_LAUNCH_OK();
        part = reduce_buf; C = 1;
    }
    EpiParams p;
    p.part = part; p.C = C; p.N = N; p.lt = lt; p.decreasing = decreasing;
    p.ctr = ctr; p.ctr_f32 = ctr_f32; p.table = table; p.table_coord = table_coord; p.n_table = n_table;
    p.eq_coord = eq_coord; p.ny = ny; p.keff_mask = keff_mask; p.increase = increase;
    p.area = area; p.intg = intg; p.latEq = latEq; p.Lmin = Lmin; p.dint = dint; p.dq = dq;
    p.Leq2 = Leq2; p.nkeff = nkeff; p.Qref = Qref; p.sorted = sorted; p.any_unsorted = any_unsorted;
    const size_t sm = ((size_t)4 * N + 2 * (size_t)n_table) * sizeof(double);
    XC_REQUIRE(sm <= 200 * 1024, "xc_keff_lwa_batch: N too large for the fused epilogue");
    if (sm > 48 * 1024)
        XC_CUDA_OK(cudaFuncSetAttribute(k_scan_epilogue, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    // one CTA per slice: 256 threads cover config 4's 361 levels / 721 rows in two or three rounds; many levels or
    // rows (config 5: 2048 / 4096) get 1024
    const unsigned nt = (N > 512 || ny > 1024 || n_table > 1024) ? 1024u : 256u;
    k_scan_epilogue<<<(unsigned)S, nt, sm, (cudaStream_t)stream>>>(p);
    XC_LAUNCH_OK();
    return 0;
}
