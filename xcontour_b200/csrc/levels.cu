// Kernel (1): per-slice NaN-skipping min/max and equally spaced contour levels
// (Contour2D.cal_contours, xcontour/core.py:222-249), plus the histogram bin
// edges of _histogram (core.py:1273-1281, 1296-1304).
//
// HBM-bound streaming read of q: 128-bit loads, 4 independent loads in flight
// per thread, one (min,max) pair per CTA written as a partial so the result is
// order-independent and needs no atomics.
#include "common.cuh"
#include "internal.h"
#include <math_constants.h>

namespace xc {

template <typename T> struct Vec4;
template <> struct Vec4<float>  { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

template <typename T>
__device__ __forceinline__ void mm_update(T v, T& mn, T& mx);
template <> __device__ __forceinline__ void mm_update<float>(float v, float& mn, float& mx) {
    mn = fminf(mn, v); mx = fmaxf(mx, v);          // fmin/fmax drop NaN operands
}
template <> __device__ __forceinline__ void mm_update<double>(double v, double& mn, double& mx) {
    mn = fmin(mn, v); mx = fmax(mx, v);
}

// grid = (C, S); CTA c reduces cells [c*per, min(P,(c+1)*per)) of slice s.
template <typename T>
__global__ void __launch_bounds__(256)
k_minmax_partial(const T* __restrict__ q, long P, long per, double* __restrict__ part)
{
    const long s = blockIdx.y;
    const int  c = blockIdx.x, C = gridDim.x;
    const T* qs = q + s * P;
    long beg = (long)c * per;
    long end = beg + per < P ? beg + per : P;
    T mn = (T)CUDART_INF, mx = (T)-CUDART_INF;

    const bool vec_ok = ((P & 3) == 0) && ((((uintptr_t)q) & 15) == 0) && ((per & 3) == 0);
    if (vec_ok) {
        // 16-byte loads (two per thread for fp64), 4-deep unroll
        const long nv = (end - beg) >> 2;
        const T* base = qs + beg;
        long i = threadIdx.x;
        for (; i + 3 * (long)blockDim.x < nv; i += 4 * (long)blockDim.x) {
            T v[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const T* p = base + ((i + u * (long)blockDim.x) << 2);
                if (sizeof(T) == 4) {
                    float4 t = __ldg((const float4*)p);
                    v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w;
                } else {
                    double2 a = __ldg((const double2*)p), b = __ldg((const double2*)p + 1);
                    v[u][0] = a.x; v[u][1] = a.y; v[u][2] = b.x; v[u][3] = b.y;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int e = 0; e < 4; ++e) mm_update<T>(v[u][e], mn, mx);
        }
        for (; i < nv; i += blockDim.x) {
            const T* p = base + (i << 2);
#pragma unroll
            for (int e = 0; e < 4; ++e) mm_update<T>(__ldg(p + e), mn, mx);
        }
    } else {
        for (long i = beg + threadIdx.x; i < end; i += blockDim.x)
            mm_update<T>(__ldg(qs + i), mn, mx);
    }
    double dmn = warp_min((double)mn), dmx = warp_max((double)mx);
    __shared__ double smn[8], smx[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { smn[w] = dmn; smx[w] = dmx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
            dmn = fmin(dmn, smn[k]); dmx = fmax(dmx, smx[k]);
        }
        part[(s * C + c) * 2 + 0] = dmn;
        part[(s * C + c) * 2 + 1] = dmx;
    }
}

// ---------------------------------------------------------------------------
// The same reduction fed by the bulk asynchronous copy engine (TMA, cp.async.bulk -> UBLKCP): one elected thread
// streams 16 KB chunks of the CTA's range into a 4-stage shared-memory ring, every chunk announced on an mbarrier
// by its byte count; the 256 threads drain a stage with four LDS.128 each and hand it back through a second
// ("empty") mbarrier.  Against the register-staged loop above this keeps 64 KB per CTA (192 KB per SM) in flight
// without a register per outstanding load and without per-load address arithmetic in the issue stream -- the
// kernel is HBM-bound (ncu r1: 76 % of the measured copy peak), so what matters is bytes in flight.
constexpr int MB_STAGES = 4, MB_BYTES = 16384, MB_NT = 256;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// grid = (C, S), block = MB_NT, dynamic shared memory = MB_STAGES * MB_BYTES; requires 16-byte aligned ranges
template <typename T>
__global__ void __launch_bounds__(MB_NT)
k_minmax_bulk(const T* __restrict__ q, long P, long per, double* __restrict__ part)
{
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) unsigned long long bars[2 * MB_STAGES];
    const long s = blockIdx.y;
    const int c = blockIdx.x, C = gridDim.x, tid = threadIdx.x;
    const long beg = (long)c * per, end = beg + per < P ? beg + per : P;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(q + s * P + beg);
    const long bytes = end > beg ? (end - beg) * (long)sizeof(T) : 0;
    const int nchunk = (int)((bytes + MB_BYTES - 1) / MB_BYTES);
    const uint32_t ring_sh = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t bar_sh = (uint32_t)__cvta_generic_to_shared(bars);
    auto full = [&](int st) { return bar_sh + 8u * st; };
    auto empty = [&](int st) { return bar_sh + 8u * (MB_STAGES + st); };
    auto chunk_bytes = [&](int k) { const long r = bytes - (long)k * MB_BYTES; return (uint32_t)(r < MB_BYTES ? r : MB_BYTES); };
    if (tid == 0) {
        for (int st = 0; st < MB_STAGES; ++st) { mbar_init(full(st), 1); mbar_init(empty(st), MB_NT); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int k = 0; k < MB_STAGES && k < nchunk; ++k) {
            mbar_expect_tx(full(k), chunk_bytes(k));
            bulk_g2s(ring_sh + k * MB_BYTES, src + (long)k * MB_BYTES, chunk_bytes(k), full(k));
        }
    T mn = (T)CUDART_INF, mx = (T)-CUDART_INF;
    for (int k = 0; k < nchunk; ++k) {
        const int st = k % MB_STAGES; const uint32_t ph = (uint32_t)(k / MB_STAGES) & 1u;
        mbar_wait(full(st), ph);
        const uint32_t nb = chunk_bytes(k);
        const uint4* b4 = reinterpret_cast<const uint4*>(ring + st * MB_BYTES);
#pragma unroll
        for (int u = 0; u < MB_BYTES / 16 / MB_NT; ++u) {
            const int i = tid + u * MB_NT;
            if ((uint32_t)i * 16u < nb) {
                const uint4 w = b4[i];
                if (sizeof(T) == 4) {
                    mm_update<T>((T)__uint_as_float(w.x), mn, mx); mm_update<T>((T)__uint_as_float(w.y), mn, mx);
                    mm_update<T>((T)__uint_as_float(w.z), mn, mx); mm_update<T>((T)__uint_as_float(w.w), mn, mx);
                } else {
                    mm_update<T>((T)__hiloint2double((int)w.y, (int)w.x), mn, mx);
                    mm_update<T>((T)__hiloint2double((int)w.w, (int)w.z), mn, mx);
                }
            }
        }
        mbar_arrive(empty(st));                                   // this thread is done with the stage
        if (tid == 0 && k + MB_STAGES < nchunk) {                 // refill it once everybody is
            mbar_wait(empty(st), ph);
            mbar_expect_tx(full(st), chunk_bytes(k + MB_STAGES));
            bulk_g2s(ring_sh + st * MB_BYTES, src + (long)(k + MB_STAGES) * MB_BYTES, chunk_bytes(k + MB_STAGES), full(st));
        }
    }
    double dmn = warp_min((double)mn), dmx = warp_max((double)mx);
    __shared__ double smn[MB_NT / 32], smx[MB_NT / 32];
    const int w = tid >> 5, l = tid & 31;
    if (l == 0) { smn[w] = dmn; smx[w] = dmx; }
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < MB_NT / 32; ++k) { dmn = fmin(dmn, smn[k]); dmx = fmax(dmx, smx[k]); }
        part[(s * C + c) * 2 + 0] = dmn;
        part[(s * C + c) * 2 + 1] = dmx;
    }
}

// grid = S.  Reduces the C partials, then writes the N levels.
//   steps = (1.0/(N-1)) * f64(end -_T start);  level_k = steps*k + f64(start)
// with separately rounded multiply/add (NumPy evaluates them as two ufuncs) and
// a final cast to the requested contour dtype (core.py:228-246).
__global__ void k_levels(const double* __restrict__ part, int C, int N, int increase,
                         int q_is_f32, int out_is_f32,
                         double* __restrict__ levels, double* __restrict__ minmax,
                         double* __restrict__ edges, int32_t* __restrict__ decreasing,
                         int32_t* __restrict__ flag_to_clear, int edges_keep_ctr_dtype)
{
    const long s = blockIdx.x;
    if (flag_to_clear && s == 0 && threadIdx.x == 0) *flag_to_clear = 0;
    __shared__ double sh[2];
    if (threadIdx.x < 32) {
        double mn = CUDART_INF, mx = -CUDART_INF;
        for (int c = threadIdx.x; c < C; c += 32) {
            mn = fmin(mn, part[(s * C + c) * 2 + 0]);
            mx = fmax(mx, part[(s * C + c) * 2 + 1]);
        }
        mn = warp_min(mn); mx = warp_max(mx);
        if (threadIdx.x == 0) {
            if (mn > mx) { mn = CUDART_NAN; mx = CUDART_NAN; }   // all-NaN slice
            sh[0] = mn; sh[1] = mx;
            if (minmax) { minmax[s * 2] = mn; minmax[s * 2 + 1] = mx; }
        }
    }
    __syncthreads();
    const double start = increase ? sh[0] : sh[1];
    const double end   = increase ? sh[1] : sh[0];
    double diff;
    if (q_is_f32) diff = (double)__fsub_rn((float)end, (float)start);
    else          diff = __dsub_rn(end, start);
    const double steps = __dmul_rn(1.0 / (double)(N - 1), diff);
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        double v = __dadd_rn(__dmul_rn(steps, (double)k), start);
        if (out_is_f32) v = (double)__double2float_rn(v);
        levels[s * N + k] = v;
    }
    if (edges) {            // per-'time' branch edges (core.py:1273-1281), same rules as k_hist_edges
        auto level = [&](int k) {
            double v = __dadd_rn(__dmul_rn(steps, (double)k), start);
            return out_is_f32 ? (double)__double2float_rn(v) : v;
        };
        const double c0 = level(0), cl = level(N - 1);
        const bool binc = c0 < cl;
        const double first = binc ? c0 : cl, last = binc ? cl : c0;
        const double d = out_is_f32 ? (double)__fsub_rn((float)last, (float)first) : __dsub_rn(last, first);
        const double step = __ddiv_rn(d, (double)(N - 1));
        const bool e32 = out_is_f32 && edges_keep_ctr_dtype;    // NumPy >= 2 scalar rules: as k_hist_edges with time_branch = 0
        double* e = edges + s * (long)(N + 1);
        for (int k = threadIdx.x; k <= N; k += blockDim.x) {
            double v = (k == 0) ? __dsub_rn(first, step) : (binc ? level(k - 1) : level(N - k));
            if (k == 0 && e32) v = (double)__double2float_rn(v);
            if (k == N) v = e32 ? (double)__fadd_rn((float)v, 1e-8f) : __dadd_rn(v, 1e-8);
            e[k] = v;
        }
        if (threadIdx.x == 0 && decreasing) decreasing[s] = binc ? 0 : 1;
    }
}

// grid = S.  edges[s][0..N] ascending, see xc_hist_edges in xcb200.h.
__global__ void k_hist_edges(const double* __restrict__ levels, int N, int ctr_is_f32,
                             int time_branch, double* __restrict__ edges,
                             int32_t* __restrict__ decreasing)
{
    const long s = blockIdx.x;
    const double* c = levels + s * N;
    double* e = edges + s * (long)(N + 1);
    const double c0 = c[0], cl = c[N - 1];
    const bool binc = c0 < cl;                                  // core.py:1273 / 1296
    const double first = binc ? c0 : cl, last = binc ? cl : c0;
    // step = (last - first)/(N-1): the difference in the contour dtype, the
    // division by a Python int promoted to fp64 (NumPy-1.x scalar rules).
    double d = ctr_is_f32 ? (double)__fsub_rn((float)last, (float)first)
                          : __dsub_rn(last, first);
    const double step = __ddiv_rn(d, (double)(N - 1));
    const bool e32 = ctr_is_f32 && !time_branch;   // np.insert keeps the array dtype
    for (int k = threadIdx.x; k <= N; k += blockDim.x) {
        double v;
        if (k == 0) {
            v = __dsub_rn(first, step);
            if (e32) v = (double)__double2float_rn(v);
        } else {
            v = binc ? c[k - 1] : c[N - k];
        }
        if (k == N) {                                           // xhistogram: + 1e-8
            if (e32) v = (double)__fadd_rn((float)v, 1e-8f);
            else     v = __dadd_rn(v, 1e-8);
        }
        e[k] = v;
    }
    if (threadIdx.x == 0 && decreasing) decreasing[s] = binc ? 0 : 1;
}

}  // namespace xc

using namespace xc;

static long minmax_ctas_per_slice(long S, long P)
{
    long want = (long)sm_count() * 8;
    long C = (want + S - 1) / S;
    long maxC = (P + 8191) / 8192;
    if (C > maxC) C = maxC;
    if (C > 1024) C = 1024;
    if (C < 1) C = 1;
    return C;
}

extern "C" size_t xc_minmax_levels_workspace_bytes(long S, long P)
{
    return 256 + (size_t)S * minmax_ctas_per_slice(S, P) * 2 * sizeof(double);
}

extern "C" int xc_minmax_levels(const void* q, int q_dtype, long S, long P,
                                int N, int increase, int out_dtype,
                                double* levels, double* minmax,
                                void* workspace, size_t ws_bytes, void* stream)
{
    return minmax_levels_impl(q, q_dtype, S, P, N, increase, out_dtype, levels, minmax,
                              nullptr, nullptr, nullptr, workspace, ws_bytes, stream, 0);
}

int xc::minmax_levels_impl(const void* q, int q_dtype, long S, long P, int N, int increase, int out_dtype,
                           double* levels, double* minmax, double* edges, int32_t* decreasing,
                           int32_t* flag_to_clear, void* workspace, size_t ws_bytes, void* stream,
                           int edges_keep_ctr_dtype)
{
    XC_REQUIRE(q && levels, "xc_minmax_levels: null pointer");
    XC_REQUIRE(S > 0 && P > 0 && N >= 2, "xc_minmax_levels: need S>0, P>0, N>=2");
    XC_REQUIRE(q_dtype == XC_F32 || q_dtype == XC_F64, "xc_minmax_levels: bad dtype");
    XC_REQUIRE(S <= 65535L * 32768L, "xc_minmax_levels: too many slices");
    long C = minmax_ctas_per_slice(S, P);
    XC_REQUIRE(workspace && ws_bytes >= xc_minmax_levels_workspace_bytes(S, P),
               "xc_minmax_levels: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(workspace, ws_bytes);
    double* part = ar.take<double>((size_t)S * C * 2);
    // bulk-copy (TMA) variant: 16-byte aligned slices, a handful of 16 KB chunks per CTA (three CTAs per SM)
    static const char* no_bulk = getenv("XCB200_NO_BULK");
    const size_t esz = q_dtype == XC_F32 ? 4 : 8;
    long Cb = ((long)sm_count() * 3 + S - 1) / S; if (Cb < 1) Cb = 1; if (Cb > C) Cb = C;
    const bool bulk = !no_bulk && (((uintptr_t)q) & 15) == 0 && (P * esz) % 16 == 0 && P * (long)esz / Cb >= 4 * MB_BYTES;
    if (bulk) C = Cb;
    long per = ((P + C - 1) / C + 3) & ~3L;
    if (bulk) {
        const int sm = MB_STAGES * MB_BYTES;
        if (q_dtype == XC_F32) XC_CUDA_OK(cudaFuncSetAttribute(k_minmax_bulk<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        else                   XC_CUDA_OK(cudaFuncSetAttribute(k_minmax_bulk<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        for (long s0 = 0; s0 < S; s0 += 65535) {
            long ns = S - s0 < 65535 ? S - s0 : 65535;
            dim3 grid((unsigned)C, (unsigned)ns);
            if (q_dtype == XC_F32) k_minmax_bulk<float><<<grid, MB_NT, sm, st>>>((const float*)q + s0 * P, P, per, part + s0 * C * 2);
            else                   k_minmax_bulk<double><<<grid, MB_NT, sm, st>>>((const double*)q + s0 * P, P, per, part + s0 * C * 2);
            XC_LAUNCH_OK();
        }
    } else
    // gridDim.y is limited to 65535: walk the slices in groups
    for (long s0 = 0; s0 < S; s0 += 65535) {
        long ns = S - s0 < 65535 ? S - s0 : 65535;
        dim3 grid((unsigned)C, (unsigned)ns);
        if (q_dtype == XC_F32)
            k_minmax_partial<float><<<grid, 256, 0, st>>>((const float*)q + s0 * P, P, per, part + s0 * C * 2);
        else
            k_minmax_partial<double><<<grid, 256, 0, st>>>((const double*)q + s0 * P, P, per, part + s0 * C * 2);
        XC_LAUNCH_OK();
    }
    k_levels<<<(unsigned)S, 128, 0, st>>>(part, (int)C, N, increase, q_dtype == XC_F32,
                                          out_dtype == XC_F32, levels, minmax, edges, decreasing, flag_to_clear,
                                          edges_keep_ctr_dtype);
    XC_LAUNCH_OK();
    return 0;
}

extern "C" int xc_hist_edges(const double* levels, long S, int N, int ctr_dtype,
                             int time_branch, double* edges, int32_t* decreasing,
                             void* stream)
{
    XC_REQUIRE(levels && edges, "xc_hist_edges: null pointer");
    XC_REQUIRE(S > 0 && N >= 2, "xc_hist_edges: need S>0, N>=2");
    k_hist_edges<<<(unsigned)S, 128, 0, (cudaStream_t)stream>>>(
        levels, N, ctr_dtype == XC_F32, time_branch, edges, decreasing);
    XC_LAUNCH_OK();
    return 0;
}
