// Contour-space (O(N) per slice) kernels: np.interp lookups, d/dA, Keff epilogue.
// Replaces Table.lookup_coordinates / interp_to_coords (core.py:1050-1174,
// 1405-1434), cal_gradient_wrt_area (core.py:463-488),
// cal_sqared_equivalent_length (core.py:635), cal_normalized_Keff
// (core.py:963-964), latitude_lengths_at / equivalent_latitudes
// (utils.py:491-534).  These are latency-, not bandwidth-bound; one thread per
// output element, the per-slice tables read through L1/L2.
#include "common.cuh"
#include "internal.h"
#include <math_constants.h>

namespace xc {

constexpr double kRearth = 6371200.0;                    // utils.py:19
constexpr double kPi = 3.141592653589793;

// reverse: 0 = ascending tables, 1 = evaluate on reversed tables, -1 = decide
// from slice 0 the way the reference does (core.py:1080-1088, 1122-1126):
// increasing iff xp[0][0] < xp[0][n-1].
__global__ void k_interp(const double* __restrict__ x, long x_stride, int M,
                         const double* __restrict__ xp, long xp_stride,
                         const double* __restrict__ fp, long fp_stride, int n,
                         int reverse, const double* __restrict__ probe,
                         double* __restrict__ out)
{
    const long s = blockIdx.y;
    bool rev = reverse == 1;
    if (reverse < 0) rev = !(probe[0] < probe[n - 1]);
    const double* xs = x + s * x_stride;
    const double* xps = xp + s * xp_stride;
    const double* fps = fp + s * fp_stride;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x)
        out[s * (long)M + m] = np_interp(xs[m], xps, fps, n, rev);
}

// np.gradient along the contour axis with unit spacing (edge_order = 1):
// interior (f[k+1]-f[k-1])/2, one-sided first differences at the ends, in f's
// own dtype.
// kind: XC_F32 (float array), XC_F64 (double array), XC_F32_AS_F64 (double
// storage of fp32-representable values, differenced in fp32).
__device__ __forceinline__ double grad_at(const void* f, int kind, long base, int k, int N)
{
    const int km = k == 0 ? 0 : k - 1, kp = k == N - 1 ? N - 1 : k + 1;
    const bool interior = (k > 0) && (k < N - 1);
    if (kind == XC_F64) {
        const double* p = (const double*)f + base;
        double d = __dsub_rn(p[kp], p[km]);
        return interior ? __ddiv_rn(d, 2.0) : d;
    }
    float a, b;
    if (kind == XC_F32) { const float* p = (const float*)f + base; a = p[kp]; b = p[km]; }
    else { const double* p = (const double*)f + base; a = (float)p[kp]; b = (float)p[km]; }
    float d = __fsub_rn(a, b);
    return (double)(interior ? __fdiv_rn(d, 2.0f) : d);
}

__global__ void k_gradient_wrt_area(const void* __restrict__ var, int var_kind,
                                    const void* __restrict__ area, int area_kind,
                                    long total, int N, double* __restrict__ out)
{
    const bool both32 = (var_kind != XC_F64) && (area_kind != XC_F64);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
         i += (long)gridDim.x * blockDim.x) {
        const long s = i / N; const int k = (int)(i - s * N);
        double dv = grad_at(var, var_kind, s * N, k, N);
        double da = grad_at(area, area_kind, s * N, k, N);
        out[i] = both32 ? (double)__fdiv_rn((float)dv, (float)da) : __ddiv_rn(dv, da);
    }
}

// np.gradient against an explicit (possibly non-uniform) contour coordinate, edge_order = 1.  The coefficients
// come from NumPy itself on the host (numpy/lib/_function_base_impl.py, gradient): uniform spacing ->
// coef = {2*dx, dx}: interior (f[k+1]-f[k-1]) / (2*dx), ends (f[1]-f[0]) / dx; otherwise
// coef = {a[N-2], b[N-2], c[N-2], dx_0, dx_n}: interior a*f[k-1] + b*f[k] + c*f[k+1].  `c32`: the coefficients are
// fp32 values and f is fp32 -> arithmetic in fp32, else in fp64; the result is rounded to f's dtype (np.gradient
// allocates its output in f.dtype).
__device__ __forceinline__ double grad_coord_at(const void* f, int kind, long base, int k, int N,
                                                const double* coef, int uniform, int c32)
{
    auto F = [&](int i) -> double {
        return kind == XC_F32 ? (double)((const float*)f)[base + i] : ((const double*)f)[base + i];
    };
    const bool f32 = kind != XC_F64;
    const bool in32 = f32 && c32;
    double r;
    if (k == 0 || k == N - 1) {
        const double dx = uniform ? coef[1] : coef[3 * (N - 2) + (k == 0 ? 0 : 1)];
        const double hi = F(k == 0 ? 1 : N - 1), lo = F(k == 0 ? 0 : N - 2);
        if (f32) { const float d = __fsub_rn((float)hi, (float)lo); r = in32 ? (double)__fdiv_rn(d, (float)dx) : __ddiv_rn((double)d, dx); }
        else r = __ddiv_rn(__dsub_rn(hi, lo), dx);
    } else if (uniform) {
        if (f32) { const float d = __fsub_rn((float)F(k + 1), (float)F(k - 1)); r = in32 ? (double)__fdiv_rn(d, (float)coef[0]) : __ddiv_rn((double)d, coef[0]); }
        else r = __ddiv_rn(__dsub_rn(F(k + 1), F(k - 1)), coef[0]);
    } else {
        const double a = coef[k - 1], b = coef[(N - 2) + k - 1], c = coef[2 * (N - 2) + k - 1];
        if (in32) r = (double)__fadd_rn(__fadd_rn(__fmul_rn((float)a, (float)F(k - 1)), __fmul_rn((float)b, (float)F(k))), __fmul_rn((float)c, (float)F(k + 1)));
        else r = __dadd_rn(__dadd_rn(__dmul_rn(a, F(k - 1)), __dmul_rn(b, F(k))), __dmul_rn(c, F(k + 1)));
    }
    return f32 ? (double)(float)r : r;
}

__global__ void k_gradient_wrt_area_coord(const void* __restrict__ var, int var_kind, const double* __restrict__ vcoef,
                                          int v_uniform, int v_c32,
                                          const void* __restrict__ area, int area_kind, const double* __restrict__ acoef,
                                          int a_uniform, int a_c32, long total, int N, double* __restrict__ out)
{
    const bool both32 = (var_kind != XC_F64) && (area_kind != XC_F64);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long s = i / N; const int k = (int)(i - s * N);
        const double dv = grad_coord_at(var, var_kind, s * N, k, N, vcoef, v_uniform, v_c32);
        const double da = grad_coord_at(area, area_kind, s * N, k, N, acoef, a_uniform, a_c32);
        out[i] = both32 ? (double)__fdiv_rn((float)dv, (float)da) : __ddiv_rn(dv, da);
    }
}

__global__ void k_leq2(const double* a, const double* b, long n, double* out)
{
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = __ddiv_rn(a[i], __dmul_rn(b[i], b[i]));
}

__global__ void k_lmin(const double* lat, long n, double* out)
{
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        double rad = __dmul_rn(lat[i], kPi / 180.0);              // np.deg2rad
        out[i] = __dmul_rn(__dmul_rn(__dmul_rn(2.0, kPi), kRearth), cos(rad));
    }
}

__global__ void k_nkeff(const double* leq2, const double* lmin, double mask, long n, double* out)
{
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        double v = __ddiv_rn(__ddiv_rn(leq2[i], lmin[i]), lmin[i]);
        out[i] = (v < mask) ? v : CUDART_NAN;
    }
}

__global__ void k_eqlat(const double* area, long n, double* out)
{
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        double r = __dsub_rn(__ddiv_rn(__ddiv_rn(__ddiv_rn(__ddiv_rn(area[i], 2.0), kPi), kRearth), kRearth), 1.0);
        if (r < -1.0) r = -1.0;
        if (r > 1.0) r = 1.0;
        out[i] = __dmul_rn(asin(r), 180.0 / kPi);                 // np.rad2deg
    }
}

static inline unsigned ew_blocks(long n)
{
    long b = (n + 255) / 256;
    long cap = (long)sm_count() * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace xc

using namespace xc;

extern "C" int xc_interp(const double* x, long x_stride, int M,
                         const double* xp, long xp_stride,
                         const double* fp, long fp_stride, int n, int reverse,
                         long S, double* out, void* stream)
{
    XC_REQUIRE(x && xp && fp && out, "xc_interp: null pointer");
    XC_REQUIRE(S > 0 && M > 0 && n > 0, "xc_interp: need S>0, M>0, n>0");
    for (long s0 = 0; s0 < S; s0 += 65535) {
        long ns = S - s0 < 65535 ? S - s0 : 65535;
        dim3 grid((unsigned)((M + 127) / 128), (unsigned)ns);
        // the direction probe (reverse < 0) always looks at slice 0 of the call
        k_interp<<<grid, 128, 0, (cudaStream_t)stream>>>(
            x + s0 * x_stride, x_stride, M, xp + s0 * xp_stride, xp_stride,
            fp + s0 * fp_stride, fp_stride, n, reverse, xp, out + s0 * (long)M);
        XC_LAUNCH_OK();
    }
    return 0;
}

extern "C" int xc_gradient_wrt_area(const void* var, int var_dtype,
                                    const void* area, int area_dtype,
                                    long S, int N, double* out, void* stream)
{
    XC_REQUIRE(var && area && out, "xc_gradient_wrt_area: null pointer");
    XC_REQUIRE(S > 0 && N >= 2, "xc_gradient_wrt_area: need S>0, N>=2");
    const long total = S * (long)N;
    k_gradient_wrt_area<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(
        var, var_dtype, area, area_dtype, total, N, out);
    XC_LAUNCH_OK();
    return 0;
}

extern "C" int xc_gradient_wrt_area_coord(const void* var, int var_dtype, const double* var_coef, int var_uniform, int var_coef_f32,
                                          const void* area, int area_dtype, const double* area_coef, int area_uniform, int area_coef_f32,
                                          long S, int N, double* out, void* stream)
{
    XC_REQUIRE(var && area && out && var_coef && area_coef, "xc_gradient_wrt_area_coord: null pointer");
    XC_REQUIRE(S > 0 && N >= 2, "xc_gradient_wrt_area_coord: need S>0, N>=2");
    XC_REQUIRE((var_dtype == XC_F32 || var_dtype == XC_F64) && (area_dtype == XC_F32 || area_dtype == XC_F64),
               "xc_gradient_wrt_area_coord: bad dtype");
    const long total = S * (long)N;
    k_gradient_wrt_area_coord<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(
        var, var_dtype, var_coef, var_uniform, var_coef_f32, area, area_dtype, area_coef, area_uniform, area_coef_f32,
        total, N, out);
    XC_LAUNCH_OK();
    return 0;
}

extern "C" int xc_leq2(const double* a, const double* b, long n, double* out, void* stream)
{
    XC_REQUIRE(a && b && out && n > 0, "xc_leq2: bad arguments");
    k_leq2<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(a, b, n, out);
    XC_LAUNCH_OK();
    return 0;
}
extern "C" int xc_lmin(const double* lat, long n, double* out, void* stream)
{
    XC_REQUIRE(lat && out && n > 0, "xc_lmin: bad arguments");
    k_lmin<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(lat, n, out);
    XC_LAUNCH_OK();
    return 0;
}
extern "C" int xc_nkeff(const double* leq2, const double* lmin, double mask, long n, double* out, void* stream)
{
    XC_REQUIRE(leq2 && lmin && out && n > 0, "xc_nkeff: bad arguments");
    k_nkeff<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(leq2, lmin, mask, n, out);
    XC_LAUNCH_OK();
    return 0;
}
extern "C" int xc_eqlat(const double* area, long n, double* out, void* stream)
{
    XC_REQUIRE(area && out && n > 0, "xc_eqlat: bad arguments");
    k_eqlat<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(area, n, out);
    XC_LAUNCH_OK();
    return 0;
}
