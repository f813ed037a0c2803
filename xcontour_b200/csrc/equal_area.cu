// Equal-area contour levels by a weighted-quantile histogram (north_star kernel (1); an extension of the
// reference, which only offers equally spaced levels -- core.py:205-266 -- and levels at prescribed equivalent
// coordinates -- core.py:316-360).  One call, no host round trips:
//   min/max -> (N-1)*refine+1 equally spaced fine levels and their bin edges      (k_minmax_*, k_levels)
//   -> area CDF of the fine levels, the histogram path of core.py:412-460           (bin_accumulate_impl)
//   -> the A(q) relation inverted by np.interp at N equally spaced areas, rounded
//      once to the contour dtype                                                    (k_equal_area_invert)
// The arithmetic is the oracle's cal_contours_equal_area statement by statement.
#include "common.cuh"
#include "internal.h"

namespace xc {

// grid = S, one CTA per slice.  area / fine: [S][Nf] fp64 (contour index ascending).
__global__ void __launch_bounds__(256)
k_equal_area_invert(const double* __restrict__ area, const double* __restrict__ fine, int Nf, int N,
                    int out_is_f32, double* __restrict__ levels)
{
    extern __shared__ __align__(16) unsigned char smem[];
    double* sa = reinterpret_cast<double*>(smem);
    double* sf = sa + Nf;
    const long s = blockIdx.x;
    for (int k = threadIdx.x; k < Nf; k += blockDim.x) { sa[k] = area[s * Nf + k]; sf[k] = fine[s * Nf + k]; }
    // direction of the relation: detected once, on the first slice (as the reference detects the direction of
    // every per-slice interpolation, core.py:1080-1088)
    const bool inc = area[0] < area[Nf - 1];
    __syncthreads();
    const double a0 = sa[0], span = __dsub_rn(sa[Nf - 1], sa[0]);
    const double wstep = __ddiv_rn(1.0, (double)(N - 1));                 // np.linspace(0, 1, N): k * step, last = stop
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const double w = (k == N - 1) ? 1.0 : __dmul_rn((double)k, wstep);
        const double tgt = __dadd_rn(a0, __dmul_rn(span, w));
        double v = np_interp(tgt, sa, sf, Nf, !inc);
        if (out_is_f32) v = (double)__double2float_rn(v);
        levels[s * N + k] = v;
    }
}

}  // namespace xc

using namespace xc;

namespace {
struct EqaPlan { int Nf; size_t ws_minmax, ws_hist, total; };
EqaPlan eqa_plan(long S, long P, int N, int refine)
{
    EqaPlan p;
    p.Nf = (N - 1) * refine + 1;
    p.ws_minmax = xc_minmax_levels_workspace_bytes(S, P);
    p.ws_hist = xc_bin_accumulate_workspace_bytes(S, P, p.Nf, 1);
    size_t t = align_up(p.ws_minmax, 256) + align_up(p.ws_hist, 256);
    t += 2 * align_up((size_t)S * p.Nf * 8, 256);                       // fine levels, area CDF
    t += align_up((size_t)S * (p.Nf + 1) * 8, 256);                     // edges
    t += align_up((size_t)S * 16, 256) + 2 * align_up((size_t)S * 4, 256);   // (min, max), decreasing, flag
    p.total = t + 2048;
    return p;
}
}  // namespace

extern "C" size_t xc_equal_area_levels_workspace_bytes(long S, long P, int N, int refine)
{
    if (S <= 0 || P <= 0 || N < 2 || refine < 1) return 0;
    return eqa_plan(S, P, N, refine).total;
}

extern "C" int xc_equal_area_levels(const void* q, int q_dtype, long S, long P,
                                    const void* dA, int dA_dtype,
                                    int N, int refine, int increase, int lt, int out_dtype, int numpy2_rules,
                                    double* levels, void* workspace, size_t ws_bytes, void* stream)
{
    XC_REQUIRE(q && dA && levels, "xc_equal_area_levels: null pointer");
    XC_REQUIRE(S > 0 && P > 0 && N >= 2 && refine >= 1, "xc_equal_area_levels: need S>0, P>0, N>=2, refine>=1");
    XC_REQUIRE(q_dtype == XC_F32 || q_dtype == XC_F64, "xc_equal_area_levels: bad q dtype");
    const EqaPlan pl = eqa_plan(S, P, N, refine);
    XC_REQUIRE((size_t)pl.Nf * 16 <= 200 * 1024, "xc_equal_area_levels: (N-1)*refine+1 = %d fine levels do not fit shared memory", pl.Nf);
    XC_REQUIRE(workspace && ws_bytes >= pl.total, "xc_equal_area_levels: workspace too small (%zu < %zu)", ws_bytes, pl.total);
    Arena ar(workspace, ws_bytes);
    char* w_minmax = ar.take<char>(pl.ws_minmax);
    char* w_hist = ar.take<char>(pl.ws_hist);
    double* fine = ar.take<double>((size_t)S * pl.Nf);
    double* area = ar.take<double>((size_t)S * pl.Nf);
    double* edges = ar.take<double>((size_t)S * (pl.Nf + 1));
    double* minmax = ar.take<double>((size_t)S * 2);
    int32_t* decr = ar.take<int32_t>((size_t)S);
    int32_t* flag = ar.take<int32_t>(1);
    XC_REQUIRE(ar.ok(), "xc_equal_area_levels: workspace accounting error");
    // bins that vary per slice take the reference's per-'time' edge rules (core.py:1273-1281); a single slice is the
    // static branch (core.py:1296-1304), whose edges keep the contour dtype in either NumPy regime
    const int keep_ctr_dtype = (numpy2_rules || S == 1) ? 1 : 0;
    if (minmax_levels_impl(q, q_dtype, S, P, pl.Nf, increase, out_dtype, fine, minmax, edges, decr, flag,
                           w_minmax, pl.ws_minmax, stream, keep_ctr_dtype)) return 1;
    ScanOut so; so.p[0] = area; so.p[1] = so.p[2] = so.p[3] = nullptr; so.stride = pl.Nf;
    if (bin_accumulate_impl(q, q_dtype, S, P, edges, pl.Nf + 1, pl.Nf, 0, dA, dA_dtype, 1,
                            nullptr, nullptr, 0, nullptr, lt ? XC_SCAN_PREFIX : XC_SCAN_TOTAL_MINUS, decr,
                            nullptr, so, nullptr, w_hist, pl.ws_hist, stream)) return 1;
    const size_t sm = (size_t)pl.Nf * 16;
    if (sm > 48 * 1024)
        XC_CUDA_OK(cudaFuncSetAttribute(k_equal_area_invert, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_equal_area_invert<<<(unsigned)S, 256, sm, (cudaStream_t)stream>>>(area, fine, pl.Nf, N, out_dtype == XC_F32, levels);
    XC_LAUNCH_OK();
    return 0;
}
