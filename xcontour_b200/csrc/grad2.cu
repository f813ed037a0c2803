// Kernel (3): |grad q|^2 on a regular lat-lon grid (centred differences,
// periodic in longitude, one-sided at the first/last latitude).  The reference
// has no such routine -- its callers import the field from xinvert / GeoApps
//  -- so the definition is the one stated in
// oracle/xcontour_oracle.py:squared_gradient_latlon.  Each CTA handles one row
// segment; the three rows it touches stream through L1/L2 (every HBM sector of q
// is fetched once per slice), the result is written with coalesced stores.
#include "common.cuh"
#include "grad2.cuh"
#include "internal.h"

namespace xc {

__global__ void k_row_metrics(const double* __restrict__ lat_rad, int ny, double two_dlam,
                              double* __restrict__ cx, double* __restrict__ cy)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ny; j += gridDim.x * blockDim.x)
        grad2_row_metrics(lat_rad, j, ny, two_dlam, cx[j], cy[j]);
}

template <typename QT, typename OT>
__global__ void __launch_bounds__(256)
k_grad2(const QT* __restrict__ q, int ny, int nx, const double* __restrict__ lat_rad,
        double dlam, OT* __restrict__ out)
{
    const long s = blockIdx.z; const int j = blockIdx.y;
    __shared__ double m[2];
    if (threadIdx.x == 0) grad2_row_metrics(lat_rad, j, ny, __dmul_rn(2.0, dlam), m[0], m[1]);
    __syncthreads();
    const QT* qs = q + s * (long)ny * nx;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x) {
        double g = grad2_cell(qs, j, i, ny, nx, m[0], m[1]);
        out[(s * ny + j) * (long)nx + i] = (OT)g;
    }
}

template <typename OT>
__global__ void k_latlon_cell_area(const double* __restrict__ lat_deg, int ny, int nx, double dlam_deg,
                                   OT* __restrict__ out)
{
    const int j = blockIdx.y;
    __shared__ double band;
    if (threadIdx.x == 0) {
        const bool asc = lat_deg[ny - 1] > lat_deg[0];
        // work on the ascending view la[k]; row j is element k of it
        const int k = asc ? j : ny - 1 - j;
        auto la = [&](int m) { return asc ? lat_deg[m] : lat_deg[ny - 1 - m]; };
        double lo = (k == 0) ? fmax(-90.0, la(0) - 0.5 * (la(1) - la(0))) : 0.5 * (la(k) + la(k - 1));
        double hi = (k == ny - 1) ? fmin(90.0, la(ny - 1) + 0.5 * (la(ny - 1) - la(ny - 2))) : 0.5 * (la(k + 1) + la(k));
        const double d2r = 3.141592653589793 / 180.0;
        band = kRearthG * kRearthG * (sin(hi * d2r) - sin(lo * d2r)) * (dlam_deg * d2r);
    }
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x)
        out[(long)j * nx + i] = (OT)band;
}

}  // namespace xc

using namespace xc;

extern "C" int xc_latlon_cell_area(const double* lat_deg, int n_y, int n_x, double dlambda_deg,
                                   void* out, int out_dtype, void* stream)
{
    XC_REQUIRE(lat_deg && out, "xc_latlon_cell_area: null pointer");
    XC_REQUIRE(n_y >= 2 && n_x >= 1 && n_y <= 65535 && dlambda_deg > 0.0, "xc_latlon_cell_area: bad sizes");
    dim3 grid((unsigned)((n_x + 255) / 256), (unsigned)n_y);
    if (out_dtype == XC_F32) k_latlon_cell_area<float><<<grid, 256, 0, (cudaStream_t)stream>>>(lat_deg, n_y, n_x, dlambda_deg, (float*)out);
    else                     k_latlon_cell_area<double><<<grid, 256, 0, (cudaStream_t)stream>>>(lat_deg, n_y, n_x, dlambda_deg, (double*)out);
    XC_LAUNCH_OK();
    return 0;
}

int xc::row_metrics(const double* lat_rad, int ny, double dlambda, double* cx, double* cy, void* stream)
{
    k_row_metrics<<<(ny + 255) / 256, 256, 0, (cudaStream_t)stream>>>(lat_rad, ny, 2.0 * dlambda, cx, cy);
    XC_LAUNCH_OK();
    return 0;
}

extern "C" int xc_grad2_latlon(const void* q, int q_dtype, long S, int n_y, int n_x,
                               const double* lat_rad, double dlambda,
                               void* out, int out_dtype, void* stream)
{
    XC_REQUIRE(q && lat_rad && out, "xc_grad2_latlon: null pointer");
    XC_REQUIRE(S > 0 && n_y >= 2 && n_x >= 2, "xc_grad2_latlon: need S>0, n_y>=2, n_x>=2");
    XC_REQUIRE(n_y <= 65535, "xc_grad2_latlon: n_y too large");
    cudaStream_t st = (cudaStream_t)stream;
    for (long s0 = 0; s0 < S; s0 += 65535) {
        long ns = S - s0 < 65535 ? S - s0 : 65535;
        dim3 grid((unsigned)((n_x + 255) / 256), (unsigned)n_y, (unsigned)ns);
        const long off = s0 * (long)n_y * n_x;
        if (q_dtype == XC_F32 && out_dtype == XC_F32)
            k_grad2<float, float><<<grid, 256, 0, st>>>((const float*)q + off, n_y, n_x, lat_rad, dlambda, (float*)out + off);
        else if (q_dtype == XC_F32)
            k_grad2<float, double><<<grid, 256, 0, st>>>((const float*)q + off, n_y, n_x, lat_rad, dlambda, (double*)out + off);
        else if (out_dtype == XC_F32)
            k_grad2<double, float><<<grid, 256, 0, st>>>((const double*)q + off, n_y, n_x, lat_rad, dlambda, (float*)out + off);
        else
            k_grad2<double, double><<<grid, 256, 0, st>>>((const double*)q + off, n_y, n_x, lat_rad, dlambda, (double*)out + off);
        XC_LAUNCH_OK();
    }
    return 0;
}
