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

}  // namespace xc

using namespace xc;

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
