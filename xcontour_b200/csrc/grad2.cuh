// Device-side definition of the |grad q|^2 stencil, shared by the standalone
// kernel (grad2.cu) and the binning kernel (hist.cu, in-flight integrand).
#pragma once
#include "common.cuh"

namespace xc {

constexpr double kRearthG = 6371200.0;   // xcontour/utils.py:19

// Row metrics (one pair per latitude row j), fp64:
//   cx[j] = 1 / ((2 dlambda) * (R cos phi_j))
//   cy[j] = 1 / ((phi_{j+1} - phi_{j-1}) * R)        one-sided at the first/last row
__device__ __forceinline__ void grad2_row_metrics(const double* __restrict__ lat_rad, int j, int ny,
                                                  double two_dlam, double& cx, double& cy)
{
    const int jm = j == 0 ? 0 : j - 1, jp = j == ny - 1 ? ny - 1 : j + 1;
    cx = __ddiv_rn(1.0, __dmul_rn(two_dlam, __dmul_rn(kRearthG, cos(lat_rad[j]))));
    cy = __ddiv_rn(1.0, __dmul_rn(__dsub_rn(lat_rad[jp], lat_rad[jm]), kRearthG));
}

// (dq/dx)^2 + (dq/dy)^2 from the four neighbours, every operation individually
// rounded (no FMA contraction) so the value is bit-identical to the NumPy
// statement in oracle/xcontour_oracle.py:squared_gradient_latlon.
__device__ __forceinline__ double grad2_from(double qe, double qw, double qn, double qso,
                                             double cx, double cy)
{
    const double dqdx = __dmul_rn(__dsub_rn(qe, qw), cx);
    const double dqdy = __dmul_rn(__dsub_rn(qn, qso), cy);
    return __dadd_rn(__dmul_rn(dqdx, dqdx), __dmul_rn(dqdy, dqdy));
}

template <typename QT>
__device__ __forceinline__ double grad2_cell(const QT* __restrict__ qs, int j, int i, int ny, int nx,
                                             double cx, double cy)
{
    const int im = i == 0 ? nx - 1 : i - 1, ip = i == nx - 1 ? 0 : i + 1;
    const int jm = j == 0 ? 0 : j - 1, jp = j == ny - 1 ? ny - 1 : j + 1;
    return grad2_from((double)__ldg(qs + (long)j * nx + ip), (double)__ldg(qs + (long)j * nx + im),
                      (double)__ldg(qs + (long)jp * nx + i), (double)__ldg(qs + (long)jm * nx + i), cx, cy);
}

}  // namespace xc
