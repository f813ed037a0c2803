// Device-side definition of the |grad q|^2 stencil, shared by the standalone
// kernel (grad2.cu) and the fused Keff+LWA batch (fused.cu).
#pragma once
#include "common.cuh"

namespace xc {

constexpr double kRearthG = 6371200.0;   // xcontour/utils.py:19

// (dq/dx)^2 + (dq/dy)^2 at cell (j, i) of one slice.
//   dq/dx = (q[j][i+1]-q[j][i-1]) / (2 dlambda) / (R cos phi_j)      periodic in i
//   dq/dy = (q[j+1][i]-q[j-1][i]) / (phi_{j+1}-phi_{j-1}) / R        one-sided at the ends
// Every operation is individually rounded (no FMA contraction) so the value is
// bit-identical to the NumPy statement in oracle/xcontour_oracle.py.
template <typename QT>
__device__ __forceinline__ double grad2_cell(const QT* __restrict__ qs, int j, int i, int ny, int nx,
                                             double rcos, double dphi, double two_dlam)
{
    const int im = i == 0 ? nx - 1 : i - 1, ip = i == nx - 1 ? 0 : i + 1;
    const int jm = j == 0 ? 0 : j - 1, jp = j == ny - 1 ? ny - 1 : j + 1;
    const double qe = (double)__ldg(qs + (long)j * nx + ip), qw = (double)__ldg(qs + (long)j * nx + im);
    const double qn = (double)__ldg(qs + (long)jp * nx + i), qso = (double)__ldg(qs + (long)jm * nx + i);
    const double dqdx = __ddiv_rn(__ddiv_rn(__dsub_rn(qe, qw), two_dlam), rcos);
    const double dqdy = __ddiv_rn(__ddiv_rn(__dsub_rn(qn, qso), dphi), kRearthG);
    return __dadd_rn(__dmul_rn(dqdx, dqdx), __dmul_rn(dqdy, dqdy));
}

}  // namespace xc
