// Internal (non-ABI) declarations shared between the kernel files.
#pragma once
#include "common.cuh"

#define XC_F32_AS_F64 2   /* fp64 storage of fp32-representable values; arithmetic in fp32 */

namespace xc {

// where the scan kernel writes the CDF of accumulator k: p[k][s*stride + n]
struct ScanOut { double* p[4]; long stride; };

// optional in-flight |grad q|^2 integrand for bin_accumulate_impl (adds one
// accumulator after the explicit integrands); rcos/dphi from row_metrics().
struct StencilArgs { int ny, nx; const double* cx; const double* cy; };
int row_metrics(const double* lat_rad, int ny, double dlambda, double* cx, double* cy, void* stream);

int bin_accumulate_impl(const void* q, int q_dtype, long S, long P,
                        const double* edges, long edges_stride, int N,
                        int closed_right,
                        const void* dA, int dA_dtype, int acc_area,
                        const void* const* integrands, const int* integrand_dtypes,
                        int n_int, const uint8_t* q_mask,
                        int scan_mode, const int32_t* decreasing,
                        double* pdf, const ScanOut& so, int32_t* bin_idx,
                        void* workspace, size_t ws_bytes, void* stream,
                        const StencilArgs* stencil = nullptr);

}  // namespace xc
