// Internal (non-ABI) declarations shared between the kernel files.
#pragma once
#include "common.cuh"

#define XC_F32_AS_F64 2   /* fp64 storage of fp32-representable values; arithmetic in fp32 */

namespace xc {

// where the scan kernel writes the CDF of accumulator k: p[k][s*stride + n]
struct ScanOut { double* p[4]; long stride; };

// optional in-flight |grad q|^2 integrand for bin_accumulate_impl (adds one
// accumulator after the explicit integrands); rcos/dphi from row_metrics().
// cx/cy: [ny] row metrics, dq/dx = (q[i+1]-q[i-1])*cx[j], dq/dy = (q[j+1]-q[j-1])*cy[j]; ghost cells by bcx/bcy
// (XC_BC_*).  dA_row (nullable): [ny] fp64 cell area of each row when dA is constant along rows -- enables the
// row-march binning kernel (bin_rows.cu); uniform_dA: all rows equal; any_degenerate: reserve the pole-row
// accumulator; minmax (nullable): [S][2] NaN-skipping (min, max) of each slice of this call.
struct StencilArgs {
    int ny, nx; const double* cx; const double* cy;
    int bcx = XC_BC_PERIODIC, bcy = XC_BC_EXTEND; float fill = 0.f;
    const double* dA_row = nullptr; int uniform_dA = 0, any_degenerate = 1;
    const double* minmax = nullptr;
};
int row_metrics(const double* lat_rad, int ny, double dlambda, double* cx, double* cy, void* stream);

struct HistOnly;
int bin_accumulate_impl(const void* q, int q_dtype, long S, long P,
                        const double* edges, long edges_stride, int N,
                        int closed_right,
                        const void* dA, int dA_dtype, int acc_area,
                        const void* const* integrands, const int* integrand_dtypes,
                        int n_int, const uint8_t* q_mask,
                        int scan_mode, const int32_t* decreasing,
                        double* pdf, const ScanOut& so, int32_t* bin_idx,
                        void* workspace, size_t ws_bytes, void* stream,
                        const StencilArgs* stencil = nullptr, struct HistOnly* hist_only = nullptr);
// workspace of bin_accumulate_impl when the caller passes stencil (ny, nx known) and hist_only
size_t bin_accumulate_ws_bytes_stencil(long S, int ny, int nx, int N);

// when passed to bin_accumulate_impl the scan kernel is skipped and the caller
// gets the per-CTA partials [S][C][K][N] (the fused epilogue reduces them itself)
struct HistOnly { const double* part; int C; };

int scan_epilogue(const double* part, int C, long S, int N, int lt, const int32_t* decreasing,
                  const double* ctr, int ctr_f32,
                  const double* table, const double* table_coord, int n_table,
                  const double* eq_coord, int ny, double keff_mask, int increase,
                  double* area, double* intg, double* latEq, double* Lmin, double* dint,
                  double* dq, double* Leq2, double* nkeff, double* Qref,
                  int32_t* sorted, int32_t* any_unsorted, void* stream, double* reduce_buf = nullptr);

// row-march fused-Keff binning kernel (bin_rows.cu): 0 launched (*C_out CTAs per slice), 1 not applicable, 2 error
int bin_rows_try(const void* q, int q_dtype, long S, const double* edges, int N,
                 const StencilArgs* st, const double* minmax, double* part, size_t part_doubles,
                 int* C_out, void* stream);

size_t bin_rows_part_doubles(long S, int ny, int nx, int N);

// levels (+ optional per-'time'-branch edges in the same launch); clears *flag_to_clear
int minmax_levels_impl(const void* q, int q_dtype, long S, long P, int N, int increase, int out_dtype,
                       double* levels, double* minmax, double* edges, int32_t* decreasing,
                       int32_t* flag_to_clear, void* workspace, size_t ws_bytes, void* stream,
                       int edges_keep_ctr_dtype = 0);

// LWA with sortedness flags already on the device (fused path)
// minmax: NaN-skipping (min, max) per slice [S][2] if the caller has them (else they are
// computed into `scratch`); scratch: lwa_scratch_doubles(S, minmax != nullptr) doubles.
// The fixed-point kernel hands slices with non-finite values to the exact loop by
// clearing sorted[s] and raising *any_unsorted.
size_t lwa_scratch_doubles(long S, bool have_minmax);
int lwa_impl(const void* q, int q_dtype, long S, int n_eq, int n_x, const double* Qref, const double* ww,
             int increase, int part, int variant, double* out, int32_t* sorted,
             int32_t* any_unsorted, bool flags_ready, const double* minmax, double* scratch, void* stream,
             const double* wmax_ready = nullptr, const double* ww_row = nullptr, int out_f32 = 0);
// partial max |ww| (lwa_wmax_doubles() values) for wmax_ready: the fused batch computes them once per call
size_t lwa_wmax_doubles();
int lwa_wmax(const double* ww, long P, double* parts, void* stream);

}  // namespace xc
