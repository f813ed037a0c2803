// (8) Fused batch: Keff + LWA for a batch of slices, chained on one stream with
// no host round-trips.  Every stage is one of the kernels the stand-alone ABI
// entries launch (same arithmetic, same parity); the intermediate contour-space
// arrays live in the caller's workspace.  This is the call bench.py times.
#include "common.cuh"
#include "internal.h"

using namespace xc;

namespace {
struct FusedPlan {
    size_t ws_minmax, ws_hist, ws_lwa;
    size_t total;
};
FusedPlan fused_plan(long S, int ny, int nx, int N, bool need_grd)
{
    FusedPlan p;
    const long P = (long)ny * nx;
    p.ws_minmax = xc_minmax_levels_workspace_bytes(S, P);
    p.ws_hist = xc_bin_accumulate_workspace_bytes(S, P, N, 2);
    p.ws_lwa = xc_lwa_workspace_bytes(S);
    size_t t = 0;
    t += align_up(p.ws_minmax, 256) + align_up(p.ws_hist, 256) + align_up(p.ws_lwa, 256);
    t += align_up((size_t)S * (N + 1) * 8, 256);            // edges
    t += align_up((size_t)S * 4, 256);                      // decreasing
    t += 10 * align_up((size_t)S * N * 8, 256);             // contour-space temporaries
    t += align_up((size_t)S * ny * 8, 256);                 // Qref
    if (need_grd) t += align_up((size_t)S * P * 8, 256);    // |grad q|^2 (fp64)
    p.total = t + 4096;
    return p;
}
}  // namespace

extern "C" size_t xc_keff_lwa_batch_workspace_bytes(long S, int n_y, int n_x, int N)
{
    if (S <= 0 || n_y <= 0 || n_x <= 0 || N <= 0) return 0;
    return fused_plan(S, n_y, n_x, N, true).total;
}

extern "C" int xc_keff_lwa_batch(const xc_keff_lwa_args* a, void* workspace, size_t ws_bytes, void* stream)
{
    XC_REQUIRE(a, "xc_keff_lwa_batch: null args");
    XC_REQUIRE(a->q && a->dA && a->table && a->table_coord && a->eq_coord && a->ww,
               "xc_keff_lwa_batch: null input pointer");
    XC_REQUIRE(a->S > 0 && a->n_y >= 2 && a->n_x >= 2 && a->N >= 2 && a->n_table >= 1,
               "xc_keff_lwa_batch: bad sizes");
    XC_REQUIRE(a->grdS || a->lat_rad, "xc_keff_lwa_batch: need grdS or lat_rad for the stencil");
    const long S = a->S; const int ny = a->n_y, nx = a->n_x, N = a->N;
    const long P = (long)ny * nx;
    const bool need_grd = a->grdS == nullptr;
    FusedPlan pl = fused_plan(S, ny, nx, N, need_grd);
    XC_REQUIRE(workspace && ws_bytes >= pl.total, "xc_keff_lwa_batch: workspace too small (%zu < %zu)",
               ws_bytes, pl.total);
    Arena ar(workspace, ws_bytes);
    char* w_minmax = ar.take<char>(pl.ws_minmax);
    char* w_hist = ar.take<char>(pl.ws_hist);
    char* w_lwa = ar.take<char>(pl.ws_lwa);
    double* edges = ar.take<double>((size_t)S * (N + 1));
    int32_t* decr = ar.take<int32_t>((size_t)S);
    auto tmp = [&](double* user) { double* t = ar.take<double>((size_t)S * N); return user ? user : t; };
    double* ctr = tmp(a->ctr);       double* area = tmp(a->area);   double* intg = tmp(a->intgrdS);
    double* latEq = tmp(a->latEq);   double* Lmin = tmp(a->Lmin);   double* dintSdA = tmp(a->dintSdA);
    double* dqdA = tmp(a->dqdA);     double* Leq2 = tmp(a->Leq2);   double* nkeff = tmp(a->nkeff);
    double* Qref = ar.take<double>((size_t)S * ny);
    if (a->Qref) Qref = a->Qref;
    double* grd = need_grd ? ar.take<double>((size_t)S * P) : nullptr;
    XC_REQUIRE(ar.ok(), "xc_keff_lwa_batch: workspace accounting error");

    // (1) levels, (1b) edges -- per-slice contours take the per-'time' branch
    if (xc_minmax_levels(a->q, a->q_dtype, S, P, N, a->increase, a->ctr_dtype, ctr, nullptr,
                         w_minmax, pl.ws_minmax, stream)) return 1;
    if (xc_hist_edges(ctr, S, N, a->ctr_dtype, 1, edges, decr, stream)) return 1;
    // (7) integrand
    const void* integ = a->grdS; int integ_dtype = a->grdS_dtype;
    if (need_grd) {
        if (xc_grad2_latlon(a->q, a->q_dtype, S, ny, nx, a->lat_rad, a->dlambda, grd, XC_F64, stream)) return 1;
        integ = grd; integ_dtype = XC_F64;
    }
    // (2) area and int |grad q|^2 dA in one pass over q
    ScanOut so; so.p[0] = area; so.p[1] = intg; so.p[2] = so.p[3] = nullptr; so.stride = N;
    const void* integs[1] = { integ };
    if (bin_accumulate_impl(a->q, a->q_dtype, S, P, edges, N + 1, N, 0, a->dA, a->dA_dtype, 1,
                            integs, &integ_dtype, 1, nullptr,
                            a->lt ? XC_SCAN_PREFIX : XC_SCAN_TOTAL_MINUS, decr,
                            nullptr, so, nullptr, w_hist, pl.ws_hist, stream)) return 1;
    // (3) latEq = Table.lookup_coordinates(area)
    if (xc_interp(area, N, N, a->table, 0, a->table_coord, 0, a->n_table, -1, S, latEq, stream)) return 1;
    // (5)/(4) Lmin, d/dA, Leq2, nkeff
    if (xc_lmin(latEq, S * (long)N, Lmin, stream)) return 1;
    if (xc_gradient_wrt_area(intg, XC_F64, area, XC_F64, S, N, dintSdA, stream)) return 1;
    if (xc_gradient_wrt_area(ctr, a->ctr_dtype == XC_F32 ? XC_F32_AS_F64 : XC_F64, area, XC_F64, S, N, dqdA, stream)) return 1;
    if (xc_leq2(dintSdA, dqdA, S * (long)N, Leq2, stream)) return 1;
    if (xc_nkeff(Leq2, Lmin, a->keff_mask, S * (long)N, nkeff, stream)) return 1;
    // (3) Q(eq_coord) = interp_to_coords(eq_coord, latEq, ctr)
    if (xc_interp(a->eq_coord, 0, ny, latEq, N, ctr, N, N, -1, S, Qref, stream)) return 1;
    // (6) LWA
    if (a->lwa)
        if (xc_lwa(a->q, a->q_dtype, S, ny, nx, Qref, a->ww, a->increase, a->part, 1, a->lwa,
                   w_lwa, pl.ws_lwa, stream)) return 1;
    return 0;
}
