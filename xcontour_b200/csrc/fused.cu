// (8) Fused batch: Keff + LWA for a batch of slices, chained on one stream with
// no host round-trips.  Every stage is one of the kernels the stand-alone ABI
// entries launch (same arithmetic, same parity); the intermediate contour-space
// arrays live in the caller's workspace.  This is the call bench.py times.
//
// The batch is walked in passes of `sub` slices, two passes in flight on internal
// streams.  A slice is read three times (min/max, binning + in-flight |grad q|^2,
// LWA), each time from HBM: a 32-slice pass (133 MB of q + 266 MB of LWA at 721x1440)
// does not fit the 126 MB L2, and shorter passes measured slower (launch tails).  The
// pipeline therefore moves ~20 MB per slice against 12.46 MB of algorithmic traffic
// (profiles/traffic.json); each kernel is judged against its own compulsory bytes.
#include "common.cuh"
#include "internal.h"
#include <stdlib.h>
#include <vector>

using namespace xc;

namespace {
struct FusedPlan {
    long sub;
    size_t ws_minmax, ws_hist, ws_lwa;
    size_t total;
};

long auto_sub_batch(long S, long P)
{
    const char* env = getenv("XCB200_SUB_BATCH");
    long sub = env ? atol(env) : 0;
    if (sub <= 0) {
        // Measured on B200 (721x1440, two passes in flight): throughput rises up to
        // ~32 slices per pass (launch / tail overheads amortise; the kernels are
        // bound by shared-memory scatter, not by HBM, so spilling the 126 MB L2
        // costs less than short launches).
        sub = (long)(400.0e6 / (double)(P * 12));
        if (sub < 1) sub = 1;
    }
    return sub < S ? sub : S;
}

// internal streams/events of the two-passes-in-flight schedule, one set per host thread
#ifndef XC_LANES
#define XC_LANES 2       /* passes in flight */
#endif
struct Overlap { cudaStream_t s[XC_LANES]; cudaEvent_t fork, join[XC_LANES]; int dev; };
bool overlap_enabled()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("XCB200_OVERLAP"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
Overlap* overlap_streams()
{
    static thread_local Overlap ov; static thread_local bool ready = false;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    if (ready && ov.dev == dev) return &ov;
    if (ready) return nullptr;                    // one device per host thread (one process per GPU)
    for (int l = 0; l < XC_LANES; ++l) {
        if (cudaStreamCreateWithFlags(&ov.s[l], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&ov.join[l], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    if (cudaEventCreateWithFlags(&ov.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ov.dev = dev; ready = true;
    return &ov;
}

FusedPlan fused_plan(long S, int ny, int nx, int N, long sub_req)
{
    FusedPlan p;
    const long P = (long)ny * nx;
    p.sub = sub_req > 0 ? (sub_req < S ? sub_req : S) : auto_sub_batch(S, P);
    p.ws_minmax = xc_minmax_levels_workspace_bytes(p.sub, P);
    p.ws_hist = bin_accumulate_ws_bytes_stencil(p.sub, ny, nx, N);
    p.ws_lwa = xc_lwa_workspace_bytes(p.sub) + 512;
    size_t t = 0;
    t += align_up(p.ws_minmax, 256) + align_up(p.ws_hist, 256) + align_up(p.ws_lwa, 256);
    t += align_up((size_t)p.sub * (N + 1) * 8, 256);        // edges
    t += 2 * align_up((size_t)p.sub * 4, 256) + 1024;       // decreasing, sorted, flag
    t += 9 * align_up((size_t)p.sub * N * 8, 256);          // contour-space temporaries
    t += align_up((size_t)p.sub * ny * 8, 256);             // Qref
    t += align_up((size_t)p.sub * 2 * N * 8, 256);          // reduced partials (many CTAs per slice)
    t += align_up((size_t)p.sub * 16, 256) + align_up(lwa_scratch_doubles(p.sub, true) * 8, 256);   // (min, max), LWA scratch
    t *= XC_LANES;                                          // passes in flight (one per internal stream)
    t += 2 * align_up((size_t)ny * 8, 256);                 // row metrics
    t += align_up(lwa_wmax_doubles() * 8, 256);              // max |ww| partials
    p.total = t + 8192;
    return p;
}
}  // namespace

extern "C" size_t xc_keff_lwa_batch_workspace_bytes(long S, int n_y, int n_x, int N)
{
    if (S <= 0 || n_y <= 0 || n_x <= 0 || N <= 0) return 0;
    // sized for the largest pass any sub_batch setting can ask for
    size_t a = fused_plan(S, n_y, n_x, N, 0).total;
    size_t b = fused_plan(S, n_y, n_x, N, S).total;
    return a > b ? a : b;
}

extern "C" int xc_keff_lwa_batch(const xc_keff_lwa_args* a, void* workspace, size_t ws_bytes, void* stream)
{
    XC_REQUIRE(a, "xc_keff_lwa_batch: null args");
    XC_REQUIRE(a->q && a->dA && a->table && a->table_coord && a->eq_coord && a->ww,
               "xc_keff_lwa_batch: null input pointer");
    XC_REQUIRE(a->S > 0 && a->n_y >= 2 && a->n_x >= 2 && a->N >= 2 && a->n_table >= 1,
               "xc_keff_lwa_batch: bad sizes");
    XC_REQUIRE(a->grdS || a->lat_rad || (a->cx && a->cy), "xc_keff_lwa_batch: need grdS, lat_rad or (cx, cy) for the stencil");
    XC_REQUIRE(a->bcx >= 0 && a->bcx <= XC_BC_FILL && a->bcy >= 0 && a->bcy <= XC_BC_FILL, "xc_keff_lwa_batch: bad boundary condition");
    XC_REQUIRE(a->q_dtype == XC_F32 || a->q_dtype == XC_F64, "xc_keff_lwa_batch: bad q dtype");
    const long S = a->S; const int ny = a->n_y, nx = a->n_x, N = a->N;
    const long P = (long)ny * nx;
    const bool stencil = a->grdS == nullptr;
    FusedPlan pl = fused_plan(S, ny, nx, N, a->sub_batch);
    XC_REQUIRE(workspace && ws_bytes >= pl.total, "xc_keff_lwa_batch: workspace too small (%zu < %zu)",
               ws_bytes, pl.total);
    Arena ar(workspace, ws_bytes);
    struct Lane {
        char *w_minmax, *w_hist; int32_t *sorted, *any_unsorted, *decr; double *edges, *minmax, *lwa_scratch;
        double *t_ctr, *t_area, *t_intg, *t_latEq, *t_Lmin, *t_dint, *t_dq, *t_Leq2, *t_nk, *t_Q, *t_red;
    } lanes[XC_LANES];
    for (Lane& L : lanes) {
        L.w_minmax = ar.take<char>(pl.ws_minmax);
        L.w_hist = ar.take<char>(pl.ws_hist);
        L.sorted = ar.take<int32_t>((size_t)pl.sub);
        L.any_unsorted = ar.take<int32_t>(1);
        L.edges = ar.take<double>((size_t)pl.sub * (N + 1));
        L.decr = ar.take<int32_t>((size_t)pl.sub);
        L.minmax = ar.take<double>((size_t)pl.sub * 2);
        L.lwa_scratch = ar.take<double>(lwa_scratch_doubles(pl.sub, true));
        L.t_ctr = ar.take<double>((size_t)pl.sub * N);   L.t_area = ar.take<double>((size_t)pl.sub * N);
        L.t_intg = ar.take<double>((size_t)pl.sub * N);  L.t_latEq = ar.take<double>((size_t)pl.sub * N);
        L.t_Lmin = ar.take<double>((size_t)pl.sub * N);  L.t_dint = ar.take<double>((size_t)pl.sub * N);
        L.t_dq = ar.take<double>((size_t)pl.sub * N);    L.t_Leq2 = ar.take<double>((size_t)pl.sub * N);
        L.t_nk = ar.take<double>((size_t)pl.sub * N);    L.t_Q = ar.take<double>((size_t)pl.sub * ny);
        L.t_red = ar.take<double>((size_t)pl.sub * 2 * N);
    }
    double* rcos = ar.take<double>((size_t)ny);
    double* dphi = ar.take<double>((size_t)ny);
    double* wmaxp = ar.take<double>(lwa_wmax_doubles());
    XC_REQUIRE(ar.ok(), "xc_keff_lwa_batch: workspace accounting error");

    StencilArgs sa; sa.ny = ny; sa.nx = nx; sa.cx = rcos; sa.cy = dphi;
    if (stencil && a->cx && a->cy) {
        sa.cx = a->cx; sa.cy = a->cy; sa.bcx = a->bcx; sa.bcy = a->bcy; sa.fill = (float)a->fill_value;
        sa.any_degenerate = a->any_degenerate;
    } else if (stencil) { if (row_metrics(a->lat_rad, ny, a->dlambda, rcos, dphi, stream)) return 1; }
    sa.dA_row = a->dA_row; sa.uniform_dA = a->uniform_dA;
    if (a->lwa) { if (lwa_wmax(a->ww, P, wmaxp, stream)) return 1; }      // once per call, before the passes fork
    cudaStream_t st = (cudaStream_t)stream;
    const long npass = (S + pl.sub - 1) / pl.sub;
    // optional per-stage timing (forces the serial schedule so that stages do not overlap)
    struct Events {                                      // destroyed on every return path
        std::vector<cudaEvent_t> v;
        ~Events() { for (cudaEvent_t e : v) if (e) cudaEventDestroy(e); }
    } evs;
    std::vector<cudaEvent_t>& ev = evs.v;
    if (a->stage_ms) {
        ev.assign((size_t)npass * (XC_N_STAGES + 1), nullptr);
        for (auto& e : ev) XC_CUDA_OK(cudaEventCreate(&e));
    }
    // Two passes in flight: pass p runs start to finish on internal stream p % 2 with
    // its own temporaries, so the HBM-bound min/max pass and the latency-bound epilogue
    // of pass p+1 fill the gaps of the shared-memory-bound LWA kernel of pass p.
    Overlap* ov = nullptr;
    if (!a->stage_ms && npass >= 2 && overlap_enabled()) ov = overlap_streams();
    if (ov) {
        XC_CUDA_OK(cudaEventRecord(ov->fork, st));
        for (int l = 0; l < XC_LANES; ++l) XC_CUDA_OK(cudaStreamWaitEvent(ov->s[l], ov->fork, 0));
    }
    long pass = 0;
    auto mark = [&](int k) { if (a->stage_ms) cudaEventRecord(ev[(size_t)pass * (XC_N_STAGES + 1) + k], st); };
    const size_t qsz = a->q_dtype == XC_F32 ? 4 : 8;
    const size_t gsz = a->grdS_dtype == XC_F32 ? 4 : 8;

    for (long s0 = 0; s0 < S; s0 += pl.sub) {
        const long ns = S - s0 < pl.sub ? S - s0 : pl.sub;
        Lane& L = lanes[pass % XC_LANES];
        void* ps = ov ? (void*)ov->s[pass % XC_LANES] : stream;              // this pass's stream
        auto at = [&](double* user, double* tmp) { return user ? user + s0 * (long)N : tmp; };
        const void* q = (const char*)a->q + (size_t)s0 * P * qsz;
        double* ctr = at(a->ctr, L.t_ctr);       double* area = at(a->area, L.t_area);
        double* intg = at(a->intgrdS, L.t_intg); double* latEq = at(a->latEq, L.t_latEq);
        double* Lmin = at(a->Lmin, L.t_Lmin);    double* dint = at(a->dintSdA, L.t_dint);
        double* dq = at(a->dqdA, L.t_dq);        double* Leq2 = at(a->Leq2, L.t_Leq2);
        double* nk = at(a->nkeff, L.t_nk);
        double* Qref = a->Qref ? a->Qref + s0 * (long)ny : L.t_Q;

        mark(0);
        // (1)+(1b) min/max, levels and per-'time'-branch edges in two launches
        if (minmax_levels_impl(q, a->q_dtype, ns, P, N, a->increase, a->ctr_dtype, ctr, L.minmax,
                               L.edges, L.decr, L.any_unsorted, L.w_minmax, pl.ws_minmax, ps, a->numpy2_rules)) return 1;
        mark(1);
        mark(2);
        // (2) area and int |grad q|^2 dA in one pass over q (per-CTA partials only)
        ScanOut so; so.p[0] = so.p[1] = so.p[2] = so.p[3] = nullptr; so.stride = N;
        const void* integs[1] = { stencil ? nullptr : (const void*)((const char*)a->grdS + (size_t)s0 * P * gsz) };
        const int integ_dt[1] = { a->grdS_dtype };
        HistOnly ho;
        sa.minmax = L.minmax;
        if (bin_accumulate_impl(q, a->q_dtype, ns, P, L.edges, N + 1, N, 0, a->dA, a->dA_dtype, 1,
                                integs, integ_dt, stencil ? 0 : 1, nullptr,
                                a->lt ? XC_SCAN_PREFIX : XC_SCAN_TOTAL_MINUS, L.decr,
                                nullptr, so, nullptr, L.w_hist, pl.ws_hist, ps,
                                stencil ? &sa : nullptr, &ho)) return 1;
        mark(3);
        // scan + (3) latEq + (5)/(4) Lmin, d/dA, Leq2, nkeff + (3) Q(eq_coord), one launch
        if (scan_epilogue(ho.part, ho.C, ns, N, a->lt, L.decr, ctr, a->ctr_dtype == XC_F32,
                          a->table, a->table_coord, a->n_table, a->eq_coord, ny, a->keff_mask, a->increase,
                          area, intg, latEq, Lmin, dint, dq, Leq2, nk, Qref, L.sorted, L.any_unsorted, ps, L.t_red)) return 1;
        mark(4);
        // (6) LWA
        if (a->lwa)
            if (lwa_impl(q, a->q_dtype, ns, ny, nx, Qref, a->ww, a->increase, a->part, 1,
                         a->lwa_f32 ? reinterpret_cast<double*>(reinterpret_cast<float*>(a->lwa) + (size_t)s0 * P) : a->lwa + (size_t)s0 * P,
                         L.sorted, L.any_unsorted, true, L.minmax, L.lwa_scratch, ps, wmaxp, a->ww_row, a->lwa_f32)) return 1;
        mark(5);
        ++pass;
    }
    if (ov) {
        for (int l = 0; l < XC_LANES; ++l) {
            XC_CUDA_OK(cudaEventRecord(ov->join[l], ov->s[l]));
            XC_CUDA_OK(cudaStreamWaitEvent(st, ov->join[l], 0));
        }
    }
    if (a->stage_ms) {
        XC_CUDA_OK(cudaStreamSynchronize(st));
        for (int k = 0; k < XC_N_STAGES; ++k) a->stage_ms[k] = 0.f;
        for (long p = 0; p < npass; ++p)
            for (int k = 0; k < XC_N_STAGES; ++k) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev[(size_t)p * (XC_N_STAGES + 1) + k], ev[(size_t)p * (XC_N_STAGES + 1) + k + 1]);
                a->stage_ms[k] += ms;
            }
    }
    return 0;
}
