/*
 * xcb200.h -- C ABI of libxcb200.so: the B200 (sm_100a) contour-coordinate hot
 * path behind xcontour's Contour2D / Table Python API.
 *
 * The reference (miniufo/xcontour) is pure Python and has no FFI layer; its
 * boundary is the public method set of Contour2D / Table (xcontour/__init__.py:2-6).
 * Each entry point below replaces the numerical body of the reference method(s)
 * cited next to it (paths relative to the reference repository).  The Python
 * host mirror in xcontour_b200/ binds them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (obtained from a
 *     torch tensor's data_ptr() / DLPack capsule); the library allocates nothing
 *     persistent.  Scratch memory is passed in as (workspace, ws_bytes) after
 *     asking the matching *_workspace_bytes().
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - arrays are C-contiguous.  A tracer is q[S][P]: S independent slices of a
 *     plane with P = n0*n1 cells; contour-space arrays are [S][N].
 *   - dtype codes: XC_F32 / XC_F64.
 *   - return value: 0 on success, non-zero on failure; xc_last_error() returns a
 *     thread-local message for the last failure.  All calls are asynchronous with
 *     respect to the host (they only enqueue work on `stream`) and re-entrant.
 */
#ifndef XCB200_H
#define XCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XC_F32 0
#define XC_F64 1

/* scan modes of xc_bin_accumulate */
#define XC_SCAN_PREFIX        0 /* cdf[k] = sum_{p<=k} pdf[p]          (lt, hist path)      */
#define XC_SCAN_TOTAL_MINUS   1 /* cdf[k] = cdf[N-1] - cdf[k]          (core.py:1322-1323)  */
#define XC_SCAN_SUFFIX        2 /* cdf[k] = sum_{p>=k} pdf[p]          (strict '>' path)    */

/* parts of the LWA integral (core.py:773-784) */
#define XC_PART_ALL   0
#define XC_PART_UPPER 1
#define XC_PART_LOWER 2

#define XC_MAX_INTEGRANDS 3

/* ghost cells of the |grad q|^2 stencil (numpy.pad vocabulary in brackets); the reference's callers choose
 * them in xinvert.FiniteDiff(BCs=...) / GeoApps (tests/test_Keff_ocean.py:26-32, tests/test_clength.py:39-45) */
#define XC_BC_PERIODIC 0 /* wrap around                              ['wrap']     */
#define XC_BC_EXTEND   1 /* ghost = value of the edge cell           ['edge']     */
#define XC_BC_REFLECT  2 /* mirror about the edge point, q[-1]=q[1]  ['reflect']  */
#define XC_BC_FILL     3 /* ghost = fill_value                       ['constant'] */

const char* xc_last_error(void);
#define XC_ABI_VERSION 3
int         xc_abi_version(void);   /* == XC_ABI_VERSION of the header the library was built from */

/* ------------------------------------------------------------------------
 * (1) contour levels -- Contour2D.cal_contours(levels:int), core.py:222-249.
 * NaN-skipping min/max over each slice, then
 *   level_k = cast_dtype( (1.0/(N-1)) * f64(end - start) * k + f64(start) )
 * with `end - start` rounded in the tracer dtype (the promotion rule pinned by
 * notebooks/1.Keff_atmos.ipynb:102-119).  `levels` receives the values already
 * rounded to the output dtype (out_dtype) and widened to fp64.
 * minmax (nullable) receives [S][2] = (min, max).
 * ---------------------------------------------------------------------- */
size_t xc_minmax_levels_workspace_bytes(long S, long P);
int xc_minmax_levels(const void* q, int q_dtype, long S, long P,
                     int N, int increase, int out_dtype,
                     double* levels, double* minmax,
                     void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------
 * (1b) histogram bin edges -- _histogram, core.py:1273-1281 (per-'time' branch,
 * time_branch=1: edges promoted to fp64) and core.py:1296-1304 (static branch,
 * time_branch=0: edges stay in the contour dtype `ctr_dtype`).  Adds the extra
 * lowest bin, reverses decreasing contours, and applies xhistogram's "+1e-8 on
 * the last edge, in the edge dtype" rule (upstream, see oracle header).
 * edges: [S][N+1] ascending.  decreasing (nullable): [S] int32, 1 when the
 * contour array of that slice decreases (result must be flipped, core.py:454).
 * ---------------------------------------------------------------------- */
int xc_hist_edges(const double* levels, long S, int N, int ctr_dtype,
                  int time_branch, double* edges, int32_t* decreasing,
                  void* stream);

/* ------------------------------------------------------------------------
 * (2) conditional accumulation -- Contour2D.cal_integral_within_contours_hist
 * (core.py:412-460) + _histogram (core.py:1202-1325), and, with other flags,
 * the strict path cal_integral_within_contours (core.py:363-409) and the table
 * builders (core.py:73-203).
 *
 * Every cell is binned against ascending edges[N+1] (per slice when
 * edges_stride = N+1, shared when 0):  closed_right = 0 -> bin p holds
 * edges[p] <= q < edges[p+1]  (np.digitize right=False);  closed_right = 1 ->
 * edges[p] < q <= edges[p+1].  Cells outside and NaN cells are discarded.
 * Accumulated per bin, in fp64:  slot 0 = dA (when acc_area), then one slot per
 * integrand = integrand*dA with the product rounded in the operands' common
 * dtype (core.py:444) and NaN products replaced by 0 (core.py:449).
 * K = acc_area + n_int.  A block-wide scan over bins then forms the CDF
 * (scan_mode); when decreasing[s] != 0 (nullable) the bins of slice s are
 * written in reversed order (core.py:454-455).
 *
 * pdf (nullable) and cdf: [S][K][N] fp64.  bin_idx (nullable): [S][P] int32,
 * the bin of every cell or -1 -- for bit-exact parity checks.
 * q_mask (nullable): uint8 [P]; cells with q_mask == 0 are treated as NaN
 * (ctrVar.where(mask==1), core.py:178).
 * ---------------------------------------------------------------------- */
size_t xc_bin_accumulate_workspace_bytes(long S, long P, int N, int K);
int xc_bin_accumulate(const void* q, int q_dtype, long S, long P,
                      const double* edges, long edges_stride, int N,
                      int closed_right,
                      const void* dA, int dA_dtype, int acc_area,
                      const void* const* integrands, const int* integrand_dtypes,
                      int n_int,
                      const uint8_t* q_mask,
                      int scan_mode, const int32_t* decreasing,
                      double* pdf, double* cdf, int32_t* bin_idx,
                      void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------
 * (3) 1-D interpolation -- _interp1d (core.py:1405-1434), used by
 * Table.lookup_coordinates (core.py:1136-1174) and interp_to_coords
 * (core.py:1050-1100).  out[s][m] = np.interp(x[s][m], xp[s][:], fp[s][:]) with
 * numpy's arithmetic (slope*(x-xp[j])+fp[j], end clamping, exact-hit shortcut);
 * reverse != 0 evaluates np.interp(x, xp[::-1], fp[::-1]).  A stride of 0 shares
 * the vector between slices.
 * ---------------------------------------------------------------------- */
int xc_interp(const double* x, long x_stride, int M,
              const double* xp, long xp_stride,
              const double* fp, long fp_stride, int n, int reverse,
              long S, double* out, void* stream);

/* ------------------------------------------------------------------------
 * (4) d/dA -- Contour2D.cal_gradient_wrt_area, core.py:463-488:
 *   np.gradient(var)/np.gradient(area) along the contour axis (unit spacing,
 *   edge_order=1), each difference in its own dtype, the quotient in the
 *   promoted dtype.  out is fp64 [S][N] (holding fp32-rounded values when both
 *   inputs are fp32).
 * ---------------------------------------------------------------------- */
int xc_gradient_wrt_area(const void* var, int var_dtype,
                         const void* area, int area_dtype,
                         long S, int N, double* out, void* stream);

/* (4b) the same against an explicit contour coordinate (ABI 2): var.differentiate('contour') is np.gradient
 *   against the array's own 'contour' coordinate values (core.py:480-483), which are the level values when the
 *   caller passed explicit -- possibly non-uniform -- levels to cal_contours (core.py:253-264).  The O(N)
 *   difference coefficients are NumPy's own, computed on the host in the coordinate's dtype and handed over as
 *   fp64 [*_coef_f32: they hold fp32 values]: uniform spacing -> {2*dx, dx}; otherwise
 *   {a[N-2], b[N-2], c[N-2], dx_0, dx_n} (numpy.gradient, edge_order = 1).  var / area dtype: XC_F32 / XC_F64
 *   arrays; out fp64 [S][N] holding values rounded to the promoted dtype. */
int xc_gradient_wrt_area_coord(const void* var, int var_dtype, const double* var_coef, int var_uniform, int var_coef_f32,
                               const void* area, int area_dtype, const double* area_coef, int area_uniform, int area_coef_f32,
                               long S, int N, double* out, void* stream);

/* ------------------------------------------------------------------------
 * (5) Keff epilogue, all fp64 element-wise over n values:
 *   xc_leq2 : Leq2  = dgrdSdA / (dqdA*dqdA)                 core.py:635
 *   xc_lmin : Lmin  = 2*pi*Rearth*cos(deg2rad(lat))         utils.py:518-534
 *   xc_nkeff: nkeff = Leq2/Lmin/Lmin, NaN unless < mask     core.py:963-964
 *   xc_eqlat: asin(clip(A/2/pi/R/R - 1)) in degrees         utils.py:491-515
 * ---------------------------------------------------------------------- */
int xc_leq2(const double* dgrdSdA, const double* dqdA, long n, double* out, void* stream);
int xc_lmin(const double* lat_deg, long n, double* out, void* stream);
int xc_nkeff(const double* Leq2, const double* Lmin, double mask, long n, double* out, void* stream);
int xc_eqlat(const double* area, long n, double* out, void* stream);

/* ------------------------------------------------------------------------
 * (6) local wave activity / local APE -- Contour2D.cal_local_wave_activity
 * (core.py:696-799), cal_local_APE (core.py:908-942) and, with variant = 2,
 * cal_local_wave_activity2 (core.py:802-905).
 *
 * xc_lwa_weights: ww = (dA/max(dA)) * dA, the quotient rounded in dA's dtype
 *   (core.py:723-724), product in fp64.  ww: [n_eq][n_x] fp64.
 * xc_lwa: q[S][n_eq][n_x] with the equivalent dimension first in the plane,
 *   Qref[S][n_eq] (the sorted profile), out[S][n_eq][n_x] fp64:
 *     out[s][j][i] = - sum_j' (q[j'][i]-Q[j]) * mask(j,j',i) * ww[j'][i]
 *   For a profile that is monotone in the direction `increase` implies, each
 *   cell is added to one contiguous j-range (two binary searches in Q, a
 *   difference array and a per-column prefix sum); any other profile (or one
 *   with NaN) takes an exact O(n_eq^2) per-column kernel.  Both run on the GPU.
 * xc_lwa_mask: the integer mask of one reference row j (core.py:759-770),
 *   mask[S][n_eq][n_x] int8.
 * ---------------------------------------------------------------------- */
int xc_lwa_weights(const void* dA, int dA_dtype, long P, double* ww,
                   void* workspace, size_t ws_bytes, void* stream);
size_t xc_lwa_weights_workspace_bytes(long P);
size_t xc_lwa_workspace_bytes(long S);
int xc_lwa(const void* q, int q_dtype, long S, int n_eq, int n_x,
           const double* Qref, const double* ww,
           int increase, int part, int variant,
           double* out, void* workspace, size_t ws_bytes, void* stream);
/* xc_lwa_ex (ABI 2): xc_lwa with the optional hint ww_row[n_eq] fp64 = the value of ww along each row when the
 * weights are constant along x (every regular lat-lon / Cartesian / X-Z grid); selects the column-tile kernel
 * whose own-slot deposits stay in registers.  Same results; ww_row == NULL is xc_lwa. */
int xc_lwa_ex(const void* q, int q_dtype, long S, int n_eq, int n_x,
              const double* Qref, const double* ww, const double* ww_row,
              int increase, int part, int variant,
              double* out, void* workspace, size_t ws_bytes, void* stream);
int xc_lwa_mask(const void* q, int q_dtype, long S, int n_eq, int n_x,
                const double* Qref, int j, int increase, int variant,
                int8_t* mask, void* stream);

/* ------------------------------------------------------------------------
 * (7) |grad q|^2 on a regular lat-lon grid (periodic in x, one-sided at the
 * first/last row).  NOT part of the reference (its callers take this field from
 * xinvert / GeoApps, tests/test_Keff_ocean.py:31-32); provided so the Keff
 * integrand never has to be staged through the host.  lat_rad[n_y] fp64,
 * dlambda in radians.  out dtype = out_dtype (XC_F32 / XC_F64).
 * ---------------------------------------------------------------------- */
int xc_grad2_latlon(const void* q, int q_dtype, long S, int n_y, int n_x,
                    const double* lat_rad, double dlambda,
                    void* out, int out_dtype, void* stream);

/* ------------------------------------------------------------------------
 * (7b) cell areas of a regular lat-lon grid, built on the device:
 *   dA[j][i] = R^2 (sin(phi_{j+1/2}) - sin(phi_{j-1/2})) * dlambda
 * with cell edges midway between grid latitudes, clipped at the poles (either
 * direction of the latitude coordinate).  Replaces the rA metric the reference
 * obtains from xgcm in add_latlon_metrics (utils.py:43-259) for this grid type.
 * lat_deg[n_y] fp64 (degrees), dlambda_deg > 0; out[n_y][n_x] in out_dtype.
 * ---------------------------------------------------------------------- */
int xc_latlon_cell_area(const double* lat_deg, int n_y, int n_x, double dlambda_deg,
                        void* out, int out_dtype, void* stream);

/* ------------------------------------------------------------------------
 * (8) fused batch: Keff + LWA for B slices in one call (the path bench.py
 * times).  Chains (1) (1b) (2) (3) (4) (5) (3) (6) on device without host
 * round-trips: levels -> edges -> {area, int|grad q|^2 dA} CDFs -> latEq by
 * table lookup -> Lmin, d/dA, Leq2, nkeff -> Q(lat) -> LWA.
 * See xc_keff_lwa_batch_workspace_bytes for the scratch size.
 * contour-space outputs, each [S][N] fp64 (nullable individually):
 *   ctr, area, intgrdS, latEq, Lmin, dintSdA, dqdA, Leq2, nkeff
 * Qref: [S][n_y] fp64;  lwa: [S][n_y][n_x] fp64.
 * grdS: [S][P] (dtype grdS_dtype) or NULL -> computed in flight from q with the
 * stencil of (7) inside the binning kernel and never written to HBM.
 * The batch is walked in passes of `sub_batch` slices (32 at 721x1440), two passes in flight on internal
 * streams.  A slice is read from HBM once per stage (min/max, binning, LWA): a pass does not fit the 126 MB L2.
 * ---------------------------------------------------------------------- */
typedef struct xc_keff_lwa_args {
    const void*   q;          int q_dtype;
    long          S;          int n_y;  int n_x;
    int           N;          int increase;  int lt;
    int           ctr_dtype;  /* Contour2D(dtype=...) : XC_F32 default */
    const void*   dA;         int dA_dtype;
    const void*   grdS;       int grdS_dtype;       /* nullable */
    const double* lat_rad;    double dlambda;       /* lat-lon stencil metrics (grdS == NULL and cx == NULL) */
    const double* table;      const double* table_coord; int n_table; /* A(Yeq), ascending coord */
    const double* eq_coord;   /* [n_y] coordinate values Q is interpolated to */
    const double* ww;         /* [n_y][n_x] from xc_lwa_weights */
    double        keff_mask;  /* cal_normalized_Keff(mask=...) */
    int           part;
    int           sub_batch;  /* slices per internal pass; 0 = auto (~400 MB of q + LWA per pass) */
    double *ctr, *area, *intgrdS, *latEq, *Lmin, *dintSdA, *dqdA, *Leq2, *nkeff;
    double *Qref, *lwa;
    /* optional HOST pointer to XC_N_STAGES floats: per-stage device time in ms
     * (CUDA events around every stage, summed over the passes of this call).  When
     * non-NULL the call synchronises the stream before returning.  Stages:
     * 0 min/max+levels, 1 edges, 2 binning+scan, 3 contour-space epilogue, 4 LWA. */
    float* stage_ms;
    /* ---- ABI version 2 ---- */
    /* general stencil (grdS == NULL): row metrics [n_y] fp64 with dq/dx = (q[i+1]-q[i-1])*cx[j] and
     * dq/dy = (q[j+1]-q[j-1])*cy[j] (lat-lon, Cartesian, X-Z ...), ghost cells by bcx / bcy (XC_BC_*).
     * cx == NULL: lat-lon metrics from lat_rad / dlambda, periodic in x, edge value in y. */
    const double* cx;         const double* cy;
    int           bcx, bcy;   double fill_value;
    /* optional hints that select the fast kernels (results are the same without them):
     * dA_row [n_y] fp64: the cell area of every row when dA is constant along x (as on every regular
     *   lat-lon / Cartesian / X-Z grid); uniform_dA: all rows have the same area; any_degenerate: some row has
     *   cx^2 > 2^30 cy^2 (a pole row) -- reserves the accumulator of such rows.
     * ww_row [n_y] fp64: ww of every row when the LWA weights are constant along x. */
    const double* dA_row;     int uniform_dA;  int any_degenerate;
    const double* ww_row;
    /* opt-in, NOT a drop-in result: lwa points to [S][n_y][n_x] fp32 and receives the fp64 result rounded once
     * (halves the bytes of the largest output; needs ww_row and n_y <= 768) */
    int           lwa_f32;
    /* ---- ABI version 3 ---- */
    /* NumPy scalar-promotion regime of the reference's per-'time' bin edges (core.py:1273-1281): 0 = NumPy 1.x
     * (step and edges promoted to fp64, as xc_hist_edges with time_branch = 1), 1 = NEP 50 / NumPy >= 2 (the edge
     * array keeps the contour dtype, as time_branch = 0).  Only the first and last edge differ. */
    int           numpy2_rules;
} xc_keff_lwa_args;
#define XC_N_STAGES 5

size_t xc_keff_lwa_batch_workspace_bytes(long S, int n_y, int n_x, int N);
int xc_keff_lwa_batch(const xc_keff_lwa_args* args,
                      void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------
 * (9) equal-area contour levels by a weighted-quantile histogram -- north_star kernel (1); an extension of the
 * reference, which offers equally spaced levels (cal_contours, core.py:205-266) and levels at prescribed
 * equivalent coordinates (cal_contours_at_hist, core.py:316-360).  One call on the device: min/max ->
 * (N-1)*refine+1 equally spaced fine levels and their bin edges -> their area CDF by the histogram path of
 * core.py:412-460 (lt: area where q < level, else q > level) -> the A(q) relation inverted by np.interp at N
 * equally spaced areas -> rounded once to `out_dtype`.  levels: [S][N] fp64 holding the rounded values; adjacent
 * levels enclose, to within one fine bin, the same area.  numpy2_rules: as in xc_keff_lwa_args.
 * ---------------------------------------------------------------------- */
size_t xc_equal_area_levels_workspace_bytes(long S, long P, int N, int refine);
int xc_equal_area_levels(const void* q, int q_dtype, long S, long P,
                         const void* dA, int dA_dtype,
                         int N, int refine, int increase, int lt, int out_dtype, int numpy2_rules,
                         double* levels, void* workspace, size_t ws_bytes, void* stream);

/* number of kernel launches issued by this library on the calling thread since
 * the last xc_reset_launch_count() (bench.py reports it as gpu_launches). */
long xc_launch_count(void);
void xc_reset_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* XCB200_H */
