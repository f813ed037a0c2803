// k_lwa_cols: the fixed-point LWA kernel for weights that are constant along a row.  See lwa.cu for the
// reformulation of the reference's j-loop (xcontour/core.py:752-794) and lwa_fx.cu for the general-weights kernel.
#include "lwa_fx.cuh"
#include <stdlib.h>

namespace xc {

// ---------------------------------------------------------------------------
// Second-generation fixed-point kernel ("column tiles with register-resident own
// deposits") for weights that are constant along a row, ww[j][i] = ww_row[j] -- every
// regular lat-lon, Cartesian and X-Z grid.  Same arithmetic as k_lwa_fx (same scales,
// same rounding of every term; 2^-kS is folded into the (Q_j - c) table, which is exact --
// the two kernels agree bit for bit); what changes is where the
// work is done (ncu of k_lwa_fx: 201 warp-instructions per 32 cells, 64 % of the time on
// the prefix side):
//   * no weight loads: -w 2^kV and -rn(w 2^kS) of the ny rows sit in shared memory;
//   * a thread keeps the own-slot deposit of each of its <= 12 cells in registers
//     between the scatter and the prefix phase (it owns the same (column, rows) in both),
//     so the prefix walk neither re-derives it from global memory nor pays four more
//     atomics: scatter = one 64-bit deposit per accumulator at the FAR end of the range;
//   * lo and hi word of an accumulator are adjacent: one LDS.64 per accumulator and row
//     in the two prefix passes, immediate offsets in the unrolled walk;
//   * the search runs on thresholds in the tracer's own type (smallest fp32 > Q_j: for an fp32
//     value v, #{Q < v} = #{T <= v}), a 4096-bucket LUT leaves at most two probes for almost every cell;
//   * persistent CTAs walk consecutive tiles of the slice-major tile list, so the tables
//     of a slice are built once per CTA and slice, not once per tile.
// grid = SM count, block = 1024 = 16 columns x 64 row segments, one CTA per SM.
constexpr int LC_TC = 16;                  // columns per tile; threads = row segments (64, or 32 with twice the rows each) x LC_TC
struct LwaColsSmem { size_t farS, farV, nxs, wrow, qc, ta, uni, lut, tot, total; };
// `ny` here is the compile-time row CAPACITY of the instantiation (256 / 512 / 736), not the run-time row count:
// every offset is then a constant and a table access is one instruction with an immediate (ncu of the run-time
// layout: ~8 % of the kernel's instructions re-derived region bases, with 64 registers per thread there is no room
// to keep them).
static __host__ __device__ constexpr LwaColsSmem lwa_cols_layout(int ny, int tbytes)
{
    LwaColsSmem L{}; size_t o = 0;
    const size_t plane = (size_t)(ny + 1) * LC_TC * 8;
    const size_t col8 = (size_t)((ny + 1) & ~1) * 8;
    L.farS = o; o += plane;
    L.farV = o; o += plane;
    // per-slice tables, built when a CTA meets a new slice
    L.nxs = o;  o += col8;                                     // -rn(w 2^kS)
    L.wrow = o; o += col8;                                     // -w 2^kV
    L.qc = o;   o += col8;                                     // (Q_j - c) 2^-kS
    L.ta = o;   o += (size_t)((ny + 2 + 3) & ~3) * tbytes;     // smallest value > Q_j, two +inf entries past the end
    // per tile: the LUT (scatter phase) and the segment totals (prefix phase) share one region
    L.uni = o; L.lut = o; L.tot = o;
    const size_t a = (size_t)((FX_LUT + 2 + 7) & ~7) * 2, b = (size_t)2 * LC_TC * FX_TOTP * 8;      // (totals: sized for 64 segments)
    L.total = o + (a > b ? a : b);
    return L;
}
// 64-bit two's-complement add into the adjacent (lo, hi) words at p with two native 32-bit shared atomics; the
// carry out of the low word is decided by the value it held when THIS add reached it, so the pair ends up as the
// exact sum modulo 2^64 in any interleaving.  (atomicAdd intrinsics, not inline PTX: the compiler may then overlap
// the independent chains of neighbouring cells.)
__device__ __forceinline__ void lc_add64(uint32_t* p, long long x)
{
    const uint32_t xl = (uint32_t)x, xh = (uint32_t)((unsigned long long)x >> 32);
    const uint32_t old = atomicAdd(p, xl);
    atomicAdd(p + 1, xh + (uint32_t)(((unsigned long long)old + xl) >> 32));
}
// smallest value of the threshold type that is > x:  for a value v of that type, (x < v) == (thr(x) <= v)
__device__ __forceinline__ float lc_thr(double x, float)
{
    float f = __double2float_rn(x);
    if (!((double)f > x)) {                                        // next float above f
        const int b = __float_as_int(f);
        f = (f == 0.0f) ? __int_as_float(1) : __int_as_float(f > 0.0f ? b + 1 : b - 1);
    }
    return f;
}
__device__ __forceinline__ double lc_thr(double x, double) { return nextafter(x, CUDART_INF); }

// The scatter phase is written without data-dependent branches: a warp executes every path of a divergent
// branch, and in the first version of this kernel (divergent search loops, region-1 / region-2 / inactive
// paths, `continue`s) that made the executed warp-instruction count 2.6x the per-thread count (ncu: 199
// warp-instructions per 32 cells, same as k_lwa_fx).  Here every cell does the same thing, six cells at a time so
// that their dependent chains (LUT -> threshold probes -> slot -> atomics) overlap:
//   A  x = #{Q < v} from the LUT and two independent probes; buckets with more than two thresholds (the flat ends
//      of the profile) are flagged, ONE warp-wide vote per six cells sends them to a bisection;
//   B  the target slot by selects and ALWAYS one far deposit per accumulator -- a cell without a range (or a NaN
//      cell) deposits at its own slot jp + 1, where the walk's unconditional own deposit cancels it exactly.
// Exact ties v == Q_j need no care: such a row contributes w (v - Q_j) = 0 whichever side it is counted on.
template <typename QT, bool INC, int NYCAP, int SEG>
__global__ void __launch_bounds__(SEG * LC_TC, 1)
k_lwa_cols(const QT* __restrict__ q, long s0, int nslices, int ny, int nx,
           const double* __restrict__ Qref, const double* __restrict__ ww_row,
           int part, const int32_t* __restrict__ sorted,
           const FxScale* __restrict__ fxs, const uint32_t* __restrict__ lutg,
           void* __restrict__ out_v, int out_f32)
{
    using TT = QT;                                               // thresholds live in the tracer's own type
    constexpr int LC_B = 4;                                      // cells per batch (LC_U is a multiple)
    constexpr int LC_NT = SEG * LC_TC;
    constexpr int LC_U = ((NYCAP + SEG - 1) / SEG + LC_B - 1) / LC_B * LC_B;      // rows a thread owns, at most
    static_assert(FX_LUT % LC_NT == 0, "the LUT is copied in whole rounds of the CTA");
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr LwaColsSmem L = lwa_cols_layout(NYCAP, (int)sizeof(TT));
    long long* nxs = reinterpret_cast<long long*>(smem + L.nxs);
    uint16_t*  lut = reinterpret_cast<uint16_t*>(smem + L.lut);
    double*    wrow = reinterpret_cast<double*>(smem + L.wrow);
    TT*        ta = reinterpret_cast<TT*>(smem + L.ta);
    long long* tot = reinterpret_cast<long long*>(smem + L.tot);
    double*    qcs = reinterpret_cast<double*>(smem + L.qc);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr double sg = INC ? 1.0 : -1.0;
    const bool keep_pos = (part == XC_PART_UPPER) == INC;
    const bool use_t1 = (part == XC_PART_ALL) || !keep_pos;   // mask -1 region
    const bool use_t2 = (part == XC_PART_ALL) || keep_pos;    // mask +1 region
    const int c = tid & (LC_TC - 1), seg = tid / LC_TC;
    const int r0 = (int)(((long)seg * ny) / SEG), r1 = (int)(((long)(seg + 1) * ny) / SEG);
    const int tps = (nx + LC_TC - 1) / LC_TC;                  // tiles per slice
    const long ntiles = (long)nslices * tps;
    const long t_beg = ntiles * blockIdx.x / gridDim.x, t_end = ntiles * (blockIdx.x + 1) / gridDim.x;
    // this thread's column of the two planes: entry (slot t) = 8 bytes at [t * LC_TC + c]
    uint32_t* farS_c = reinterpret_cast<uint32_t*>(smem + L.farS) + 2 * c;
    uint32_t* farV_c = reinterpret_cast<uint32_t*>(smem + L.farV) + 2 * c;
    const long long* colS = reinterpret_cast<const long long*>(smem + L.farS) + (size_t)r0 * LC_TC + c;
    const long long* colV = reinterpret_cast<const long long*>(smem + L.farV) + (size_t)r0 * LC_TC + c;

    long cur_slice = -1;
    double fc = 0.0, fiV = 0.0; float scalef = 0.f, qminf = 0.f;
    bool slice_ok = false;
    for (long tile = t_beg; tile < t_end; ++tile) {
        const long sl = tile / tps; const int tx = (int)(tile - sl * tps);
        const long s = s0 + sl;
        const bool fresh = sl != cur_slice;
        if (fresh) { cur_slice = sl; slice_ok = sorted[s] != 0; }
        if (!slice_ok) continue;                                  // uniform: the exact loop takes this slice
        const FxScale* fp = fxs + sl;
        const double* Qg = Qref + s * (long)ny;
        // ---- phase 0: zero the planes, tables of a new slice, LUT ----
        // (the LUT words come from L2 and are only needed at the end of the phase: requested first, so that their
        // latency is spent behind the barrier and the zeroing instead of in front of the copy loop)
        uint32_t pk[FX_LUT / LC_NT];
#pragma unroll
        for (int t = 0; t < FX_LUT / LC_NT; ++t) pk[t] = __ldg(lutg + (size_t)sl * FX_LUT + tid + t * LC_NT);
        __syncthreads();                                          // previous tile's walk is done with the planes / totals
        {
            uint4* zS = reinterpret_cast<uint4*>(smem + L.farS);          // the planes are NYCAP + 1 slots apart: zero the
            uint4* zV = reinterpret_cast<uint4*>(smem + L.farV);          // ny + 1 slots in use of each
            const int n16 = (ny + 1) * (LC_TC / 2);
            for (int k = tid; k < n16; k += LC_NT) { zS[k] = make_uint4(0u, 0u, 0u, 0u); zV[k] = make_uint4(0u, 0u, 0u, 0u); }
        }
        if (fresh) {                                              // tables of this slice (shared by all its tiles)
            fc = __ldg(&fp->c);
            const double fsS = __ldg(&fp->sS), fsV = __ldg(&fp->sV), fiS = __ldg(&fp->iS);
            fiV = __ldg(&fp->iV);
            for (int j = tid; j < ny + 2; j += LC_NT) {
                if (j >= ny) { ta[j] = (TT)CUDART_INF; continue; }              // probes may look one past the end
                const double Qj = sg * Qg[j], w = __ldg(ww_row + j);
                ta[j] = lc_thr(Qj, TT());
                wrow[j] = (w == w) ? __dmul_rn(w, -fsV) : 0.0;              // a NaN weight deposits nothing
                nxs[j] = (w == w) ? fx_rn(__dmul_rn(w, -fsS)) : 0ll;
                qcs[j] = __dmul_rn(__dsub_rn(Qj, fc), fiS);
            }
            const double qmin = sg * Qg[0], qmax = sg * Qg[ny - 1];
            qminf = (float)qmin;
            scalef = fx_scale(qmin, qmax);
        }
        // LUT: first row of every bucket (+ the end marker), unpacked from the (first[b], first[b+1]) pairs
#pragma unroll
        for (int t = 0; t < FX_LUT / LC_NT; ++t) {
            const int k = tid + t * LC_NT;
            lut[k] = (uint16_t)(pk[t] & 0xffffu);
            if (k == FX_LUT - 1) lut[FX_LUT] = (uint16_t)(pk[t] >> 16);
        }
        __syncthreads();

        // ---- phase 1: scatter -- one deposit of -X per accumulator at the far end of each cell's range ----
        const int i = tx * LC_TC + c;
        const bool col_ok = i < nx;
        long long NVr[LC_U];                                      // -X_V of the thread's cells
        long long ownS = 0;
        {
            const QT* qp = q + (s * (long)ny + r0) * nx + (col_ok ? i : 0);
#pragma unroll
            for (int u0 = 0; u0 < LC_U; u0 += LC_B) {
                TT vts[LC_B]; int xs[LC_B]; unsigned oddm = 0u;
#pragma unroll
                for (int k = 0; k < LC_B; ++k) {                              // loads of the batch in flight together
                    const QT qraw = (col_ok && r0 + u0 + k < r1) ? __ldg(qp) : (QT)CUDART_NAN;
                    vts[k] = INC ? (TT)qraw : -(TT)qraw;
                    qp += nx;
                }
                // A: every lane runs the whole body (the vote is warp-wide); lanes without a cell -- past the last
                // column, or the 12th row of an 11-row segment -- carry NaN and skip only the deposits
#pragma unroll
                for (int k = 0; k < LC_B; ++k) {
                    const int u = u0 + k, jp = r0 + u;
                    const bool live = col_ok && jp < r1;
                    const TT vt = vts[k];
                    NVr[u] = live ? fx_rn(__dmul_rn(__dsub_rn((double)vt, fc), wrow[live ? jp : 0])) : 0ll;     // NaN -> 0
                    const int bk = fx_bucket((float)vt, qminf, scalef);
                    const int x0 = (int)lut[bk], cnt = (int)lut[bk + 1] - x0;
                    const TT t0 = ta[x0], t1 = ta[x0 + 1];
                    xs[k] = x0 + ((cnt > 0 && t0 <= vt) ? 1 : 0) + ((cnt > 1 && t1 <= vt) ? 1 : 0);            // #{Q < v}
                    if (live && cnt > 2) oddm |= 1u << k;
                }
                if (__any_sync(XC_FULL, oddm != 0u)) {                        // rare: the flat ends of the profile
#pragma unroll
                    for (int k = 0; k < LC_B; ++k)
                        if (oddm & (1u << k)) {
                            const TT vt = vts[k];
                            const int bk = fx_bucket((float)vt, qminf, scalef);
                            int x = (int)lut[bk], e = (int)lut[bk + 1];
                            while (x < e) { const int mid = (x + e) >> 1; if (ta[mid] <= vt) x = mid + 1; else e = mid; }
                            xs[k] = x;
                        }
                }
                // B: slots and deposits
#pragma unroll
                for (int k = 0; k < LC_B; ++k) {
                    const int u = u0 + k, jp = r0 + u;
                    if (!(col_ok && jp < r1)) continue;                       // (thread-level: tile edge / short segment)
                    const int x = xs[k];
                    const bool act = (vts[k] == vts[k]) && ((x > jp + 1 && use_t1) || (x <= jp && use_t2));
                    const int target = act ? x : jp + 1;                      // no range, or a NaN cell: cancels at its own slot
                    const long long NS = nxs[jp];
                    lc_add64(farS_c + (size_t)target * (2 * LC_TC), NS);
                    lc_add64(farV_c + (size_t)target * (2 * LC_TC), NVr[u]);
                    ownS -= NS;
                }
            }
        }
        __syncthreads();

        // ---- phase 2: segment totals, block scan over the 64 segments of every (accumulator, column) ----
        {
            long long aS = ownS, aV = 0;
#pragma unroll
            for (int u = 0; u < LC_U; ++u) {
                aV -= NVr[u];                                             // own deposits: +X at slot jp + 1
                if (r0 + u < r1) { aS += colS[u * LC_TC]; aV += colV[u * LC_TC]; }
            }
            // the shared region changes hands here: LUT -> totals; every thread is past the scatter phase
            tot[c * FX_TOTP + seg] = aS;
            tot[(LC_TC + c) * FX_TOTP + seg] = aV;
        }
        __syncthreads();
        {                                                         // a warp per (accumulator, column) row: exclusive scan over segments
            constexpr int RPW = (2 * LC_TC) / (LC_NT / 32), EPL = SEG / 32;
#pragma unroll
            for (int rr = 0; rr < RPW; ++rr) {
                long long* row = tot + (size_t)(warp * RPW + rr) * FX_TOTP;
                const long long a0 = row[EPL * lane], a1 = EPL == 2 ? row[2 * lane + 1] : 0ll;
                long long x = a0 + a1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(XC_FULL, x, o); if (lane >= o) x += t; }
                const long long ex = x - (a0 + a1);
                row[EPL * lane] = ex;
                if (EPL == 2) row[2 * lane + 1] = ex + a0;
            }
        }
        __syncthreads();

        // ---- phase 3: the walk ----
        if (col_ok) {
            long long RS = tot[c * FX_TOTP + seg], RV = tot[(LC_TC + c) * FX_TOTP + seg];
            const long o0 = (s * (long)ny + r0) * nx + i;
            char* op = reinterpret_cast<char*>(out_v) + o0 * (out_f32 ? 4 : 8);
            const long ostep = (long)nx * (out_f32 ? 4 : 8);
#pragma unroll
            for (int u = 0; u < LC_U; ++u) {
                if (r0 + u >= r1) break;
                RS += colS[u * LC_TC];
                RV += colV[u * LC_TC];
                const double Vj = __dmul_rn(fx_to_double(RV), fiV);
                const double val = sg * (Vj - qcs[r0 + u] * fx_to_double(RS));
                if (out_f32) *reinterpret_cast<float*>(op) = (float)val;      // opt-in: the fp64 result rounded once
                else *reinterpret_cast<double*>(op) = val;
                op += ostep;
                RS -= nxs[r0 + u];                                        // own deposits of cell r0 + u, at slot r0 + u + 1
                RV -= NVr[u];
            }
        }
    }
}

}  // namespace xc

using namespace xc;

static int lwa_cols_cap(int n_eq) { return n_eq <= 256 ? 256 : n_eq <= 512 ? 512 : n_eq <= 736 ? 736 : 0; }

bool xc::lwa_cols_fits(int n_eq, int qbytes)
{
    const int cap = lwa_cols_cap(n_eq);
    return n_eq >= 2 && cap > 0 && lwa_cols_layout(cap, qbytes).total <= 227 * 1024;
}

int xc::lwa_cols_launch(const void* q, int q_dtype, long s0, long ns, int n_eq, int n_x, const double* Qref, const double* ww_row,
                        int increase, int part, const int32_t* sorted, const FxScale* fxs, const uint32_t* lutg,
                        void* out, int out_f32, void* stream)
{
    const int cap = lwa_cols_cap(n_eq);
    const size_t smem = lwa_cols_layout(cap, q_dtype == XC_F32 ? 4 : 8).total;
    const long tiles = ns * ((n_x + LC_TC - 1) / LC_TC);
    const unsigned grid = (unsigned)(tiles < sm_count() ? tiles : sm_count());
    auto go = [&](auto kern, auto qptr, int nt) -> int {
        XC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, nt, smem, (cudaStream_t)stream>>>(qptr, s0, (int)ns, n_eq, n_x, Qref, ww_row, part, sorted, fxs, lutg, out, out_f32);
        XC_LAUNCH_OK();
        return 0;
    };
    // (32 row segments per column -- 512 threads with 128 registers and twice the rows each -- measured 0.277 ms per
    // 32 slices against 0.253 for 64 segments: the template parameter stays, only 64 is instantiated)
#define XC_LC_GO(QT, CAP) (increase ? go(k_lwa_cols<QT, true, CAP, 64>, (const QT*)q, 64 * LC_TC) : go(k_lwa_cols<QT, false, CAP, 64>, (const QT*)q, 64 * LC_TC))
    if (q_dtype == XC_F32) return cap == 256 ? XC_LC_GO(float, 256) : cap == 512 ? XC_LC_GO(float, 512) : XC_LC_GO(float, 736);
    return cap == 256 ? XC_LC_GO(double, 256) : cap == 512 ? XC_LC_GO(double, 512) : XC_LC_GO(double, 736);
#undef XC_LC_GO
}
