// k_lwa_fx + its per-slice preparation kernel.  See lwa.cu for the reformulation of the reference's j-loop
// (xcontour/core.py:752-794) as difference arrays + one prefix sum per column.
#include "lwa_fx.cuh"
#include <stdlib.h>

namespace xc {

// ---------------------------------------------------------------------------
// Fixed-point LWA kernel for GENERAL weights ww[n_eq][n_x] (the fallback of lwa_cols.cu, which covers weights
// that are constant along a row).  Measured on B200
// (scripts/micro/atoms_bench.cu): a native shared-memory integer atomic
// (ATOMS.ADD.U32) costs 2.6-3.4 cycles per warp instruction, a 64-bit add built
// from two of them plus a carry 4.3-5.0, against 15-18.5 for ONE conflict-free
// read-modify-write round of an fp64 pair (and 2.4 election rounds on average in
// k_lwa_fast).  So the difference arrays become 64-bit two's-complement integers:
//     X_S = rn(w * 2^kS),   X_V = rn(w * (v - c) * 2^kV)
// with c the mid-range of the slice and kS, kV chosen per slice so that a column
// of n_eq terms cannot overflow 63 bits (|X| < 2^min(51, 62 - ceil(log2(n_eq+1))): 51
// significant bits at n_eq = 721, i.e. the resolution of the largest term's own
// fp64 ulp).  Integer adds are exact and order-independent, so
//   * any thread may deposit into any slot of its column tile: lanes run along
//     x (coalesced loads, no transposed staging), there is no lane election, no
//     warp-private array, and a CTA has as many warps as registers allow;
//   * the +X deposit at slot j'+1 never touches shared memory: the prefix pass
//     re-derives it from (q, ww) with the same rounding, and an inactive cell
//     deposits -X at j'+1, which cancels exactly;
//   * results are bit-reproducible whatever the schedule.
//   LWA[j] = sg * ( V_j 2^-kV - (Q_j - c) * S_j 2^-kS ).
// grid = (ceil(n_x / 8), slices); block = FX_NT threads = (column c, row segment).
// Shared memory: four u32 planes [slot][column] (lo/hi words of S and V), Q, a
// 4096-bucket LUT over Q and the per-segment totals of the block scan.
// Two tile shapes: 16 columns x 1024 threads (one CTA per SM; a warp-wide global
// access then covers 2 rows x 64/128 B instead of 4 rows x 32/64 B, which is what
// the LSU data pipe is paid in) while the four planes fit 227 KB, else 8 x 512.

struct LwaFxSmem { size_t off_Q, off_far, off_lut, off_tot, total; int plane; };
static __host__ __device__ inline LwaFxSmem lwa_fx_layout(int ny, int FX_TC)
{
    LwaFxSmem L;
    L.plane = ((ny + 1) * FX_TC + 3) & ~3;                       // u32 words per plane
    size_t o = 0;
    L.off_Q = o;   o += (size_t)((ny + 1) & ~1) * 8;
    L.off_far = o; o += (size_t)4 * L.plane * 4;
    L.off_lut = o; o += (size_t)FX_LUT * 4;
    L.off_tot = o; o += (size_t)2 * FX_TC * FX_TOTP * 8;
    L.total = o;
    return L;
}


// 64-bit two's-complement add into (lo[idx], hi[idx]) with two native 32-bit
// shared atomics; the carry out of the low word is decided by the value the low
// word held when THIS add reached it, so the pair ends up as the exact sum modulo
// 2^64 in any interleaving.
__device__ __forceinline__ void fx_add64(uint32_t* lo, uint32_t* hi, int idx, long long x)
{
    const uint32_t xl = (uint32_t)x, xh = (uint32_t)((unsigned long long)x >> 32);
    const uint32_t old = atomicAdd(lo + idx, xl);
    const uint32_t carry = (uint32_t)((old + xl) < xl);
    atomicAdd(hi + idx, xh + carry);
}

// Per-slice preparation (one CTA per slice): the fixed-point scales from the
// NaN-skipping (min, max) of the slice and max |ww|, and the LUT over Q
// (first row whose bucket is >= b, packed (first[b], first[b+1])), both shared by
// every column tile of the slice.  A slice with an infinite value is handed to the
// exact loop (sorted[s] = 0, *any_unsorted = 1).

__global__ void __launch_bounds__(FX_PREP_NT)
k_lwa_fx_prep(long s0, int ny, const double* __restrict__ Qref, int increase,
              int32_t* sorted, int32_t* any_unsorted,
              const double* __restrict__ rng, int rngC, const double* __restrict__ wmax_part, int n_wmax,
              FxScale* __restrict__ fxs, uint32_t* __restrict__ lutg)
{
    const long s = s0 + blockIdx.x;
    if (!sorted[s]) return;
    __shared__ uint16_t first[FX_LUT + 2];
    __shared__ double swm[FX_PREP_NT / 32];
    const int tid = threadIdx.x;
    const double sg = increase ? 1.0 : -1.0;
    const double* Qg = Qref + s * (long)ny;
    double wm = 0.0;
    for (int k = tid; k < n_wmax; k += FX_PREP_NT) wm = fmax(wm, wmax_part[k]);
    wm = warp_max(wm);
    if ((tid & 31) == 0) swm[tid >> 5] = wm;
    const double qmin = sg * Qg[0], qmax = sg * Qg[ny - 1];
    const float qminf = (float)qmin;
    const float scalef = fx_scale(qmin, qmax);
    for (int j = tid; j <= ny; j += FX_PREP_NT) {
        const int bj = (j < ny) ? fx_bucket((float)(sg * Qg[j]), qminf, scalef) : FX_LUT;
        const int bp = (j > 0) ? fx_bucket((float)(sg * Qg[j - 1]), qminf, scalef) : -1;
        for (int b = bp + 1; b <= bj; ++b) first[b] = (uint16_t)j;
    }
    __syncthreads();
    uint32_t* lut = lutg + (size_t)blockIdx.x * FX_LUT;
    for (int b = tid; b < FX_LUT; b += FX_PREP_NT) lut[b] = (uint32_t)first[b] | ((uint32_t)first[b + 1] << 16);
    if (tid == 0) {
        for (int k = 1; k < FX_PREP_NT / 32; ++k) wm = fmax(wm, swm[k]);
        double lo = CUDART_INF, hi = -CUDART_INF;
        for (int k = 0; k < rngC; ++k) {
            lo = fmin(lo, rng[(s * rngC + k) * 2]); hi = fmax(hi, rng[(s * rngC + k) * 2 + 1]);
        }
        const double a = sg * lo, b = sg * hi;
        const double vlo = fmin(a, b), vhi = fmax(a, b);
        FxScale f;
        f.c = 0.5 * vlo + 0.5 * vhi;
        const double vabs = fmax(vhi - f.c, f.c - vlo);
        const double MV = wm * vabs * 1.0000001, MS = wm;
        const bool empty = !(lo <= hi);                              // slice without a finite value
        if (!empty && !(isfinite(MV) && isfinite(MS) && isfinite(f.c))) {  // inf in q or ww: exact loop instead
            sorted[s] = 0; if (any_unsorted) *any_unsorted = 1;
        }
        int hb = 1; while ((1 << hb) < ny + 1) ++hb;                 // sums of up to ny terms
        const int kb = min(51, 62 - hb) - 1;                         // |X| < 2^(kb+1): fx_rn range and 63-bit column sums
        const int kS = (MS > 0.0 && isfinite(MS) && !empty) ? kb - ilogb(MS) : 0;
        const int kV = (MV > 0.0 && isfinite(MV) && !empty) ? kb - ilogb(MV) : 0;
        if (empty || !isfinite(f.c)) f.c = 0.0;
        f.sS = scalbn(1.0, kS); f.iS = scalbn(1.0, -kS);
        f.sV = scalbn(1.0, kV); f.iV = scalbn(1.0, -kV);
        f.pad0 = f.pad1 = f.pad2 = 0.0;
        fxs[blockIdx.x] = f;
    }
}

// the deposit of one cell; used by the scatter phase (with negated scales: rn() is
// odd, so rn(-x) = -rn(x) bit for bit) and re-derived by the prefix phase
__device__ __forceinline__ void fx_terms(double v, double w, double c, double sS, double sV, long long& XS, long long& XV)
{
    XS = fx_rn(__dmul_rn(w, sS));
    XV = fx_rn(__dmul_rn(__dmul_rn(w, __dsub_rn(v, c)), sV));
}

constexpr int FX_U = 4;        // rows whose loads are in flight together

template <typename QT, int FX_TC>
__global__ void __launch_bounds__(FX_SEG * FX_TC, FX_TC <= 8 ? 2 : 1)
k_lwa_fx(const QT* __restrict__ q, long s0, long sbase, int ny, int nx,
         const double* __restrict__ Qref, const double* __restrict__ ww,
         int increase, int part, const int32_t* __restrict__ sorted,
         const FxScale* __restrict__ fxs, const uint32_t* __restrict__ lutg,
         double* __restrict__ out)
{
    const long s = s0 + blockIdx.y;
    if (!sorted[s]) return;
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int FX_NT = FX_SEG * FX_TC;
    static_assert(FX_NT / 32 >= 2 * FX_TC && FX_LUT % FX_NT == 0, "one scan warp per (accumulator, column)");
    const LwaFxSmem L = lwa_fx_layout(ny, FX_TC);
    double*    Qs  = reinterpret_cast<double*>(smem + L.off_Q);
    uint32_t*  far = reinterpret_cast<uint32_t*>(smem + L.off_far);
    uint32_t*  lut = reinterpret_cast<uint32_t*>(smem + L.off_lut);
    long long* tot = reinterpret_cast<long long*>(smem + L.off_tot);     // [2][FX_TC][FX_TOTP]
    const int plane = L.plane;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double sg = increase ? 1.0 : -1.0;
    const float sgf = increase ? 1.0f : -1.0f;
    const double* Qg = Qref + s * (long)ny;
    const uint32_t* lg = lutg + (size_t)(s - sbase) * FX_LUT;
    for (int j = tid; j < ny; j += FX_NT) Qs[j] = sg * Qg[j];
#pragma unroll
    for (int k = 0; k < FX_LUT / FX_NT; ++k) lut[tid + k * FX_NT] = __ldg(lg + tid + k * FX_NT);
    {
        uint4* z = reinterpret_cast<uint4*>(far);
        for (int k = tid; k < plane; k += FX_NT) z[k] = make_uint4(0u, 0u, 0u, 0u);   // 4 planes of `plane` words
    }
    const FxScale* fp = fxs + (s - sbase);
    const double fc = __ldg(&fp->c), fsS = __ldg(&fp->sS), fsV = __ldg(&fp->sV);
    __syncthreads();
    const double qmin = Qs[0], qmax = Qs[ny - 1];
    const float qminf = (float)qmin;
    const float scalef = fx_scale(qmin, qmax);

    const bool keep_pos = (part == XC_PART_UPPER) == (increase != 0);
    const bool use_t1 = (part == XC_PART_ALL) || !keep_pos;   // mask -1 region
    const bool use_t2 = (part == XC_PART_ALL) || keep_pos;    // mask +1 region

    // thread = (column c, row segment seg); the segments split the rows evenly
    const int c = tid & (FX_TC - 1), seg = tid / FX_TC;
    const int i = blockIdx.x * FX_TC + c;
    const bool col_ok = i < nx;
    const int r0 = (int)(((long)seg * ny) / FX_SEG), r1 = (int)(((long)(seg + 1) * ny) / FX_SEG);
    const QT* qc = q + (s * (long)ny + r0) * nx + i;
    const double* wc = ww + (long)r0 * nx + i;
    uint32_t* fcol = far + c;                                  // word (slot t, plane k) = fcol[t * FX_TC + k * plane]

    // ---- scatter: one deposit of -X at the far end of each cell's range ----
    long long ownS = 0, ownV = 0;
    if (col_ok) {
        const double nsS = -fsS, nsV = -fsV;
        const QT* qp = qc; const double* wp = wc;
        for (int jb = r0; jb < r1; jb += FX_U) {
            QT qv[FX_U]; double wv[FX_U];
#pragma unroll
            for (int u = 0; u < FX_U; ++u) {
                const bool ok = jb + u < r1;
                qv[u] = ok ? __ldg(qp + (long)u * nx) : (QT)CUDART_NAN;
                wv[u] = ok ? __ldg(wp + (long)u * nx) : 0.0;
            }
            qp += (long)FX_U * nx; wp += (long)FX_U * nx;
#pragma unroll
            for (int u = 0; u < FX_U; ++u) {
                const int jp = jb + u;
                double v; float vf;
                fx_value(qv[u], sgf, sg, v, vf);
                const double w = wv[u];
                if (v != v || w != w) continue;                  // NaN cell / NaN weight / past the segment
                long long NS, NV;                                // -X_S, -X_V
                fx_terms(v, w, fc, nsS, nsV, NS, NV);
                const uint32_t pk = lut[fx_bucket(vf, qminf, scalef)];
                int x = (int)(pk & 0xffffu), e = (int)(pk >> 16);
                const int e0 = e;                                // rows >= e0 have Q > v
                while (x < e) { const int mid = (x + e) >> 1; if (Qs[mid] < v) x = mid + 1; else e = mid; }
                int target = jp + 1;                             // inactive: cancels the own deposit
                if (x > jp + 1) { if (use_t1) target = x; }      // x = #{Q < v}
                else {
                    int h = x;                                   // #{Q <= v}; ties live in v's bucket only
                    if (h < e0 && Qs[h] == v) {
                        int y = e0; ++h;
                        while (h < y) { const int mid = (h + y) >> 1; if (Qs[mid] <= v) h = mid + 1; else y = mid; }
                    }
                    if (h <= jp && use_t2) target = h;
                }
                ownS -= NS; ownV -= NV;
                uint32_t* slot = fcol + target * FX_TC;
                fx_add64(slot, slot + plane, 0, NS);
                fx_add64(slot + 2 * plane, slot + 3 * plane, 0, NV);
            }
        }
    }
    __syncthreads();

    // ---- prefix down the columns: segment totals, block scan, final walk ----
    {
        unsigned long long aSl = 0, aVl = 0; long long aSh = 0, aVh = 0;
        const uint32_t* sl = fcol + r0 * FX_TC;
        for (int j = r0; j < r1; ++j, sl += FX_TC) {
            aSl += sl[0]; aSh += (int32_t)sl[plane]; aVl += sl[2 * plane]; aVh += (int32_t)sl[3 * plane];
        }
        tot[c * FX_TOTP + seg] = (long long)aSl + (aSh << 32) + ownS;
        tot[(FX_TC + c) * FX_TOTP + seg] = (long long)aVl + (aVh << 32) + ownV;
    }
    __syncthreads();
    if (warp < 2 * FX_TC) {                                      // warp = (which, column): exclusive scan over segments
        long long* row = tot + (size_t)warp * FX_TOTP;
        const long long a0 = row[2 * lane], a1 = row[2 * lane + 1];
        long long x = a0 + a1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(XC_FULL, x, o); if (lane >= o) x += t; }
        const long long ex = x - (a0 + a1);
        row[2 * lane] = ex; row[2 * lane + 1] = ex + a0;
    }
    __syncthreads();
    if (col_ok) {
        const double fiS = __ldg(&fp->iS), fiV = __ldg(&fp->iV);
        long long RS = tot[c * FX_TOTP + seg], RV = tot[(FX_TC + c) * FX_TOTP + seg];
        double* op = out + (s * (long)ny + r0) * nx + i;
        const QT* qp = qc; const double* wp = wc;
        const uint32_t* sl = fcol + r0 * FX_TC;
        const double* Qj = Qs + r0;
        for (int jb = r0; jb < r1; jb += FX_U) {
            QT qv[FX_U]; double wv[FX_U];
#pragma unroll
            for (int u = 0; u < FX_U; ++u) {
                const bool ok = jb + u < r1;
                qv[u] = ok ? __ldg(qp + (long)u * nx) : (QT)CUDART_NAN;
                wv[u] = ok ? __ldg(wp + (long)u * nx) : 0.0;
            }
            qp += (long)FX_U * nx; wp += (long)FX_U * nx;
#pragma unroll
            for (int u = 0; u < FX_U; ++u) {
                if (jb + u >= r1) break;
                RS += (long long)(((unsigned long long)sl[plane] << 32) | sl[0]);
                RV += (long long)(((unsigned long long)sl[3 * plane] << 32) | sl[2 * plane]);
                const double Sj = __dmul_rn(fx_to_double(RS), fiS), Vj = __dmul_rn(fx_to_double(RV), fiV);
                *op = sg * (Vj - (*Qj - fc) * Sj);
                op += nx; sl += FX_TC; ++Qj;
                double v; float vf;
                fx_value(qv[u], sgf, sg, v, vf);
                if (v == v && wv[u] == wv[u]) {
                    long long XS, XV;
                    fx_terms(v, wv[u], fc, fsS, fsV, XS, XV);
                    RS += XS; RV += XV;
                }
            }
        }
    }
}

}  // namespace xc

using namespace xc;

int xc::lwa_fx_prep_launch(long s0, long ns, int n_eq, const double* Qref, int increase, int32_t* sorted, int32_t* any_unsorted,
                           const double* rng, int rngC, const double* wmax_parts, int n_wmax, FxScale* fxs, uint32_t* lutg, void* stream)
{
    k_lwa_fx_prep<<<(unsigned)ns, FX_PREP_NT, 0, (cudaStream_t)stream>>>(s0, n_eq, Qref, increase, sorted, any_unsorted,
                                                                          rng, rngC, wmax_parts, n_wmax, fxs, lutg);
    XC_LAUNCH_OK();
    return 0;
}

bool xc::lwa_fx_fits(int n_eq) { return lwa_fx_layout(n_eq, 8).total <= 227 * 1024; }

template <typename QT, typename... A> static const QT* lwa_qptr(void (*)(const QT*, A...)) { return nullptr; }

int xc::lwa_fx_launch(const void* q, int q_dtype, long s0, long ns, int n_eq, int n_x, const double* Qref, const double* ww,
                      int increase, int part, const int32_t* sorted, const FxScale* fxs, const uint32_t* lutg, double* out, void* stream)
{
    const int fx_tc = lwa_fx_layout(n_eq, 16).total <= 227 * 1024 ? 16 : 8;
    const LwaFxSmem FL = lwa_fx_layout(n_eq, fx_tc);
    auto launch = [&](auto kern, int tcv) -> int {
        XC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FL.total));
        dim3 grid((unsigned)((n_x + tcv - 1) / tcv), (unsigned)ns);
        kern<<<grid, FX_SEG * tcv, FL.total, (cudaStream_t)stream>>>((decltype(lwa_qptr(kern)))q, s0, s0, n_eq, n_x, Qref, ww, increase, part, sorted, fxs, lutg, out);
        XC_LAUNCH_OK();
        return 0;
    };
    if (q_dtype == XC_F32) return fx_tc == 16 ? launch(k_lwa_fx<float, 16>, 16) : launch(k_lwa_fx<float, 8>, 8);
    return fx_tc == 16 ? launch(k_lwa_fx<double, 16>, 16) : launch(k_lwa_fx<double, 8>, 8);
}
