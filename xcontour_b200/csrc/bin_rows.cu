// Binning kernel of the fused Keff pass, second generation ("row march").
//
// Replaces, for fp32 tracers on grids whose cell area is constant along a row (every
// lat-lon, Cartesian and X-Z grid in BASELINE.json), what xhistogram + cumsum do in
// Contour2D.cal_integral_within_contours_hist (xcontour/core.py:412-460, _histogram
// core.py:1202-1325) for the two Keff integrands {dA, |grad q|^2 dA}, with |grad q|^2
// evaluated in flight (reference callers: tests/test_Keff_ocean.py:26-32,
// tests/test_clength.py:39-45).
//
// Why it looks the way it does (ncu of the round-1 kernel: 169 warp-instructions per 32
// cells, 55 % of them in the exponent-windowed accumulation, 72 % of the shared-memory
// wavefronts bank conflicts):
//   * a thread owns 4 consecutive columns and MARCHES DOWN the rows of its CTA: the
//     north/south neighbours are the previous / next iteration's registers, so a cell
//     costs one quarter of ONE 128-bit load and 1.5 fp32->fp64 conversions instead of
//     three loads and 3.5 conversions;
//   * the area is not accumulated at all: dA is constant along a row, so a cell only
//     bumps a 16-bit counter of its (row, bin) -- one ATOMS.ADD without return -- and
//     the bin's area is sum_rows dA[row] * count at the end (exact products);
//   * |grad q|^2 dA goes into ONE 96-bit fixed-point accumulator per (bin, copy): the
//     term is scaled by 2^k (k per CTA, folded into the row metrics, so scaling costs
//     nothing) such that the largest possible term of the CTA, (qmax-qmin)^2 (cx^2+cy^2) dA,
//     stays below 2^(96-h) with h = log2(cells per CTA); the three 32-bit words come
//     from two round-to-zero magic additions (no exponent extraction, no variable
//     shifts), the carries from add.cc/addc on the values the atomics return.  A term
//     d bits below that bound keeps min(53, 96-h-d) significant bits: full fp64
//     precision for |grad q| down to 2^-13 of the largest representable gradient and
//     1e-10 down to fp32-ulp sized differences.  Integer adds commute: the result is
//     bit-reproducible whatever the schedule;
//   * rows whose zonal metric dwarfs the meridional one (the pole rows of a lat-lon
//     grid: cx^2 > 2^30 cy^2) would waste that range; they take their scale from the
//     largest term actually present in the row (one block-wide max, two barriers) and
//     are flushed row by row -- still exact integer sums, still deterministic;
//   * COPIES accumulator copies selected by lane & (COPIES-1), laid out
//     [bin][copy] in 32-bit planes: lanes of a warp that sit in neighbouring bins hit
//     different banks and lanes in the same bin different words.
// One wave: the grid is sized to the resident CTA slots so every CTA runs concurrently
// and owns ceil(ny / row blocks) rows.
#include "common.cuh"
#include "internal.h"
#include <math_constants.h>

namespace xc {

constexpr int BR_MAXT = 384;

struct BinRowsParams {
    const float* q; int ny, nx; long s0;
    const double* edges; int N;                 // [S][N+1] ascending
    const double* cx; const double* cy;         // [ny] row metrics
    const double* dA_row;                       // [ny] cell area of each row
    const double* minmax;                       // [S][2] NaN-skipping (min, max) of each slice
    int bcx, bcy; float fill;
    int strips, rows_per, strip_w;              // column strips per row block, rows per CTA, columns per strip
    int uniform_dA;                             // every row has the same area: one count row
    int any_degenerate;                         // acc2 present
    int hbits;                                  // ceil(log2(cells per CTA))
    double* part;                               // [S][gridDim.x][2][N]
};

__device__ __forceinline__ int br_map(int i, int n, int bc)
{
    if ((unsigned)i < (unsigned)n) return i;
    if (bc == XC_BC_PERIODIC) return i < 0 ? i + n : i - n;
    if (bc == XC_BC_EXTEND) return i < 0 ? 0 : n - 1;
    if (bc == XC_BC_REFLECT) return i < 0 ? min(-i, n - 1) : max(2 * n - 2 - i, 0);
    return -1;                                                      // XC_BC_FILL
}

// smallest fp32 >= x: for an fp32 value v, (v >= x) == (v >= up32(x)) and (v < x) == (v < up32(x))
__device__ __forceinline__ float br_up32(double x)
{
    float f = __double2float_rn(x);
    if ((double)f < x) f = __int_as_float(__float_as_int(f) + (f >= 0.0f ? 1 : -1));
    if (f == 0.0f && x > 0.0) f = __int_as_float(1);
    return f;
}

// 96-bit add of floor(G), 0 <= G < 2^84, into the three words at shared address a, a + PLB, a + 2 PLB
template <int PLB>
__device__ __forceinline__ void br_add96(uint32_t a, double G)
{
    const double t = __dadd_rz(G, 0x1p84);                     // mantissa = floor(G / 2^32)
    const uint32_t v1 = (uint32_t)__double2loint(t);
    const uint32_t v2 = (uint32_t)__double2hiint(t) & 0xfffffu;
    const double r = __dsub_rn(G, __dsub_rn(t, 0x1p84));       // exact, in [0, 2^32)
    const uint32_t v0 = (uint32_t)__double2loint(__dadd_rz(r, 0x1p52));
    asm volatile("{\n\t.reg .u32 o0, o1, t1, k1, d;\n\t"
                 "atom.shared.add.u32 o0, [%0], %1;\n\t"
                 "add.cc.u32 d, o0, %1;\n\t"
                 "addc.cc.u32 t1, %2, 0;\n\t"
                 "addc.u32 k1, 0, 0;\n\t"
                 "atom.shared.add.u32 o1, [%0+%4], t1;\n\t"
                 "add.cc.u32 d, o1, t1;\n\t"
                 "addc.u32 d, %3, k1;\n\t"
                 "red.shared.add.u32 [%0+%5], d;\n\t}"
                 :: "r"(a), "r"(v0), "r"(v1), "r"(v2), "n"(PLB), "n"(2 * PLB) : "memory");
}
// a value the compiler must keep in a register (it would otherwise rebuild shared-window addresses from
// special registers at every use)
__device__ __forceinline__ uint32_t br_keep(uint32_t x) { uint32_t r; asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(x)); return r; }
__device__ __forceinline__ void br_red(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
// the same with a run-time plane stride (pole rows only)
__device__ __forceinline__ void br_add96_rt(uint32_t* w, int stride, double G)
{
    const double t = __dadd_rz(G, 0x1p84);
    const uint32_t v1 = (uint32_t)__double2loint(t);
    const uint32_t v2 = (uint32_t)__double2hiint(t) & 0xfffffu;
    const double r = __dsub_rn(G, __dsub_rn(t, 0x1p84));
    const uint32_t v0 = (uint32_t)__double2loint(__dadd_rz(r, 0x1p52));
    const uint32_t o0 = atomicAdd(w, v0);
    const unsigned long long t0 = (unsigned long long)o0 + v0;
    const unsigned long long t1 = (unsigned long long)v1 + (uint32_t)(t0 >> 32);
    const uint32_t o1 = atomicAdd(w + stride, (uint32_t)t1);
    const unsigned long long u1 = (unsigned long long)o1 + (uint32_t)t1;
    atomicAdd(w + 2 * stride, v2 + (uint32_t)(t1 >> 32) + (uint32_t)(u1 >> 32));
}
// value of a 96-bit accumulator (w2 may exceed 32 bits after summing copies)
__device__ __forceinline__ double br_value(unsigned long long w0, unsigned long long w1, unsigned long long w2)
{
    w1 += w0 >> 32; w2 += w1 >> 32;
    return fma((double)w2, 18446744073709551616.0, fma((double)(uint32_t)w1, 4294967296.0, (double)(uint32_t)w0));
}

struct BinRowsSmem { size_t e2, acc, esc, rowc, acc2, cnt, total; };
template <int PLW>
__host__ __device__ inline BinRowsSmem bin_rows_layout(int N, int rows_per, int uniform_dA, int any_deg)
{
    BinRowsSmem L; size_t o = 0;
    auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };   // every region 16-byte aligned (v2.f64 / uint4 accesses)
    L.e2 = o;   o = al(o + (size_t)N * 8);
    L.acc = o;  o = al(o + (size_t)3 * PLW * 4);
    L.esc = o;  o = al(o + (size_t)N * 8);
    L.rowc = o; o = al(o + (size_t)rows_per * 32);              // per row: cx', cy', dA, flag
    L.acc2 = o; o = al(o + (any_deg ? (size_t)3 * N * 4 : 0));
    L.cnt = o;  o = al(o + (size_t)(uniform_dA ? 1 : (rows_per + 1) / 2) * N * 4);
    L.total = o;
    return L;
}

__device__ __forceinline__ float2 br_lds_f2(uint32_t a)
{
    float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v;
}
__device__ __noinline__ int br_bin_slow(float v, uint32_t e2_sh, int N, int k)
{
    if (!(v == v)) return -1;
    for (;;) {
        const float2 e = br_lds_f2(e2_sh + (uint32_t)k * 8u);
        if (v < e.x) { if (--k < 0) return -1; }
        else if (v >= e.y) { if (++k >= N) return -1; }
        else return k;
    }
}
// bin p with e[p] <= v < e[p+1] (-1 outside / NaN): arithmetic guess, verified against the true thresholds
__device__ __forceinline__ int br_bin(float v, uint32_t e2_sh, int N, float e0, float inv)
{
    int k = __float2int_rz((v - e0) * inv);
    k = min(max(k, 0), N - 1);
    const float2 e = br_lds_f2(e2_sh + (uint32_t)k * 8u);
    if (v >= e.x && v < e.y) return k;
    return br_bin_slow(v, e2_sh, N, k);
}

struct BrRow { float4 raw; double d[4]; };
__device__ __forceinline__ void br_cvt(BrRow& r)
{
    r.d[0] = (double)r.raw.x; r.d[1] = (double)r.raw.y; r.d[2] = (double)r.raw.z; r.d[3] = (double)r.raw.w;
}

template <int COPIES, int NPAD>
__global__ void __launch_bounds__(BR_MAXT, 2)
k_bin_rows(const BinRowsParams p)
{
    constexpr int PLW = COPIES * NPAD;
    constexpr int PLB = PLW * 4;
    extern __shared__ __align__(16) unsigned char smem[];
    const int N = p.N, ny = p.ny, nx = p.nx, NT = blockDim.x;
    const BinRowsSmem L = bin_rows_layout<PLW>(N, p.rows_per, p.uniform_dA, p.any_degenerate);
    float2*   e2   = reinterpret_cast<float2*>(smem + L.e2);
    uint32_t* acc  = reinterpret_cast<uint32_t*>(smem + L.acc);
    double*   esc  = reinterpret_cast<double*>(smem + L.esc);
    double*   rowc = reinterpret_cast<double*>(smem + L.rowc);     // [row][4]: cx', cy', dA, flag
    uint32_t* acc2 = reinterpret_cast<uint32_t*>(smem + L.acc2);
    uint32_t* cnt  = reinterpret_cast<uint32_t*>(smem + L.cnt);
    __shared__ int s_emax;
    __shared__ unsigned s_hmax;

    const int tid = threadIdx.x, lane = tid & 31;
    const long s = p.s0 + blockIdx.y;
    const int strip = blockIdx.x % p.strips, rb = blockIdx.x / p.strips;
    const int r0 = rb * p.rows_per, r1 = min(ny, r0 + p.rows_per), nrows = r1 - r0;
    const int cbeg = strip * p.strip_w, cend = min(nx, cbeg + p.strip_w);
    const int col = cbeg + 4 * tid;
    const bool act = col < cend;
    const int cnt_rows = p.uniform_dA ? 1 : (p.rows_per + 1) / 2;

    // ---- per-CTA setup: thresholds, zeroed accumulators, scaled row metrics ----
    const double* eg = p.edges + s * (long)(N + 1);
    for (int k = tid; k < N; k += NT) e2[k] = make_float2(br_up32(eg[k]), br_up32(eg[k + 1]));
    for (int i = tid; i < 3 * PLW; i += NT) acc[i] = 0u;
    for (int i = tid; i < N; i += NT) esc[i] = 0.0;
    if (p.any_degenerate) for (int i = tid; i < 3 * N; i += NT) acc2[i] = 0u;
    for (int i = tid; i < cnt_rows * N; i += NT) cnt[i] = 0u;
    if (tid == 0) { s_emax = -100000; s_hmax = 0u; }
    __syncthreads();
    const double qlo = p.minmax[2 * s], qhi = p.minmax[2 * s + 1];
    double Rq = qhi - qlo;
    if (p.bcx == XC_BC_FILL || p.bcy == XC_BC_FILL) Rq = fmax(Rq, fmax(fabs(qhi - (double)p.fill), fabs(qlo - (double)p.fill)));
    if (!(Rq >= 0.0) || !isfinite(Rq)) Rq = 0.0;
    for (int r = tid; r < nrows; r += NT) {
        const double cx = p.cx[r0 + r], cy = p.cy[r0 + r];
        double a = p.dA_row[r0 + r];
        if (!(a == a)) a = 0.0;                                        // fillna(0) of wei, core.py:446-449
        const double b = Rq * Rq * (cx * cx + cy * cy) * fabs(a) * 1.000001;
        const bool deg = !isfinite(b) || !isfinite(cx) || !isfinite(cy) || (cx * cx > 1073741824.0 * cy * cy && cy != 0.0);
        rowc[4 * r + 2] = a;
        rowc[4 * r + 3] = deg ? 1.0 : 0.0;
        if (!deg && b > 0.0) atomicMax(&s_emax, ilogb(b) + 1);
    }
    __syncthreads();
    int ksc = 0;                                                         // G = 2^ksc * |grad q|^2 dA
    if (s_emax > -100000) { ksc = 96 - p.hbits - s_emax; ksc = (ksc >= 0 ? ksc : ksc - 1) / 2 * 2; ksc = max(-1600, min(1600, ksc)); }
    for (int r = tid; r < nrows; r += NT) {
        const int h = rowc[4 * r + 3] != 0.0 ? 0 : ksc / 2;
        rowc[4 * r] = scalbn(p.cx[r0 + r], h);
        rowc[4 * r + 1] = scalbn(p.cy[r0 + r], h);
    }
    const unsigned hi_limit = (unsigned)(1023 + 96 - p.hbits) << 20;   // hi word of 2^(96-h)
    const float e0 = br_up32(eg[0]);
    const double span = eg[N] - eg[0];
    const float inv = (span > 0.0 && isfinite(span)) ? (float)((double)N / span) : 0.0f;
    __syncthreads();

    // ---- the march ----
    const float* qs = p.q + s * (long)ny * nx;
    const float fillv = p.fill;
    const float4 fill4 = make_float4(fillv, fillv, fillv, fillv);
    auto ld4 = [&](int j) -> float4 {                                   // any row index, ghost rows included
        const int jj = br_map(j, ny, p.bcy);
        if (!act) return make_float4(0.f, 0.f, 0.f, 0.f);
        if (jj < 0) return fill4;
        return __ldg(reinterpret_cast<const float4*>(qs + (long)jj * nx + col));
    };
    const bool wneed = act && lane == 0;
    const bool eneed = act && (lane == 31 || col + 4 >= cend);
    const int cw = br_map(col - 1, nx, p.bcx), ce = br_map(col + 4, nx, p.bcx);
    const uint32_t e2_sh = br_keep((uint32_t)__cvta_generic_to_shared(e2));
    const uint32_t acc_sh = br_keep((uint32_t)__cvta_generic_to_shared(acc) + (uint32_t)(lane & (COPIES - 1)) * 4u);
    const uint32_t cnt_sh = br_keep((uint32_t)__cvta_generic_to_shared(cnt));
    const uint32_t N4 = (uint32_t)N * 4u;

    // running pointers: pq -> (row j+2, col), pw / pe -> halo cells of row j+1
    const float* pq = qs + (long)min(r0 + 2, ny - 1) * nx + (act ? col : 0);
    const float* pw = qs + (long)min(r0 + 1, ny - 1) * nx + (cw < 0 ? 0 : cw);
    const float* pe = qs + (long)min(r0 + 1, ny - 1) * nx + (ce < 0 ? 0 : ce);

    BrRow A, B, C;                                                      // roles rotate: (prev, cur, next)
    float wraw = 0.f, eraw = 0.f;
    A.raw = ld4(r0 - 1); br_cvt(A);
    B.raw = ld4(r0); br_cvt(B);
    C.raw = ld4(r0 + 1);
    if (wneed) wraw = cw < 0 ? fillv : __ldg(qs + (long)r0 * nx + cw);
    if (eneed) eraw = ce < 0 ? fillv : __ldg(qs + (long)r0 * nx + ce);

    uint32_t rc_sh = br_keep((uint32_t)__cvta_generic_to_shared(rowc));      // row constants of the current row
    const uint32_t esc_sh = 0; (void)esc_sh;
    auto slow_cell = [&](int bin, double G, bool degenerate_row) {            // everything that is not the common case
        if (bin < 0 || !(G == G)) return;
        if (degenerate_row) { atomicAdd(esc + bin, G); return; }
        if ((unsigned)__double2hiint(G) < hi_limit) br_add96<PLB>(acc_sh + (uint32_t)bin * (COPIES * 4u), G);
        else atomicAdd(esc + bin, scalbn(G, -ksc));                            // inf / out of range (never on sane data)
    };
    auto pole_cell = [&](int bin, double G, int k2) {
        if (bin < 0) return;
        if ((unsigned)__double2hiint(G) < 0x7ff00000u) br_add96_rt(acc2 + bin, N, scalbn(G, k2));
        else if (G == G) atomicAdd(esc + bin, G);
    };
    auto step = [&](BrRow& P, BrRow& Cu, BrRow& Nx, const int j) {
        const int r = j - r0;
        br_cvt(Nx);
        float wv = __shfl_up_sync(XC_FULL, Cu.raw.w, 1), ev = __shfl_down_sync(XC_FULL, Cu.raw.x, 1);
        if (wneed) wv = wraw;
        if (eneed) ev = eraw;
        const float4 cur = Cu.raw;
        // prefetch row j+2 into the slot whose fp64 image is still in use as `prev`, and the halo of row j+1
        if (j + 2 < r1 || (j + 2 == r1 && r1 < ny)) { if (act) P.raw = __ldg(reinterpret_cast<const float4*>(pq)); }
        else if (j + 2 == r1) P.raw = ld4(ny);
        if (j + 1 < r1) {
            if (wneed) wraw = cw < 0 ? fillv : __ldg(pw);
            if (eneed) eraw = ce < 0 ? fillv : __ldg(pe);
        }
        pq += nx; pw += nx; pe += nx;
        double cxs, cys, das, flag;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(cxs), "=d"(cys) : "r"(rc_sh));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(das), "=d"(flag) : "r"(rc_sh));
        rc_sh += 32u;
        const double Wd = (double)wv, Ed = (double)ev;
        double G[4];
        {
            const double x0 = __dmul_rn(__dsub_rn(Cu.d[1], Wd), cxs),      y0 = __dmul_rn(__dsub_rn(Nx.d[0], P.d[0]), cys);
            const double x1 = __dmul_rn(__dsub_rn(Cu.d[2], Cu.d[0]), cxs), y1 = __dmul_rn(__dsub_rn(Nx.d[1], P.d[1]), cys);
            const double x2 = __dmul_rn(__dsub_rn(Cu.d[3], Cu.d[1]), cxs), y2 = __dmul_rn(__dsub_rn(Nx.d[2], P.d[2]), cys);
            const double x3 = __dmul_rn(__dsub_rn(Ed, Cu.d[2]), cxs),      y3 = __dmul_rn(__dsub_rn(Nx.d[3], P.d[3]), cys);
            G[0] = __dmul_rn(__dadd_rn(__dmul_rn(x0, x0), __dmul_rn(y0, y0)), das);
            G[1] = __dmul_rn(__dadd_rn(__dmul_rn(x1, x1), __dmul_rn(y1, y1)), das);
            G[2] = __dmul_rn(__dadd_rn(__dmul_rn(x2, x2), __dmul_rn(y2, y2)), das);
            G[3] = __dmul_rn(__dadd_rn(__dmul_rn(x3, x3), __dmul_rn(y3, y3)), das);
        }
        int b[4];
        b[0] = act ? br_bin(cur.x, e2_sh, N, e0, inv) : -1;
        b[1] = act ? br_bin(cur.y, e2_sh, N, e0, inv) : -1;
        b[2] = act ? br_bin(cur.z, e2_sh, N, e0, inv) : -1;
        b[3] = act ? br_bin(cur.w, e2_sh, N, e0, inv) : -1;
        // area: one 16-bit counter per (row, bin), two rows per word
        const uint32_t crow = p.uniform_dA ? cnt_sh : cnt_sh + (uint32_t)(r >> 1) * N4;
        const uint32_t cinc = p.uniform_dA ? 1u : 1u << ((r & 1) << 4);
        const bool normal = flag == 0.0;
        // common case of a warp step: every lane's four cells are binned and every term is in range
        // (lanes past the last column of the strip have nothing to add and do not spoil the vote)
        const bool ok = !act || ((b[0] | b[1] | b[2] | b[3]) >= 0 &&
                        (unsigned)__double2hiint(G[0]) < hi_limit && (unsigned)__double2hiint(G[1]) < hi_limit &&
                        (unsigned)__double2hiint(G[2]) < hi_limit && (unsigned)__double2hiint(G[3]) < hi_limit);
        if (normal && __all_sync(XC_FULL, ok)) {
            // smooth stretch of the field: every lane's four cells sit in one bin -> one counter bump and one 96-bit
            // add per lane instead of four (the lanes of such a step share one or two bins, so the atomics would
            // serialise on them); the four terms are summed in fp64 first, in a fixed order
            const bool one = !act || (b[0] == b[1] && b[1] == b[2] && b[2] == b[3]);
            if (__all_sync(XC_FULL, one)) {
                if (act) {
                    br_red(crow + (uint32_t)b[0] * 4u, cinc * 4u);
                    br_add96<PLB>(acc_sh + (uint32_t)b[0] * (COPIES * 4u), __dadd_rn(__dadd_rn(G[0], G[1]), __dadd_rn(G[2], G[3])));
                }
                return;
            }
            if (act) {
#pragma unroll
                for (int c = 0; c < 4; ++c) br_red(crow + (uint32_t)b[c] * 4u, cinc);
#pragma unroll
                for (int c = 0; c < 4; ++c) br_add96<PLB>(acc_sh + (uint32_t)b[c] * (COPIES * 4u), G[c]);
            }
            return;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) if (b[c] >= 0) br_red(crow + (uint32_t)b[c] * 4u, cinc);
        if (normal || !p.any_degenerate) {
            slow_cell(b[0], G[0], !normal); slow_cell(b[1], G[1], !normal);
            slow_cell(b[2], G[2], !normal); slow_cell(b[3], G[3], !normal);
            return;
        }
        // pole-like row: scale from the largest term actually present, flushed at once
        unsigned hm = 0u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const unsigned h = (unsigned)__double2hiint(G[c]);
            if (b[c] >= 0 && h < 0x7ff00000u) hm = max(hm, h);
        }
        hm = __reduce_max_sync(XC_FULL, hm);
        if (lane == 0 && hm) atomicMax(&s_hmax, hm);
        __syncthreads();
        const int ex = (int)(s_hmax >> 20) - 1023;                       // largest term < 2^(ex+1)
        const int k2 = 96 - p.hbits - (ex + 1);
        pole_cell(b[0], G[0], k2); pole_cell(b[1], G[1], k2); pole_cell(b[2], G[2], k2); pole_cell(b[3], G[3], k2);
        __syncthreads();
        for (int n = tid; n < N; n += NT) {
            const double v = br_value(acc2[n], acc2[N + n], acc2[2 * N + n]);
            if (v != 0.0) esc[n] += scalbn(v, -k2);
            acc2[n] = 0u; acc2[N + n] = 0u; acc2[2 * N + n] = 0u;
        }
        if (tid == 0) s_hmax = 0u;
        __syncthreads();
    };
    for (int j = r0; j < r1; j += 3) {
        step(A, B, C, j);
        if (j + 1 < r1) step(B, C, A, j + 1);
        if (j + 2 < r1) step(C, A, B, j + 2);
    }
    __syncthreads();

    // ---- flush: this CTA's partial sums, bin by bin ----
    double* out = p.part + ((size_t)(blockIdx.y + p.s0) * gridDim.x + blockIdx.x) * 2 * (size_t)N;
    for (int n = tid; n < N; n += NT) {
        double area = 0.0;
        if (p.uniform_dA) area = __dmul_rn(rowc[2], (double)cnt[n]);
        else
            for (int r = 0; r < nrows; ++r) {
                const uint32_t cn = (cnt[(size_t)(r >> 1) * N + n] >> ((r & 1) << 4)) & 0xffffu;
                if (cn) area += rowc[4 * r + 2] * (double)cn;
            }
        unsigned long long w0 = 0, w1 = 0, w2 = 0;
#pragma unroll
        for (int c = 0; c < COPIES; ++c) {
            w0 += acc[n * COPIES + c]; w1 += acc[PLW + n * COPIES + c]; w2 += acc[2 * PLW + n * COPIES + c];
        }
        out[n] = area;
        out[N + n] = scalbn(br_value(w0, w1, w2), -ksc) + esc[n];
    }
}

}  // namespace xc

using namespace xc;

namespace {
struct BinRowsPlan { bool ok; bool small; int NT, strips, strip_w, rows_per, hbits; long C; size_t smem; };
// uniform_dA / any_degenerate < 0: unknown (workspace sizing) -> the combination that needs the most shared memory
BinRowsPlan bin_rows_plan(long S, int ny, int nx, int N, int uniform_dA, int any_degenerate)
{
    BinRowsPlan pl; pl.ok = false;
    if ((nx & 3) || nx < 8 || ny < 2 || N < 1 || N > 2048 || S < 1) return pl;
    const int uni = uniform_dA < 0 ? 0 : uniform_dA, deg = any_degenerate < 0 ? 1 : any_degenerate;
    pl.small = N <= 512;
    if (nx <= 4 * BR_MAXT) { pl.NT = ((nx / 4 + 31) / 32) * 32; pl.strips = 1; }
    else { pl.NT = 256; pl.strips = (nx + 4 * pl.NT - 1) / (4 * pl.NT); }
    pl.strip_w = 4 * pl.NT;
    const long slots = (long)sm_count() * 2;
    long per_slice = slots / S; if (per_slice < 1) per_slice = 1;
    long rbs = per_slice / pl.strips; if (rbs < 1) rbs = 1;
    int rows_per = (int)((ny + rbs - 1) / rbs);
    if (rows_per < 8) rows_per = ny < 8 ? ny : 8;                  // the two halo rows of a march stay a small share
    auto lay = [&](int rp) {
        return pl.small ? bin_rows_layout<4 * 512>(N, rp, uni, deg).total : bin_rows_layout<2 * 2048>(N, rp, uni, deg).total;
    };
    const size_t budget = 110 * 1024;                                // two CTAs per SM
    while (lay(rows_per) > budget) {                                 // strictly decreasing: 111, 82, 60, ... 4, 2
        if (rows_per <= 2) return pl;
        const int next = (rows_per * 3 / 4) & ~1;
        rows_per = next >= 2 ? next : 2;
    }
    pl.rows_per = rows_per;
    pl.C = (long)((ny + rows_per - 1) / rows_per) * pl.strips;
    pl.smem = lay(rows_per);
    const long cells = (long)rows_per * pl.strip_w;
    pl.hbits = 14; while ((1L << pl.hbits) < cells) ++pl.hbits;     // >= 14: a thread's four summed terms stay below 2^84 (br_add96)
    pl.ok = pl.C <= 65535 && pl.hbits <= 30 && pl.strip_w <= 65535;
    return pl;
}
}  // namespace

// doubles of per-CTA partials the row-march kernel may need for S slices (0: it never applies)
size_t xc::bin_rows_part_doubles(long S, int ny, int nx, int N)
{
    const BinRowsPlan pl = bin_rows_plan(S, ny, nx, N, -1, -1);
    return pl.ok ? (size_t)S * pl.C * 2 * N : 0;
}

// Plans and launches the row-march kernel.  Returns 0 launched (C_out = CTAs per slice), 1 not applicable, 2 error.
int xc::bin_rows_try(const void* q, int q_dtype, long S, const double* edges, int N,
                     const StencilArgs* st, const double* minmax, double* part, size_t part_doubles,
                     int* C_out, void* stream)
{
    if (q_dtype != XC_F32 || !st || !st->dA_row || !minmax) return 1;
    const int ny = st->ny, nx = st->nx;
    if ((((uintptr_t)q) & 15)) return 1;
    static const char* off = getenv("XCB200_NO_BIN_ROWS");
    if (off) return 1;
    BinRowsPlan pl = bin_rows_plan(S, ny, nx, N, st->uniform_dA, st->any_degenerate);
    if (!pl.ok) return 1;
    if ((size_t)S * pl.C * 2 * N > part_doubles) {                   // sized for the worst-case flags: cannot happen
        set_error("k_bin_rows: partial buffer too small (%zu < %zu doubles)", part_doubles, (size_t)S * pl.C * 2 * N); return 2;
    }
    BinRowsParams p;
    p.q = (const float*)q; p.ny = ny; p.nx = nx; p.edges = edges; p.N = N;
    p.cx = st->cx; p.cy = st->cy; p.dA_row = st->dA_row; p.minmax = minmax;
    p.bcx = st->bcx; p.bcy = st->bcy; p.fill = st->fill;
    p.strips = pl.strips; p.rows_per = pl.rows_per; p.strip_w = pl.strip_w;
    p.uniform_dA = st->uniform_dA; p.any_degenerate = st->any_degenerate; p.hbits = pl.hbits; p.part = part;
    auto kern = pl.small ? k_bin_rows<4, 512> : k_bin_rows<2, 2048>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess) {
        set_error("k_bin_rows: cannot reserve %zu bytes of shared memory", pl.smem); return 2;
    }
    for (long s0 = 0; s0 < S; s0 += 65535) {
        const long ns = S - s0 < 65535 ? S - s0 : 65535;
        p.s0 = s0;
        kern<<<dim3((unsigned)pl.C, (unsigned)ns), pl.NT, pl.smem, (cudaStream_t)stream>>>(p);
        count_launch();
        if (cudaGetLastError() != cudaSuccess) { set_error("k_bin_rows launch failed"); return 2; }
    }
    *C_out = (int)pl.C;
    return 0;
}
