// Kernel (4): local wave activity / local APE column integrals
// (Contour2D.cal_local_wave_activity, xcontour/core.py:696-799; cal_local_APE,
// core.py:908-942; cal_local_wave_activity2, core.py:802-905).
//
// The reference loops over every row j of the equivalent dimension and reduces a
// masked full slice each time: O(n_eq^2 * n_x).  For a profile Q that is sorted
// in the direction `increase` implies, every cell (j', i) contributes
// sign*(v - Q_j)*w to ONE contiguous j-range of its own column,
//     j' < j < lo(v)      when lo(v) = #{Q < v} > j'+1     (mask -1 region)
//     hi(v) <= j <= j'    when hi(v) = #{Q <= v} <= j'     (mask +1 region)
// so a column is two difference arrays (sum w, sum w*v) followed by one prefix
// sum:  LWA[j] = V[j] - Q_j * S[j].   One warp owns one column (difference
// arrays in shared memory; rows are staged 96 at a time through a transposed
// tile so that global loads stay coalesced); lo/hi come from a 2048-bucket lookup
// table over Q plus a short exact search (comparisons against Q decide, never
// arithmetic).  The slot j'+1 of each cell is updated without conflicts; the other
// end of the range is a warp-private scatter whose colliding lanes are serialised
// by a byte-tag election (default) or a MATCH.ANY peel, as in hist.cu.
// Profiles that are not sorted (or contain NaN) take the exact O(n_eq^2) kernel;
// variant 2 (cal_local_wave_activity2) has its own gather-only kernel.
// Bound: shared-memory scatter throughput, not HBM (profiles/README.md).
#include "lwa_fx.cuh"
#include <stdlib.h>

namespace xc {

constexpr int LWA_LUT = 2048;
constexpr int LWA_MAX_TC = 16;

__global__ void k_check_sorted(const double* __restrict__ Q, int ny, int increase,
                               int32_t* __restrict__ flag)
{
    const long s = blockIdx.x;
    const double sg = increase ? 1.0 : -1.0;
    const double* Qs = Q + s * ny;
    int bad = 0;
    for (int j = threadIdx.x; j < ny; j += blockDim.x) {
        double a = sg * Qs[j];
        if (isnan(a)) bad = 1;
        if (j + 1 < ny && !(a <= sg * Qs[j + 1])) bad = 1;
    }
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) flag[s] = bad ? 0 : 1;
}

// Rows are staged in blocks of LWA_BIG = 32*LWA_NI.  Inside a block lane l works
// on rows {NI*l + u : u < NI}: an odd stride between the lanes of one item set
// keeps the 128-bit accesses to the difference array bank-conflict free and
// makes two lanes of a set share a far-end slot less often than adjacent rows do.
#ifndef XC_LWA_NI
#define XC_LWA_NI 3
#endif
constexpr int LWA_NI  = XC_LWA_NI;
constexpr int LWA_BIG = 32 * LWA_NI;       // rows per staged block
// staging strides chosen so that the transposed store [col][row] of a 2-row x 16-col
// warp tile and the stride-NI column reads are both free of bank conflicts
constexpr int LWA_RSW = LWA_BIG + 1;       // ww (doubles): odd
constexpr int LWA_RSQ = LWA_BIG + 2;       // q: == 2 (mod 32) for BIG = 96

struct LwaSmem {
    size_t off_Q, off_D, off_lut, off_tag, off_sq, off_sw, total;
    int nyp, tagp;
};
static __host__ __device__ inline LwaSmem lwa_layout(int ny, int TC, int qbytes, bool tags)
{
    LwaSmem L;
    L.nyp = ny + 2;
    L.tagp = (ny + 2 + 15) & ~15;
    size_t o = 0;
    L.off_Q = o;   o += (size_t)((ny + 1) & ~1) * 8;
    L.off_D = o;   o += (size_t)TC * L.nyp * 16;
    L.off_sw = o;  o += (size_t)TC * LWA_RSW * 8;
    L.off_sq = o;  o += (size_t)TC * LWA_RSQ * qbytes;
    o = (o + 15) & ~(size_t)15;
    L.off_lut = o; o += (size_t)(LWA_LUT + 1) * 4;
    o = (o + 15) & ~(size_t)15;
    L.off_tag = o; if (tags) o += (size_t)TC * L.tagp;
    L.total = o;
    return L;
}

// fp32 bucket of a value; monotone in its argument, so bucket(Q_j) < bucket(v)
// implies Q_j < v (and > implies >): the LUT only narrows the range, the exact
// fp64 comparisons against Q decide.
__device__ __forceinline__ int lwa_bucket(float vf, float qminf, float scalef)
{
    const float t = fminf(fmaxf((vf - qminf) * scalef, 0.0f), (float)(LWA_LUT - 1));
    return __float2int_rz(t);
}

// Many lanes share a slot (homogenised regions): combine them in registers by
// pointer jumping over the MATCH.ANY peer list (see hist.cu), then one
// read-modify-write per distinct slot.  Out of line: it is the rare path.
__device__ __noinline__ void lwa_scatter_heavy(double2* Dw, int target, double s0, double s1, bool a, int lane)
{
    unsigned pr = __match_any_sync(XC_FULL, a ? (unsigned)target : (0x80000000u | (unsigned)lane));
    if (!a) pr = 0u;
    const unsigned above = (lane == 31) ? 0u : (pr & (0xffffffffu << (lane + 1)));
    int nxt = above ? (__ffs(above) - 1) : -1;
    const bool leader = a && ((__ffs(pr) - 1) == lane);
    while (__any_sync(XC_FULL, nxt >= 0)) {
        const int src = nxt >= 0 ? nxt : lane;
        const double g0 = __shfl_sync(XC_FULL, s0, src), g1 = __shfl_sync(XC_FULL, s1, src);
        const int gn = __shfl_sync(XC_FULL, nxt, src);
        if (nxt >= 0) { s0 += g0; s1 += g1; nxt = gn; }
    }
    if (leader) { double2 t = Dw[target]; t.x += s0; t.y += s1; Dw[target] = t; }
    __syncwarp();
}

// Warp-private scatter-add of (nw, nwv) into Dw[target] for NI items per lane;
// lanes of one item set that share a target are serialised (MATCH.ANY peel or
// byte tags, see hist.cu).
template <bool MATCH, int HEAVY>
__device__ __forceinline__ void lwa_scatter(double2* Dw, uint8_t* tagw, const int (&target)[LWA_NI],
                                            const double (&nw)[LWA_NI], const double (&nwv)[LWA_NI],
                                            const bool (&act)[LWA_NI], int lane)
{
    if (MATCH) {
#pragma unroll
        for (int u = 0; u < LWA_NI; ++u) {
            unsigned pr = __match_any_sync(XC_FULL, act[u] ? (unsigned)target[u] : (0x80000000u | (unsigned)lane));
            if (!act[u]) pr = 0u;
            do {
                if (pr && (__ffs(pr) - 1) == lane) {
                    double2 t = Dw[target[u]]; t.x += nw[u]; t.y += nwv[u]; Dw[target[u]] = t;
                }
                pr &= pr - 1u;
                __syncwarp();
            } while (__any_sync(XC_FULL, pr != 0u));
        }
    } else {
#pragma unroll
        for (int u = 0; u < LWA_NI; ++u) {
            bool a = act[u];
            unsigned pending = __ballot_sync(XC_FULL, a);
            for (int round = 0; pending && (HEAVY == 0 || round < HEAVY); ++round) {
                if (a) tag_store(tagw + target[u], (unsigned)lane);
                __syncwarp();
                if (a && tag_load(tagw + target[u]) == (unsigned)lane) {
                    double2 t = Dw[target[u]]; t.x += nw[u]; t.y += nwv[u]; Dw[target[u]] = t;
                    a = false;
                }
                __syncwarp();
                pending = __ballot_sync(XC_FULL, a);
            }
            if (HEAVY > 0 && pending) lwa_scatter_heavy(Dw, target[u], nw[u], nwv[u], a, lane);   // homogenised region
        }
    }
}

// grid = (ceil(nx/TC), nslices), block = TC warps; warp w owns column i0 + w.
#define XC_LWA_BOUNDS __launch_bounds__(LWA_MAX_TC * 32, 1)
#define LWA_ROW(lane, u) (LWA_NI * (lane) + (u))

template <typename QT, bool MATCH, int HEAVY>
__global__ void XC_LWA_BOUNDS
k_lwa_fast(const QT* __restrict__ q, long s0, int ny, int nx,
           const double* __restrict__ Qref, const double* __restrict__ ww,
           int increase, int part, const int32_t* __restrict__ sorted,
           double* __restrict__ out, int TC)
{
    const long s = s0 + blockIdx.y;
    if (!sorted[s]) return;
    extern __shared__ __align__(16) unsigned char smem[];
    const LwaSmem L = lwa_layout(ny, TC, (int)sizeof(QT), !MATCH);
    double*   Qs  = reinterpret_cast<double*>(smem + L.off_Q);
    double2*  D   = reinterpret_cast<double2*>(smem + L.off_D);
    double*   sw  = reinterpret_cast<double*>(smem + L.off_sw);    // [TC][LWA_RSW]
    QT*       sq  = reinterpret_cast<QT*>(smem + L.off_sq);        // [TC][LWA_RSQ]
    uint32_t* lut = reinterpret_cast<uint32_t*>(smem + L.off_lut); // lut[b] | lut[b+1] << 16
    uint8_t*  tag = smem + L.off_tag;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
    const double sg = increase ? 1.0 : -1.0;
    const int i0 = blockIdx.x * TC;
    const QT* qs = q + s * (long)ny * nx;
    const double* Qg = Qref + s * (long)ny;

    // staging map: element k*nthr + tid of a BIG x TC block -> row 32*k + tid/TC,
    // col tid%TC (nthr = 32*TC): a warp reads 32/TC... rows of TC contiguous values
    const int rr = tid / TC, cc = tid - rr * TC;
    const bool col_ok = (i0 + cc) < nx;
    const QT*     gq = qs + (long)rr * nx + i0 + cc;               // element (row rr, col cc) of block 0
    const double* gw = ww + (long)rr * nx + i0 + cc;
    const long    gstep = 32L * nx;
    QT pq[LWA_NI]; double pw[LWA_NI];
    auto fetch = [&](int b) {
        const QT* a = gq + (long)b * LWA_NI * gstep;
        const double* c = gw + (long)b * LWA_NI * gstep;
        const int jb = b * LWA_BIG + rr;
#pragma unroll
        for (int k = 0; k < LWA_NI; ++k) {
            if (jb + 32 * k < ny && col_ok) { pq[k] = __ldg(a + k * gstep); pw[k] = __ldg(c + k * gstep); }
            else { pq[k] = (QT)CUDART_NAN; pw[k] = 0.0; }
        }
    };
    fetch(0);

    for (int j = tid; j < ny; j += nthr) Qs[j] = sg * Qg[j];
    __syncthreads();
    const double qmin = Qs[0], qmax = Qs[ny - 1];
    const float qminf = (float)qmin;
    const float scalef = (qmax > qmin) ? (float)((double)LWA_LUT / (qmax - qmin)) : 0.0f;
    // first j whose bucket is >= b, for b = 0..LWA_LUT (buckets are monotone in Q);
    // entry b of the packed table holds (first[b], first[b+1])
    {
        uint16_t* first = reinterpret_cast<uint16_t*>(D);           // D is zeroed right after
        for (int j = tid; j <= ny; j += nthr) {
            int bj = (j < ny) ? lwa_bucket((float)Qs[j], qminf, scalef) : LWA_LUT;
            int bp = (j > 0) ? lwa_bucket((float)Qs[j - 1], qminf, scalef) : -1;
            for (int b = bp + 1; b <= bj; ++b) first[b] = (uint16_t)j;
        }
        __syncthreads();
        for (int b = tid; b < LWA_LUT; b += nthr) lut[b] = (uint32_t)first[b] | ((uint32_t)first[b + 1] << 16);
        __syncthreads();
        for (int k = tid; k < TC * L.nyp; k += nthr) D[k] = make_double2(0.0, 0.0);
        // (the sync after the first staging store below also covers D)
    }

    // which mask regions are integrated (core.py:773-784)
    const bool keep_pos = (part == XC_PART_UPPER) == (increase != 0);
    const bool use_t1 = (part == XC_PART_ALL) || !keep_pos;   // mask -1 region
    const bool use_t2 = (part == XC_PART_ALL) || keep_pos;    // mask +1 region

    double2* Dw = D + (size_t)warp * L.nyp;
    uint8_t* tagw = tag + (size_t)warp * L.tagp;
    const QT*     sqw = sq + (size_t)warp * LWA_RSQ;
    const double* sww = sw + (size_t)warp * LWA_RSW;
    const int nblk = (ny + LWA_BIG - 1) / LWA_BIG;

    for (int b = 0; b < nblk; ++b) {
        if (b > 0) __syncthreads();                          // everyone is done with the staging buffers
#pragma unroll
        for (int k = 0; k < LWA_NI; ++k) {                   // transposed store: [col][row]
            sq[cc * LWA_RSQ + 32 * k + rr] = pq[k];
            sw[cc * LWA_RSW + 32 * k + rr] = pw[k];
        }
        __syncthreads();
        if (b + 1 < nblk) fetch(b + 1);                      // next block's loads fly during the work below

        const int jblk = b * LWA_BIG;
        int tgt[LWA_NI]; bool act[LWA_NI]; double nw[LWA_NI], nwv[LWA_NI];
#pragma unroll
        for (int u = 0; u < LWA_NI; ++u) {
            const int jp = jblk + LWA_ROW(lane, u);
            const QT qraw = sqw[LWA_ROW(lane, u)];
            const double v = sg * (double)qraw;
            const double w = sww[LWA_ROW(lane, u)];
            bool a = (jp < ny) && (v == v);
            int target = 0;
            {
                const uint32_t pk = lut[lwa_bucket((float)v, qminf, scalef)];
                int x = (int)(pk & 0xffffu), e = (int)(pk >> 16);
                const int e0 = e;                            // end of v's bucket: rows >= e0 have Q > v
                while (x < e) {                              // wide buckets (flat stretches of Q)
                    const int mid = (x + e) >> 1;
                    if (Qs[mid] < v) x = mid + 1; else e = mid;
                }
                const int lo = x;                            // #{Q < v}
                if (lo > jp + 1) { target = lo; a = a && use_t1; }
                else {
                    int hi = lo;                             // #{Q <= v}; ties live in v's bucket only
                    if (hi < e0 && Qs[hi] == v) {            // first idx in (lo, e0] with Qs > v
                        int y = e0; ++hi;
                        while (hi < y) { const int mid = (hi + y) >> 1; if (Qs[mid] <= v) hi = mid + 1; else y = mid; }
                    }
                    target = hi;
                    a = a && (hi <= jp) && use_t2;
                }
            }
            tgt[u] = target; act[u] = a; nw[u] = -w; nwv[u] = -(w * v);
        }
#pragma unroll
        for (int u = 0; u < LWA_NI; ++u) {                   // own slot j'+1: +(w, w v), no conflicts
            const int jo = jblk + LWA_ROW(lane, u) + 1;
            if (act[u]) { double2 t = Dw[jo]; t.x -= nw[u]; t.y -= nwv[u]; Dw[jo] = t; }
        }
        __syncwarp();
        lwa_scatter<MATCH, HEAVY>(Dw, tagw, tgt, nw, nwv, act, lane);   // far end of the range: -(w, w v)
    }

    // prefix sums down the column: each lane owns a contiguous run of rows (odd
    // length -> conflict-free 128-bit accesses), one warp scan joins the runs;
    // the result overwrites D in place and is written out with coalesced stores.
    {
        const int Lr = ((ny + 31) / 32) | 1;
        const int j0 = lane * Lr, j1 = min(ny, j0 + Lr);
        double aS = 0.0, aV = 0.0;
        for (int j = j0; j < j1; ++j) { double2 d = Dw[j]; aS += d.x; aV += d.y; }
        double xs = aS, xv = aV;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double ts = __shfl_up_sync(XC_FULL, xs, o), tv = __shfl_up_sync(XC_FULL, xv, o);
            if (lane >= o) { xs += ts; xv += tv; }
        }
        double rS = xs - aS, rV = xv - aV;                   // exclusive prefix of this lane's run
        for (int j = j0; j < j1; ++j) {
            double2 d = Dw[j];
            rS += d.x; rV += d.y;
            reinterpret_cast<double*>(Dw + j)[0] = sg * (rV - Qs[j] * rS);
        }
    }
    __syncthreads();
    if (col_ok) {
        const double2* Dc = D + (size_t)cc * L.nyp;
        double* o = out + (s * ny + rr) * (long)nx + i0 + cc;
        for (int j = rr; j < ny; j += 32, o += gstep) *o = Dc[j].x;
    }
}

// NaN-skipping (min, max) of every slice in rngC partials (stand-alone xc_lwa; the
// fused batch already has them from the levels stage) and max |ww|
template <typename QT>
__global__ void k_lwa_range(const QT* __restrict__ q, long P, int C, double* __restrict__ rng)
{
    const long s = blockIdx.y;
    const QT* qs = q + s * P;
    const long per = (P + C - 1) / C, beg = blockIdx.x * per, end = min(P, beg + per);
    double lo = CUDART_INF, hi = -CUDART_INF;
    for (long k = beg + threadIdx.x; k < end; k += blockDim.x) {
        const double v = (double)__ldg(qs + k);
        lo = fmin(lo, v); hi = fmax(hi, v);                      // fmin/fmax skip NaN
    }
    lo = warp_min(lo); hi = warp_max(hi);
    __shared__ double sl[8], sh[8];
    if ((threadIdx.x & 31) == 0) { sl[threadIdx.x >> 5] = lo; sh[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) { lo = fmin(lo, sl[k]); hi = fmax(hi, sh[k]); }
        rng[(s * C + blockIdx.x) * 2] = lo; rng[(s * C + blockIdx.x) * 2 + 1] = hi;
    }
}
__global__ void k_absmax_partial(const double* __restrict__ x, long P, double* __restrict__ part)
{
    double mx = 0.0;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < P; i += (long)gridDim.x * blockDim.x)
        mx = fmax(mx, fabs(x[i]));
    mx = warp_max(mx);
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) mx = fmax(mx, sm[k]);
        part[blockIdx.x] = mx;
    }
}

// Variant 2 (cal_local_wave_activity2, core.py:802-905: the point is fixed, the
// profile varies).  With a sorted profile it needs no scatter at all:
//   out[j] = sg * ( [hi<j] (v (W_j - W_hi) - (QW_j - QW_hi))  -  [lo>j] (v (W_lo - W_j) - (QW_lo - QW_j)) )
// with v = sg*q[j][i], lo = #{Q < v}, hi = #{Q <= v} and the column prefix sums
// W_k = sum_{j'<k} ww[j'][i], QW_k = sum_{j'<k} Q_j' ww[j'][i].  One warp per column
// builds the two prefix arrays in shared memory, then every row is a gather.
constexpr int LWA2_TC = 8;
template <typename QT>
__global__ void __launch_bounds__(LWA2_TC * 32)
k_lwa2_fast(const QT* __restrict__ q, long s0, int ny, int nx,
            const double* __restrict__ Qref, const double* __restrict__ ww,
            int increase, int part, const int32_t* __restrict__ sorted, double* __restrict__ out)
{
    const long s = s0 + blockIdx.y;
    if (!sorted[s]) return;
    extern __shared__ __align__(16) unsigned char smem[];
    double*   Qs  = reinterpret_cast<double*>(smem);                       // [ny]
    double2*  PW  = reinterpret_cast<double2*>(Qs + ((ny + 1) & ~1));      // [TC][ny+1] (W, QW)
    uint32_t* lut = reinterpret_cast<uint32_t*>(PW + (size_t)LWA2_TC * (ny + 1));
    uint16_t* first = reinterpret_cast<uint16_t*>(lut + LWA_LUT);          // [LWA_LUT+1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
    const double sg = increase ? 1.0 : -1.0;
    const QT* qs = q + s * (long)ny * nx;
    const double* Qg = Qref + s * (long)ny;
    for (int j = tid; j < ny; j += nthr) Qs[j] = sg * Qg[j];
    __syncthreads();
    const double qmin = Qs[0], qmax = Qs[ny - 1];
    const float qminf = (float)qmin;
    const float scalef = (qmax > qmin) ? (float)((double)LWA_LUT / (qmax - qmin)) : 0.0f;
    for (int j = tid; j <= ny; j += nthr) {
        const int bj = (j < ny) ? lwa_bucket((float)Qs[j], qminf, scalef) : LWA_LUT;
        const int bp = (j > 0) ? lwa_bucket((float)Qs[j - 1], qminf, scalef) : -1;
        for (int b = bp + 1; b <= bj; ++b) first[b] = (uint16_t)j;
    }
    __syncthreads();
    for (int b = tid; b < LWA_LUT; b += nthr) lut[b] = (uint32_t)first[b] | ((uint32_t)first[b + 1] << 16);

    const int i = blockIdx.x * LWA2_TC + warp;                             // this warp's column
    const bool col_ok = i < nx;
    double2* P = PW + (size_t)warp * (ny + 1);
    // column prefix sums, 32 rows per step (NaN weights contribute nothing: nansum)
    double cW = 0.0, cQ = 0.0;
    for (int r0 = 0; r0 < ny; r0 += 32) {
        const int j = r0 + lane;
        double w = (j < ny && col_ok) ? __ldg(ww + (long)j * nx + i) : 0.0;
        if (w != w) w = 0.0;
        double xw = w, xq = (j < ny) ? sg * Qs[j] * w : 0.0;                // Q_j (original sign) * ww
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double tw = __shfl_up_sync(XC_FULL, xw, o), tq = __shfl_up_sync(XC_FULL, xq, o);
            if (lane >= o) { xw += tw; xq += tq; }
        }
        if (j < ny) P[j + 1] = make_double2(cW + xw, cQ + xq);
        cW += __shfl_sync(XC_FULL, xw, 31); cQ += __shfl_sync(XC_FULL, xq, 31);
    }
    if (lane == 0) P[0] = make_double2(0.0, 0.0);
    __syncthreads();                                                       // lut + prefix arrays ready

    const bool keep_pos = (part == XC_PART_UPPER) == (increase != 0);
    const bool use_t1 = (part == XC_PART_ALL) || !keep_pos;   // rows j' <  j (mask -1 region)
    const bool use_t2 = (part == XC_PART_ALL) || keep_pos;    // rows j' >= j (mask +1 region)
    if (!col_ok) return;
    for (int j = lane; j < ny; j += 32) {
        const double qv = (double)__ldg(qs + (long)j * nx + i);
        double res = 0.0;
        if (qv == qv) {
            const double v = sg * qv;
            const uint32_t pk = lut[lwa_bucket((float)v, qminf, scalef)];
            int x = (int)(pk & 0xffffu), e = (int)(pk >> 16);
            const int e0 = e;
            while (x < e) { const int mid = (x + e) >> 1; if (Qs[mid] < v) x = mid + 1; else e = mid; }
            const int lo = x;
            int hi = lo;
            if (hi < e0 && Qs[hi] == v) {
                int y = e0; ++hi;
                while (hi < y) { const int mid = (hi + y) >> 1; if (Qs[mid] <= v) hi = mid + 1; else y = mid; }
            }
            const double2 pj = P[j];
            // sums use the original-sign profile: sum (q - Q_j') ww = qv*dW - dQW
            if (use_t1 && hi < j) { const double2 a = P[hi]; res += qv * (pj.x - a.x) - (pj.y - a.y); }
            if (use_t2 && lo > j) { const double2 b = P[lo]; res -= qv * (b.x - pj.x) - (b.y - pj.y); }
        }
        out[(s * ny + j) * (long)nx + i] = res;
    }
}

// Exact reference loop (any profile, both variants).  Persistent grid: each CTA
// walks (slice, tile) work items; slices already handled by k_lwa_fast are
// skipped.  block = (32, 8): 32 columns x 8 output rows.
template <typename QT>
__global__ void __launch_bounds__(256)
k_lwa_brute(const QT* __restrict__ q, long S, int ny, int nx,
            const double* __restrict__ Qref, const double* __restrict__ ww,
            int increase, int part, int variant, const int32_t* __restrict__ skip,
            const int32_t* __restrict__ gate, double* __restrict__ out, int out_f32)
{
    if (gate && *gate == 0) return;            // every slice went down the fast path
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tiles_x = (nx + 31) / 32, tiles_y = (ny + 7) / 8;
    const long tiles = (long)tiles_x * tiles_y;
    const bool f = variant == 1 ? (increase != 0) : (increase == 0);   // core.py:759 / 865
    const bool keep_pos = (part == XC_PART_UPPER) == (increase != 0);
    for (long s = 0; s < S; ++s) {
        if (skip && skip[s]) continue;
        const QT* qs = q + s * (long)ny * nx;
        const double* Qs = Qref + s * (long)ny;
        for (long t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int i = (int)(t % tiles_x) * 32 + tx;
            const int j = (int)(t / tiles_x) * 8 + ty;
            if (i >= nx || j >= ny) continue;
            const double Qj = Qs[j];
            const double qj = (double)__ldg(qs + (long)j * nx + i);
            double acc = 0.0;
            for (int jp = 0; jp < ny; ++jp) {
                const double qe = variant == 1
                    ? (double)__ldg(qs + (long)jp * nx + i) - Qj
                    : qj - Qs[jp];
                const bool m = jp >= j;
                int mask = 0;
                if (f) { if (qe > 0.0 && !m) mask = -1; else if (qe < 0.0 && m) mask = 1; }
                else   { if (qe < 0.0 && !m) mask = -1; else if (qe > 0.0 && m) mask = 1; }
                if (part != XC_PART_ALL && ((mask > 0) != keep_pos)) mask = 0;
                if (mask != 0) {
                    double term = qe * (double)mask * __ldg(ww + (long)jp * nx + i);
                    if (!isnan(term)) acc += term;
                }
            }
            if (out_f32) reinterpret_cast<float*>(out)[(s * ny + j) * (long)nx + i] = (float)-acc;
            else out[(s * ny + j) * (long)nx + i] = -acc;
        }
    }
}

template <typename QT>
__global__ void k_lwa_mask(const QT* __restrict__ q, long total, int ny, int nx,
                           const double* __restrict__ Qref, int j, int increase,
                           int variant, int8_t* __restrict__ mask)
{
    const bool f = variant == 1 ? (increase != 0) : (increase == 0);
    const long plane = (long)ny * nx;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const long s = idx / plane; const long r = idx - s * plane;
        const int jp = (int)(r / nx); const int i = (int)(r - (long)jp * nx);
        const double* Qs = Qref + s * (long)ny;
        const double qe = variant == 1 ? (double)q[idx] - Qs[j]
                                       : (double)q[s * plane + (long)j * nx + i] - Qs[jp];
        const bool m = jp >= j;
        int8_t mk = 0;
        if (f) { if (qe > 0.0 && !m) mk = -1; else if (qe < 0.0 && m) mk = 1; }
        else   { if (qe < 0.0 && !m) mk = -1; else if (qe > 0.0 && m) mk = 1; }
        mask[idx] = mk;
    }
}

// ---- ww = (dA/max(dA)) * dA ------------------------------------------------
__global__ void k_max_partial(const void* __restrict__ dA, int is_f32, long P, double* __restrict__ part)
{
    double mx = -CUDART_INF;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < P; i += (long)gridDim.x * blockDim.x)
        mx = fmax(mx, ld_as_f64(dA, i, is_f32));
    mx = warp_max(mx);
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) mx = fmax(mx, sm[k]);
        part[blockIdx.x] = mx;
    }
}
__global__ void k_lwa_weights(const void* __restrict__ dA, int is_f32, long P,
                              const double* __restrict__ part, int nparts, double* __restrict__ ww)
{
    double mx = -CUDART_INF;
    for (int k = 0; k < nparts; ++k) mx = fmax(mx, part[k]);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < P; i += (long)gridDim.x * blockDim.x) {
        if (is_f32) {
            float a = ((const float*)dA)[i];
            float wei = __fdiv_rn(a, (float)mx);             // rounded in dA's dtype
            ww[i] = __dmul_rn((double)wei, (double)a);
        } else {
            double a = ((const double*)dA)[i];
            ww[i] = __dmul_rn(__ddiv_rn(a, mx), a);
        }
    }
}

}  // namespace xc

using namespace xc;

static bool lwa_use_match()
{
    static int v = -1;
    // measured on B200 (scripts/ab_variants.sh): the byte-tag protocol is ~10 % faster here
    if (v < 0) { const char* e = getenv("XCB200_LWA_DEDUP"); v = (e && e[0] == 'm') ? 1 : 0; }
    return v == 1;
}

static int lwa_pick_tc(int ny, int qbytes, bool tags)
{
    for (int tc = LWA_MAX_TC; tc >= 1; --tc)
        if (lwa_layout(ny, tc, qbytes, tags).total <= 227 * 1024) return tc;
    return 0;
}

// XCB200_LWA_HEAVY=n (n > 0): after n tag-election rounds the lanes that still
// share a slot are combined in registers (robust against homogenised regions where
// dozens of cells map to one row: 1.5x faster there, ~8 % slower on the ERA5-like
// benchmark field, hence off by default).
static int lwa_heavy_rounds()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("XCB200_LWA_HEAVY"); v = e ? atoi(e) : 0; if (v < 0) v = 0; }
    return v;
}

template <typename QT, bool MATCH, int HEAVY>
static int launch_lwa_fast_h(const QT* q, long S, int n_eq, int n_x, const double* Qref, const double* ww,
                           int increase, int part, const int32_t* sorted, double* out, int tc, cudaStream_t st)
{
    const LwaSmem L = lwa_layout(n_eq, tc, (int)sizeof(QT), !MATCH);
    XC_CUDA_OK(cudaFuncSetAttribute(k_lwa_fast<QT, MATCH, HEAVY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    for (long s0 = 0; s0 < S; s0 += 65535) {
        long ns = S - s0 < 65535 ? S - s0 : 65535;
        dim3 grid((unsigned)((n_x + tc - 1) / tc), (unsigned)ns);
        k_lwa_fast<QT, MATCH, HEAVY><<<grid, tc * 32, L.total, st>>>(q, s0, n_eq, n_x, Qref, ww, increase, part, sorted, out, tc);
        XC_LAUNCH_OK();
    }
    return 0;
}

template <typename QT, bool MATCH>
static int launch_lwa_fast(const QT* q, long S, int n_eq, int n_x, const double* Qref, const double* ww,
                           int increase, int part, const int32_t* sorted, double* out, int tc, cudaStream_t st)
{
    return lwa_heavy_rounds() > 0
        ? launch_lwa_fast_h<QT, MATCH, 3>(q, S, n_eq, n_x, Qref, ww, increase, part, sorted, out, tc, st)
        : launch_lwa_fast_h<QT, MATCH, 0>(q, S, n_eq, n_x, Qref, ww, increase, part, sorted, out, tc, st);
}

extern "C" size_t xc_lwa_weights_workspace_bytes(long P) { (void)P; return 256 + 256 * sizeof(double); }

extern "C" int xc_lwa_weights(const void* dA, int dA_dtype, long P, double* ww,
                              void* workspace, size_t ws_bytes, void* stream)
{
    XC_REQUIRE(dA && ww && P > 0, "xc_lwa_weights: bad arguments");
    XC_REQUIRE(workspace && ws_bytes >= xc_lwa_weights_workspace_bytes(P), "xc_lwa_weights: workspace too small");
    Arena ar(workspace, ws_bytes);
    double* part = ar.take<double>(256);
    cudaStream_t st = (cudaStream_t)stream;
    int nb = (int)((P + 255) / 256); if (nb > 256) nb = 256;
    k_max_partial<<<nb, 256, 0, st>>>(dA, dA_dtype == XC_F32, P, part);
    XC_LAUNCH_OK();
    k_lwa_weights<<<nb, 256, 0, st>>>(dA, dA_dtype == XC_F32, P, part, nb, ww);
    XC_LAUNCH_OK();
    return 0;
}

constexpr int LWA_RNG_C = 8;        // partial (min, max) CTAs per slice in the stand-alone path
constexpr int LWA_WMAX_N = 128;     // partial max |ww| CTAs
size_t xc::lwa_scratch_doubles(long S, bool have_minmax)
{
    const long Sp = S > 0 ? S : 0, ch = Sp < FX_CHUNK ? Sp : FX_CHUNK;
    return (size_t)LWA_WMAX_N + 32 + (have_minmax ? 0 : (size_t)Sp * LWA_RNG_C * 2) +
           (size_t)ch * (sizeof(FxScale) / 8 + FX_LUT / 2) + 64;
}

extern "C" size_t xc_lwa_workspace_bytes(long S)
{
    return 1024 + (size_t)(S > 0 ? S : 0) * sizeof(int32_t) + lwa_scratch_doubles(S, false) * sizeof(double);
}

extern "C" int xc_lwa_ex(const void* q, int q_dtype, long S, int n_eq, int n_x,
                         const double* Qref, const double* ww, const double* ww_row,
                         int increase, int part, int variant,
                         double* out, void* workspace, size_t ws_bytes, void* stream)
{
    XC_REQUIRE(workspace && ws_bytes >= xc_lwa_workspace_bytes(S), "xc_lwa: workspace too small");
    Arena ar(workspace, ws_bytes);
    int32_t* sorted = ar.take<int32_t>((size_t)(S > 0 ? S : 0));
    double* scratch = ar.take<double>(lwa_scratch_doubles(S, false));
    return lwa_impl(q, q_dtype, S, n_eq, n_x, Qref, ww, increase, part, variant, out, sorted,
                    nullptr, false, nullptr, scratch, stream, nullptr, ww_row);
}

extern "C" int xc_lwa(const void* q, int q_dtype, long S, int n_eq, int n_x,
                      const double* Qref, const double* ww,
                      int increase, int part, int variant,
                      double* out, void* workspace, size_t ws_bytes, void* stream)
{
    return xc_lwa_ex(q, q_dtype, S, n_eq, n_x, Qref, ww, nullptr, increase, part, variant, out, workspace, ws_bytes, stream);
}

// XCB200_LWA_FX=0 selects the fp64 read-modify-write kernel (k_lwa_fast) instead of
// the fixed-point one (kept for A/B timing and as a cross-check in the tests).
static bool lwa_use_fx()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("XCB200_LWA_FX"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

size_t xc::lwa_wmax_doubles() { return LWA_WMAX_N; }
int xc::lwa_wmax(const double* ww, long P, double* parts, void* stream)
{
    k_absmax_partial<<<LWA_WMAX_N, 256, 0, (cudaStream_t)stream>>>(ww, P, parts);
    XC_LAUNCH_OK();
    return 0;
}

int xc::lwa_impl(const void* q, int q_dtype, long S, int n_eq, int n_x, const double* Qref, const double* ww,
                 int increase, int part, int variant, double* out, int32_t* sorted,
                 int32_t* any_unsorted, bool flags_ready, const double* minmax, double* scratch, void* stream,
                 const double* wmax_ready, const double* ww_row, int out_f32)
{
    XC_REQUIRE(q && Qref && ww && out, "xc_lwa: null pointer");
    XC_REQUIRE(S > 0 && n_eq >= 1 && n_x >= 1, "xc_lwa: need S>0, n_eq>=1, n_x>=1");
    XC_REQUIRE(q_dtype == XC_F32 || q_dtype == XC_F64, "xc_lwa: bad dtype");
    XC_REQUIRE(part >= XC_PART_ALL && part <= XC_PART_LOWER,
               "invalid part, should be in ['all', 'upper', 'lower']");
    XC_REQUIRE(variant == 1 || variant == 2, "xc_lwa: variant must be 1 or 2");
    cudaStream_t st = (cudaStream_t)stream;
    const bool match = lwa_use_match();
    const int qbytes = q_dtype == XC_F32 ? 4 : 8;
    const int tc = (n_eq < 65535) ? lwa_pick_tc(n_eq, qbytes, !match) : 0;
    static const char* no_cols0 = getenv("XCB200_NO_LWA_COLS");
    const bool fx = (variant == 1) && lwa_use_fx() && scratch && n_eq < 65535 &&
                    (lwa_fx_fits(n_eq) || (ww_row && !no_cols0 && lwa_cols_fits(n_eq, qbytes)));
    XC_REQUIRE(!out_f32 || fx, "xc_keff_lwa_batch: lwa_f32 needs the fixed-point column-tile kernel");
    const bool fast = fx || ((variant == 1) && tc >= 1 && (size_t)tc * (n_eq + 2) * 16 >= (size_t)(LWA_LUT + 1) * 2);
    if (fast && !flags_ready) {
        k_check_sorted<<<(unsigned)S, 256, 0, st>>>(Qref, n_eq, increase, sorted);
        XC_LAUNCH_OK();
    }
    if (fx) {
        const long P = (long)n_eq * n_x;
        double* wpart = scratch;
        const double* rng = minmax; int rngC = 1;
        if (!rng) {
            double* r = scratch + LWA_WMAX_N + 32; rng = r; rngC = LWA_RNG_C;
            for (long s0 = 0; s0 < S; s0 += 65535) {
                const long ns = S - s0 < 65535 ? S - s0 : 65535;
                dim3 grid(LWA_RNG_C, (unsigned)ns);
                if (q_dtype == XC_F32) k_lwa_range<float><<<grid, 256, 0, st>>>((const float*)q + s0 * P, P, LWA_RNG_C, r + s0 * LWA_RNG_C * 2);
                else                   k_lwa_range<double><<<grid, 256, 0, st>>>((const double*)q + s0 * P, P, LWA_RNG_C, r + s0 * LWA_RNG_C * 2);
                XC_LAUNCH_OK();
            }
        }
        const double* wparts = wmax_ready;
        if (!wparts) {
            k_absmax_partial<<<LWA_WMAX_N, 256, 0, st>>>(ww, P, wpart);
            XC_LAUNCH_OK();
            wparts = wpart;
        }
        char* cur = reinterpret_cast<char*>(scratch + LWA_WMAX_N + 32 + (minmax ? 0 : (size_t)S * LWA_RNG_C * 2));
        cur = reinterpret_cast<char*>(align_up((size_t)cur, 64));
        FxScale* fxs = reinterpret_cast<FxScale*>(cur);
        const long ch = S < FX_CHUNK ? S : FX_CHUNK;
        uint32_t* lutg = reinterpret_cast<uint32_t*>(cur + (size_t)ch * sizeof(FxScale));
        // row-constant weights: the column-tile kernel with register-resident own deposits (lwa_cols.cu);
        // general weights: k_lwa_fx (lwa_fx.cu)
        static const char* no_cols = getenv("XCB200_NO_LWA_COLS");
        const bool cols = ww_row && !no_cols && lwa_cols_fits(n_eq, qbytes);
        XC_REQUIRE(cols || !out_f32, "xc_keff_lwa_batch: lwa_f32 needs the column-tile kernel (row-constant weights, n_y <= 768)");
        for (long s0 = 0; s0 < S; s0 += FX_CHUNK) {
            const long ns = S - s0 < FX_CHUNK ? S - s0 : FX_CHUNK;
            if (lwa_fx_prep_launch(s0, ns, n_eq, Qref, increase, sorted, any_unsorted, rng, rngC, wparts, LWA_WMAX_N, fxs, lutg, stream)) return 1;
            const int rc = cols ? lwa_cols_launch(q, q_dtype, s0, ns, n_eq, n_x, Qref, ww_row, increase, part, sorted, fxs, lutg, out, out_f32, stream)
                                : lwa_fx_launch(q, q_dtype, s0, ns, n_eq, n_x, Qref, ww, increase, part, sorted, fxs, lutg, out, stream);
            if (rc) return rc;
        }
    } else if (fast) {
        int rc;
        if (q_dtype == XC_F32)
            rc = match ? launch_lwa_fast<float, true>((const float*)q, S, n_eq, n_x, Qref, ww, increase, part, sorted, out, tc, st)
                       : launch_lwa_fast<float, false>((const float*)q, S, n_eq, n_x, Qref, ww, increase, part, sorted, out, tc, st);
        else
            rc = match ? launch_lwa_fast<double, true>((const double*)q, S, n_eq, n_x, Qref, ww, increase, part, sorted, out, tc, st)
                       : launch_lwa_fast<double, false>((const double*)q, S, n_eq, n_x, Qref, ww, increase, part, sorted, out, tc, st);
        if (rc) return rc;
    }
    const size_t sm2 = (size_t)((n_eq + 1) & ~1) * 8 + (size_t)LWA2_TC * (n_eq + 1) * 16 +
                       (size_t)LWA_LUT * 4 + (size_t)(LWA_LUT + 2) * 2;
    const bool fast2 = (variant == 2) && n_eq < 65535 && sm2 <= 227 * 1024;
    if (fast2) {
        k_check_sorted<<<(unsigned)S, 256, 0, st>>>(Qref, n_eq, increase, sorted);
        XC_LAUNCH_OK();
        if (q_dtype == XC_F32) XC_CUDA_OK(cudaFuncSetAttribute(k_lwa2_fast<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        else                   XC_CUDA_OK(cudaFuncSetAttribute(k_lwa2_fast<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        for (long s0 = 0; s0 < S; s0 += 65535) {
            const long ns = S - s0 < 65535 ? S - s0 : 65535;
            dim3 grid((unsigned)((n_x + LWA2_TC - 1) / LWA2_TC), (unsigned)ns);
            if (q_dtype == XC_F32)
                k_lwa2_fast<float><<<grid, LWA2_TC * 32, sm2, st>>>((const float*)q, s0, n_eq, n_x, Qref, ww, increase, part, sorted, out);
            else
                k_lwa2_fast<double><<<grid, LWA2_TC * 32, sm2, st>>>((const double*)q, s0, n_eq, n_x, Qref, ww, increase, part, sorted, out);
            XC_LAUNCH_OK();
        }
    }
    // exact loop for whatever the fast paths did not take (exits at once when the
    // fused epilogue reported that every profile is sorted)
    dim3 blk(32, 8);
    const bool gated = fast && flags_ready && any_unsorted;
    unsigned nb = (unsigned)(sm_count() * (gated ? 1 : 8));
    if (q_dtype == XC_F32)
        k_lwa_brute<float><<<nb, blk, 0, st>>>((const float*)q, S, n_eq, n_x, Qref, ww, increase, part,
                                               variant, (fast || fast2) ? sorted : nullptr, gated ? any_unsorted : nullptr, out, out_f32);
    else
        k_lwa_brute<double><<<nb, blk, 0, st>>>((const double*)q, S, n_eq, n_x, Qref, ww, increase, part,
                                                variant, (fast || fast2) ? sorted : nullptr, gated ? any_unsorted : nullptr, out, out_f32);
    XC_LAUNCH_OK();
    return 0;
}

extern "C" int xc_lwa_mask(const void* q, int q_dtype, long S, int n_eq, int n_x,
                           const double* Qref, int j, int increase, int variant,
                           int8_t* mask, void* stream)
{
    XC_REQUIRE(q && Qref && mask, "xc_lwa_mask: null pointer");
    XC_REQUIRE(j >= 0 && j < n_eq, "indices in mask_idx out of boundary");
    const long total = S * (long)n_eq * n_x;
    long b = (total + 255) / 256; long cap = (long)sm_count() * 16; if (b > cap) b = cap;
    if (q_dtype == XC_F32)
        k_lwa_mask<float><<<(unsigned)b, 256, 0, (cudaStream_t)stream>>>((const float*)q, total, n_eq, n_x, Qref, j, increase, variant, mask);
    else
        k_lwa_mask<double><<<(unsigned)b, 256, 0, (cudaStream_t)stream>>>((const double*)q, total, n_eq, n_x, Qref, j, increase, variant, mask);
    XC_LAUNCH_OK();
    return 0;
}
