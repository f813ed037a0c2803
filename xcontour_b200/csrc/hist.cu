// Kernel (2): conditional accumulation.  Bins every cell of a slice against the
// slice's ascending fp64 edges (an exact np.digitize: arithmetic guess, then a
// fix-up against the true edge values held in shared memory), accumulates K
// weighted sums per bin in fp64 in warp-private shared-memory histograms, and a
// second small kernel reduces the per-CTA partials in a fixed order and runs a
// block-wide scan over bins.  Replaces xhistogram + cumsum
// (xcontour/core.py:412-460, 1202-1325) and, with closed_right / SUFFIX, the 4-D
// broadcast of core.py:363-409.
//
// Why warp-private, non-atomic histograms: sm_100a has no native shared-memory
// fp64 (or 64-bit integer) add -- atomicAdd(double*) on shared memory compiles
// to an ATOMS.CAST.SPIN loop -- so each warp owns a copy and resolves the lanes
// that hit the same bin itself: one MATCH.ANY gives every lane the mask of its
// peers and the lowest remaining peer of each group does a plain 128-bit
// read-modify-write per round (default); alternatively every pending lane
// stores its id into a byte tag of the bin and the lane that reads its own id
// back owns the bin for the round (XCB200_HIST_DEDUP=t).  Bins too many for
// private copies (e.g. N = 2048 with two accumulators) fall back to shared copies
// with CAS atomics.  The fused Keff pass has a specialised kernel in bin_rows.cu.
#include "common.cuh"
#include "internal.h"
#include "grad2.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace xc {

constexpr int HIST_MAX_WARPS = 16;

struct HistParams {
    const void* q; long P; long per; long s0;
    const double* edges; long edges_stride; int N; int closed_right;
    const void* dA; int dA_f32; int acc_area;
    const void* integ[XC_MAX_INTEGRANDS]; int integ_f32[XC_MAX_INTEGRANDS]; int n_int;
    const uint8_t* q_mask;
    // in-flight |grad q|^2 integrand (last accumulator slot) -- see grad2.cuh
    int stencil; int ny, nx; const double* cx; const double* cy;
    double* part;            // [S][C][K][N]
    int32_t* bin_idx;        // [S][P] or null
    int ncopy;
};

struct EdgeGuess { double base, inv; float basef, invf; int off; int uniform; };

// Exact digitize against ascending e[0..N]; -1 when the cell is discarded.
// Uniform edges: an fp32 guess of the bin, then the guess is walked to the bin
// whose true fp64 edges bracket v (usually zero or one step) -- the comparisons
// against e[] decide, the arithmetic never does.
__device__ __forceinline__ int find_bin(double v, float vf, const double* e, int N,
                                        const EdgeGuess& g, int closed_right)
{
    if (vf != vf) return -1;
    int p;
    if (g.uniform) {
        const float t = (vf - g.basef) * g.invf;
        p = g.off + __float2int_rd(fminf(fmaxf(t, -1.0f), (float)N));
        p = min(max(p, 0), N - 1);
        for (;;) {
            const double lo = e[p], hi = e[p + 1];
            const int d = closed_right ? ((v > hi) - (v <= lo)) : ((v >= hi) - (v < lo));
            if (d == 0) return p;
            p += d;
            if (p < 0 || p >= N) return -1;
        }
    }
    if (closed_right) { if (!(v > e[0]) || !(v <= e[N])) return -1; }
    else              { if (!(v >= e[0]) || !(v < e[N])) return -1; }
    int lo = 0, hi = N;                // first index with e[idx] > v  (or >= v)
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        bool right = closed_right ? (e[mid] < v) : (e[mid] <= v);
        if (right) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

template <int K>
__device__ __forceinline__ void rmw_add(double* h, const double (&w)[K])
{
    if (K == 2) {
        double2 t = *reinterpret_cast<double2*>(h);
        t.x += w[0]; t.y += w[1];
        *reinterpret_cast<double2*>(h) = t;
    } else if (K == 4) {
        double2 a = reinterpret_cast<double2*>(h)[0], b = reinterpret_cast<double2*>(h)[1];
        a.x += w[0]; a.y += w[1]; b.x += w[2]; b.y += w[3];
        reinterpret_cast<double2*>(h)[0] = a; reinterpret_cast<double2*>(h)[1] = b;
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) h[k] += w[k];
    }
}

enum { HIST_TAG = 0, HIST_MATCH = 1, HIST_ATOMIC = 2 };

// Warp-collective scatter-add of 4 items per lane into a warp-private histogram
// (all 32 lanes must call).  Lanes that hit the same bin are serialised:
//   HIST_MATCH: four independent MATCH.ANY give each item the mask of its peers;
//               per round the lowest remaining peer of every group does a plain
//               128-bit read-modify-write, all members then clear that bit.
//   HIST_TAG  : byte-tag protocol (write lane id, winner reads its own id back).
template <int K, int MODE>
__device__ __forceinline__ void scatter_add4(double* H, uint8_t* tag, const int (&bin)[4],
                                             const double (&w)[4][K], int lane)
{
    if (MODE == HIST_MATCH) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool a = bin[u] >= 0;
            unsigned pr = __match_any_sync(XC_FULL, a ? (unsigned)bin[u] : (0x80000000u | (unsigned)lane));
            if (!a) pr = 0u;
            if (__reduce_max_sync(XC_FULL, __popc(pr)) <= 3) {
                // few duplicates: the lowest remaining peer of every group adds, all clear that bit
                do {
                    if (pr && (__ffs(pr) - 1) == lane) rmw_add<K>(H + (size_t)bin[u] * K, w[u]);
                    pr &= pr - 1u;
                    __syncwarp();
                } while (__any_sync(XC_FULL, pr != 0u));
            } else {
                // many lanes share a bin (smooth fields): combine them in registers first.
                // The peers of a bin form a linked list in lane order; pointer jumping
                // leaves the group total in the lowest lane after log2(group size) steps.
                const unsigned above = (lane == 31) ? 0u : (pr & (0xffffffffu << (lane + 1)));
                int nxt = above ? (__ffs(above) - 1) : -1;
                const bool leader = a && ((__ffs(pr) - 1) == lane);
                double acc[K];
#pragma unroll
                for (int k = 0; k < K; ++k) acc[k] = w[u][k];
                while (__any_sync(XC_FULL, nxt >= 0)) {
                    const int src = nxt >= 0 ? nxt : lane;
                    double g[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) g[k] = __shfl_sync(XC_FULL, acc[k], src);
                    const int gn = __shfl_sync(XC_FULL, nxt, src);
                    if (nxt >= 0) {
#pragma unroll
                        for (int k = 0; k < K; ++k) acc[k] += g[k];
                        nxt = gn;
                    }
                }
                if (leader) rmw_add<K>(H + (size_t)bin[u] * K, acc);
                __syncwarp();
            }
        }
    } else if (MODE == HIST_TAG) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            bool active = bin[u] >= 0;
            unsigned pending = __ballot_sync(XC_FULL, active);
            while (pending) {
                if (active) tag_store(tag + bin[u], (unsigned)lane);
                __syncwarp();
                if (active && tag_load(tag + bin[u]) == (unsigned)lane) {
                    rmw_add<K>(H + (size_t)bin[u] * K, w[u]);
                    active = false;
                }
                __syncwarp();
                pending = __ballot_sync(XC_FULL, active);
            }
        }
    } else {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (bin[u] >= 0) {
#pragma unroll
                for (int k = 0; k < K; ++k) atomicAdd(H + (size_t)bin[u] * K + k, w[u][k]);
            }
    }
}

// w[k] = x with k only known at run time, without spilling w to local memory
template <int K>
__device__ __forceinline__ void put(double (&w)[K], int& k, double x)
{
#pragma unroll
    for (int kk = 0; kk < K; ++kk) if (kk == k) w[kk] = x;
    ++k;
}

// grid = (C, nslices).  MODE != HIST_ATOMIC: one histogram copy per warp, no atomics.
template <typename QT, int K, int MODE>
__global__ void __launch_bounds__(HIST_MAX_WARPS * 32)
k_hist(const HistParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int N = p.N;
    double* e = reinterpret_cast<double*>(smem);
    const int eN = (N + 2) & ~1;
    double* H = e + eN;
    uint8_t* tags = reinterpret_cast<uint8_t*>(H + (size_t)p.ncopy * N * K);
    const int tagN = (N + 15) & ~15;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const long s = p.s0 + blockIdx.y;
    const int  c = blockIdx.x, C = gridDim.x;

    const double* eg = p.edges + s * p.edges_stride;
    for (int k = tid; k <= N; k += blockDim.x) e[k] = eg[k];
    for (int i = tid; i < p.ncopy * N * K; i += blockDim.x) H[i] = 0.0;
    __syncthreads();

    EdgeGuess g;
    {
        int a = isinf(e[0]) ? 1 : 0, b = isinf(e[N]) ? N - 1 : N;
        double span = e[b] - e[a];
        g.base = e[a]; g.off = a;
        g.inv = (b > a && span > 0.0) ? (double)(b - a) / span : 0.0;
        int bad = (g.inv == 0.0) || !isfinite(g.inv);
        double h = (b > a) ? span / (double)(b - a) : 0.0;
        for (int k = a + tid; k <= b; k += blockDim.x)
            if (fabs(e[k] - (g.base + (double)(k - a) * h)) > h) bad = 1;
        g.uniform = !__syncthreads_or(bad);
        g.basef = (float)g.base; g.invf = (float)g.inv;
        if (!isfinite(g.invf) || !isfinite(g.basef)) g.uniform = 0;
    }

    const QT* qs = reinterpret_cast<const QT*>(p.q) + s * p.P;
    const long beg = (long)c * p.per;
    const long end = beg + p.per < p.P ? beg + p.per : p.P;
    double*  Hw   = H + (size_t)(warp % p.ncopy) * N * K;
    uint8_t* tagw = tags + (size_t)warp * tagN;
    const bool vec_ok = ((p.P & 3) == 0) && ((((uintptr_t)p.q) & 15) == 0);
    // 4 consecutive cells share a row when nx % 4 == 0: the stencil then needs one
    // 128-bit load from the row above, one from the row below and two scalars
    const bool st_vec = p.stencil && vec_ok && ((p.nx & 3) == 0) && sizeof(QT) == 4;

    for (long base = beg + (long)warp * 128; base < end; base += (long)nwarps * 128) {
        const long i0 = base + lane * 4;
        const bool full = vec_ok && i0 + 3 < end;
        QT qv[4];
        if (full) {
            if (sizeof(QT) == 4) {
                float4 t = __ldg(reinterpret_cast<const float4*>(qs + i0));
                qv[0] = t.x; qv[1] = t.y; qv[2] = t.z; qv[3] = t.w;
            } else {
                double2 a = __ldg(reinterpret_cast<const double2*>(qs + i0));
                double2 b = __ldg(reinterpret_cast<const double2*>(qs + i0) + 1);
                qv[0] = a.x; qv[1] = a.y; qv[2] = b.x; qv[3] = b.y;
            }
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) qv[u] = (i0 + u < end) ? __ldg(qs + i0 + u) : (QT)CUDART_NAN;
        }
        // weights of the 4 cells
        double ad[4]; float af[4];
        if (full && p.dA_f32) {
            float4 t = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.dA) + i0));
            af[0] = t.x; af[1] = t.y; af[2] = t.z; af[3] = t.w;
#pragma unroll
            for (int u = 0; u < 4; ++u) ad[u] = (double)af[u];
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                af[u] = 0.f; ad[u] = 0.0;
                if (i0 + u < end) {
                    if (p.dA_f32) { af[u] = __ldg(reinterpret_cast<const float*>(p.dA) + i0 + u); ad[u] = (double)af[u]; }
                    else          { ad[u] = __ldg(reinterpret_cast<const double*>(p.dA) + i0 + u); }
                }
            }
        }
        double gq[4] = {0.0, 0.0, 0.0, 0.0};
        if (p.stencil && i0 < end) {
            const int j = (int)((unsigned long long)i0 / (unsigned)p.nx);
            const int col = (int)(i0 - (long)j * p.nx);
            if (st_vec && full) {
                const int jm = j == 0 ? 0 : j - 1, jp = j == p.ny - 1 ? p.ny - 1 : j + 1;
                const float4 nn = __ldg(reinterpret_cast<const float4*>(qs + (long)jp * p.nx + col));
                const float4 ss = __ldg(reinterpret_cast<const float4*>(qs + (long)jm * p.nx + col));
                const float wv = __ldg(reinterpret_cast<const float*>(qs) + (long)j * p.nx + (col == 0 ? p.nx - 1 : col - 1));
                const float ev = __ldg(reinterpret_cast<const float*>(qs) + (long)j * p.nx + (col + 4 == p.nx ? 0 : col + 4));
                const double cx = __ldg(p.cx + j), cy = __ldg(p.cy + j);
                const double c0 = (double)qv[0], c1 = (double)qv[1], c2 = (double)qv[2], c3 = (double)qv[3];
                gq[0] = grad2_from(c1, (double)wv, (double)nn.x, (double)ss.x, cx, cy);
                gq[1] = grad2_from(c2, c0, (double)nn.y, (double)ss.y, cx, cy);
                gq[2] = grad2_from(c3, c1, (double)nn.z, (double)ss.z, cx, cy);
                gq[3] = grad2_from((double)ev, c2, (double)nn.w, (double)ss.w, cx, cy);
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (i0 + u < end) {
                        int jj = j, cu = col + u;
                        while (cu >= p.nx) { cu -= p.nx; ++jj; }
                        gq[u] = grad2_cell(qs, jj, cu, p.ny, p.nx, __ldg(p.cx + jj), __ldg(p.cy + jj));
                    }
                }
            }
        }
        int bins[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long i = i0 + u;
            bool ok = i < end;
            if (ok && p.q_mask) ok = p.q_mask[i] != 0;
            bins[u] = ok ? find_bin((double)qv[u], (float)qv[u], e, N, g, p.closed_right) : -1;
        }
        if (p.bin_idx) {
            int32_t* bo = p.bin_idx + s * p.P;
#pragma unroll
            for (int u = 0; u < 4; ++u) if (i0 + u < end) bo[i0 + u] = bins[u];
        }
        double w[4][K];
        const bool fast_layout = (K == 2) && p.acc_area && p.n_int == 0 && p.stencil;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long i = i0 + u;
            const bool act = bins[u] >= 0;
#pragma unroll
            for (int k = 0; k < K; ++k) w[u][k] = 0.0;
            if (fast_layout) {                       // {dA, |grad q|^2 dA}: the fused Keff pass
                const double pr = __dmul_rn(gq[u], ad[u]);
                w[u][0] = isnan(ad[u]) ? 0.0 : ad[u];
                w[u][K - 1] = isnan(pr) ? 0.0 : pr;
            } else if (act) {
                int k = 0;
                if (p.acc_area) put<K>(w[u], k, isnan(ad[u]) ? 0.0 : ad[u]);
#pragma unroll
                for (int n = 0; n < XC_MAX_INTEGRANDS; ++n) {
                    if (n < p.n_int) {
                        double pr;
                        if (p.integ_f32[n]) {
                            float gf = __ldg(reinterpret_cast<const float*>(p.integ[n]) + s * p.P + i);
                            // product rounded in the common dtype (core.py:444)
                            pr = p.dA_f32 ? (double)__fmul_rn(gf, af[u]) : __dmul_rn((double)gf, ad[u]);
                        } else {
                            double gd = __ldg(reinterpret_cast<const double*>(p.integ[n]) + s * p.P + i);
                            pr = __dmul_rn(gd, ad[u]);
                        }
                        put<K>(w[u], k, isnan(pr) ? 0.0 : pr);     // fillna(0), core.py:449
                    }
                }
                if (p.stencil) {
                    const double pr = __dmul_rn(gq[u], ad[u]);
                    put<K>(w[u], k, isnan(pr) ? 0.0 : pr);
                }
            }
        }
        scatter_add4<K, MODE>(Hw, tagw, bins, w, lane);
    }
    __syncthreads();
    double* out = p.part + ((size_t)(blockIdx.y + p.s0) * C + c) * K * N;
    for (int idx = tid; idx < K * N; idx += blockDim.x) {
        const int k = idx / N, n = idx - k * N;
        double acc = 0.0;
        for (int cp = 0; cp < p.ncopy; ++cp) acc += H[((size_t)cp * N + n) * K + k];
        out[idx] = acc;
    }
}

// grid = (S, K).  Fixed-order reduction of the C partials, then a block scan.
__global__ void __launch_bounds__(256)
k_reduce_scan(const double* __restrict__ part, int C, int K, int N, int scan_mode,
              const int32_t* __restrict__ decreasing,
              double* __restrict__ pdf, const ScanOut so)
{
    extern __shared__ __align__(16) unsigned char smem[];
    double* pd = reinterpret_cast<double*>(smem);       // N values, scan order
    __shared__ double total_s;
    const long s = blockIdx.x; const int k = blockIdx.y;
    const int tid = threadIdx.x;
    const bool rev = decreasing && decreasing[s] != 0;
    const bool suffix = scan_mode == XC_SCAN_SUFFIX;

    for (int n = tid; n < N; n += blockDim.x) {
        double acc = 0.0;
        const double* pp = part + ((size_t)s * C * K + k) * N + n;
        const size_t cs = (size_t)K * N;
        int c = 0;
        for (; c + 8 <= C; c += 8) {                 // 8 loads in flight, summed in order
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = pp[(size_t)(c + u) * cs];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += t[u];
        }
        for (; c < C; ++c) acc += pp[(size_t)c * cs];
        pd[suffix ? N - 1 - n : n] = acc;
        if (pdf) pdf[((size_t)s * K + k) * N + (rev ? N - 1 - n : n)] = acc;
    }
    __syncthreads();
    // Sequential running sum, exactly np.cumsum's order (core.py:1320): empty bins
    // leave the CDF bit-for-bit flat and the CDF of non-negative weights is monotone,
    // which the d/dA step and check_mono rely on.  N additions by one thread cost a
    // few microseconds; the parallelism of this kernel is across (slice, accumulator).
    if (tid == 0) serial_cumsum(pd, N);
    __syncthreads();
    if (tid == 0) total_s = pd[N - 1];
    __syncthreads();
    const double total = total_s;
    for (int r = tid; r < N; r += blockDim.x) {
        const int n = suffix ? N - 1 - r : r;          // bin position (ascending edges)
        double v = pd[r];
        if (scan_mode == XC_SCAN_TOTAL_MINUS) v = total - v;
        so.p[k][(size_t)s * so.stride + (rev ? N - 1 - n : n)] = v;
    }
}

struct HistPlan { int C; int warps; int ncopy; bool priv; size_t smem; };

static HistPlan plan_hist(long S, long P, int N, int K)
{
    HistPlan pl;
    const size_t budget = 200 * 1024;
    const size_t e_bytes = (size_t)((N + 2) & ~1) * 8;
    const size_t copy_bytes = (size_t)N * K * 8, tag_bytes = (size_t)((N + 15) & ~15);
    long fit = (long)((budget - e_bytes) / (copy_bytes + tag_bytes));
    // two resident CTAs per SM when 16 private copies fit in half the budget
    long fit_half = (long)((budget / 2 - e_bytes) / (copy_bytes + tag_bytes));
    if (fit_half >= HIST_MAX_WARPS)      { pl.priv = true;  pl.warps = HIST_MAX_WARPS; pl.ncopy = pl.warps; }
    else if (fit >= 8)                   { pl.priv = true;  pl.warps = (int)(fit > HIST_MAX_WARPS ? HIST_MAX_WARPS : fit); pl.ncopy = pl.warps; }
    else {
        pl.priv = false; pl.warps = HIST_MAX_WARPS;
        long nc = (long)((budget - e_bytes) / copy_bytes);
        pl.ncopy = (int)(nc > 8 ? 8 : nc);
    }
    pl.smem = e_bytes + (size_t)(pl.ncopy > 0 ? pl.ncopy : 0) * copy_bytes +
              (pl.priv ? (size_t)pl.warps * tag_bytes : 0);
    // CTAs per slice: fill two resident CTAs per SM for two waves without a ragged
    // third wave (floor, not ceil)
    long want = (long)sm_count() * 4;
    long C = want / S;
    long maxC = (P + 16383) / 16384;
    if (C > maxC) C = maxC;
    if (C < 1) C = 1;
    pl.C = (int)C;
    return pl;
}

static bool hist_use_match()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("XCB200_HIST_DEDUP"); v = (e && e[0] == 't') ? 0 : 1; }
    return v == 1;
}

template <typename QT, int K, int MODE>
static int launch_hist_mode(const HistParams& hp, const HistPlan& pl, long ns, cudaStream_t st)
{
    dim3 grid((unsigned)pl.C, (unsigned)ns);
    XC_CUDA_OK(cudaFuncSetAttribute(k_hist<QT, K, MODE>,
               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    k_hist<QT, K, MODE><<<grid, pl.warps * 32, pl.smem, st>>>(hp);
    XC_LAUNCH_OK();
    return 0;
}

template <typename QT, int K>
static int launch_hist(const HistParams& hp, const HistPlan& pl, long ns, cudaStream_t st)
{
    if (!pl.priv) return launch_hist_mode<QT, K, HIST_ATOMIC>(hp, pl, ns, st);
    return hist_use_match() ? launch_hist_mode<QT, K, HIST_MATCH>(hp, pl, ns, st)
                            : launch_hist_mode<QT, K, HIST_TAG>(hp, pl, ns, st);
}

}  // namespace xc

using namespace xc;

extern "C" size_t xc_bin_accumulate_workspace_bytes(long S, long P, int N, int K)
{
    if (S <= 0 || P <= 0 || N <= 0 || K <= 0) return 0;
    HistPlan pl = plan_hist(S, P, N, K);
    return 256 + (size_t)S * pl.C * K * N * sizeof(double);
}

size_t xc::bin_accumulate_ws_bytes_stencil(long S, int ny, int nx, int N)
{
    const size_t a = xc_bin_accumulate_workspace_bytes(S, (long)ny * nx, N, 2);
    const size_t b = 256 + bin_rows_part_doubles(S, ny, nx, N) * sizeof(double);
    return a > b ? a : b;
}

extern "C" int xc_bin_accumulate(const void* q, int q_dtype, long S, long P,
                                 const double* edges, long edges_stride, int N,
                                 int closed_right,
                                 const void* dA, int dA_dtype, int acc_area,
                                 const void* const* integrands, const int* integrand_dtypes,
                                 int n_int, const uint8_t* q_mask,
                                 int scan_mode, const int32_t* decreasing,
                                 double* pdf, double* cdf, int32_t* bin_idx,
                                 void* workspace, size_t ws_bytes, void* stream)
{
    XC_REQUIRE(cdf, "xc_bin_accumulate: null pointer");
    const int K = (acc_area ? 1 : 0) + n_int;
    ScanOut so;
    for (int k = 0; k < 4; ++k) so.p[k] = cdf + (size_t)k * N;
    so.stride = (long)K * N;
    return bin_accumulate_impl(q, q_dtype, S, P, edges, edges_stride, N, closed_right, dA, dA_dtype,
                               acc_area, integrands, integrand_dtypes, n_int, q_mask, scan_mode,
                               decreasing, pdf, so, bin_idx, workspace, ws_bytes, stream);
}

int xc::bin_accumulate_impl(const void* q, int q_dtype, long S, long P,
                            const double* edges, long edges_stride, int N,
                            int closed_right,
                            const void* dA, int dA_dtype, int acc_area,
                            const void* const* integrands, const int* integrand_dtypes,
                            int n_int, const uint8_t* q_mask,
                            int scan_mode, const int32_t* decreasing,
                            double* pdf, const ScanOut& so, int32_t* bin_idx,
                            void* workspace, size_t ws_bytes, void* stream,
                            const StencilArgs* stencil, HistOnly* hist_only)
{
    XC_REQUIRE(q && edges && dA, "xc_bin_accumulate: null pointer");
    XC_REQUIRE(S > 0 && P > 0 && N >= 1, "xc_bin_accumulate: need S>0, P>0, N>=1");
    XC_REQUIRE(q_dtype == XC_F32 || q_dtype == XC_F64, "xc_bin_accumulate: bad q dtype");
    XC_REQUIRE(n_int >= 0 && n_int <= XC_MAX_INTEGRANDS, "xc_bin_accumulate: n_int out of range");
    XC_REQUIRE(edges_stride == 0 || edges_stride == N + 1, "xc_bin_accumulate: edges_stride must be 0 or N+1");
    const int K = (acc_area ? 1 : 0) + n_int + (stencil ? 1 : 0);
    XC_REQUIRE(K >= 1 && K <= 4, "xc_bin_accumulate: nothing to accumulate");
    XC_REQUIRE(!stencil || (long)stencil->ny * stencil->nx == P, "xc_bin_accumulate: stencil shape");
    HistPlan pl = plan_hist(S, P, N, K);
    XC_REQUIRE(pl.ncopy >= 1, "xc_bin_accumulate: N=%d with K=%d does not fit shared memory", N, K);
    XC_REQUIRE(workspace && ws_bytes >= xc_bin_accumulate_workspace_bytes(S, P, N, K),
               "xc_bin_accumulate: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(workspace, ws_bytes);
    size_t part_doubles = (size_t)S * pl.C * K * N;
    if (hist_only && stencil && ws_bytes >= bin_accumulate_ws_bytes_stencil(S, stencil->ny, stencil->nx, N)) {
        const size_t r = bin_rows_part_doubles(S, stencil->ny, stencil->nx, N);
        if (r > part_doubles) part_doubles = r;
    }
    HistParams hp;
    hp.q = q; hp.P = P; hp.per = ((P + pl.C - 1) / pl.C + 3) & ~3L;
    hp.edges = edges; hp.edges_stride = edges_stride; hp.N = N; hp.closed_right = closed_right;
    hp.dA = dA; hp.dA_f32 = dA_dtype == XC_F32; hp.acc_area = acc_area;
    hp.n_int = n_int;
    for (int n = 0; n < XC_MAX_INTEGRANDS; ++n) {
        hp.integ[n] = n < n_int ? integrands[n] : nullptr;
        hp.integ_f32[n] = n < n_int ? (integrand_dtypes[n] == XC_F32) : 0;
        XC_REQUIRE(n >= n_int || hp.integ[n], "xc_bin_accumulate: null integrand");
    }
    hp.q_mask = q_mask;
    hp.stencil = stencil ? 1 : 0;
    if (stencil) {
        hp.ny = stencil->ny; hp.nx = stencil->nx; hp.cx = stencil->cx; hp.cy = stencil->cy;
    } else { hp.ny = hp.nx = 0; hp.cx = hp.cy = nullptr; }
    hp.part = ar.take<double>(part_doubles);
    hp.bin_idx = bin_idx;
    hp.ncopy = pl.ncopy;
    // the fused Keff pass (fp32 tracer and areas, {dA, |grad q|^2 dA}, uniform per-slice
    // edges, no mask / bin output) has a dedicated lean kernel
    if (hist_only && stencil && acc_area && n_int == 0 && !q_mask && !bin_idx && !closed_right &&
        edges_stride == N + 1 && stencil->dA_row && stencil->minmax) {
        int Cr = 0;
        const int r = bin_rows_try(q, q_dtype, S, edges, N, stencil, stencil->minmax, hp.part,
                                   part_doubles, &Cr, stream);
        if (r == 2) return 1;
        if (r == 0) { hist_only->part = hp.part; hist_only->C = Cr; return 0; }
    }
    const bool plain_latlon = stencil && stencil->bcx == XC_BC_PERIODIC && stencil->bcy == XC_BC_EXTEND;
    XC_REQUIRE(!stencil || plain_latlon, "xc_bin_accumulate: ghost-cell rules other than (periodic, extend) need an fp32 "
               "tracer with nx %% 4 == 0 and cell areas that are constant along x (dA_row)");
    for (long s0 = 0; s0 < S; s0 += 65535) {
        long ns = S - s0 < 65535 ? S - s0 : 65535;
        hp.s0 = s0;
        int rc;
        if (q_dtype == XC_F32) {
            rc = K == 1 ? launch_hist<float, 1>(hp, pl, ns, st) : K == 2 ? launch_hist<float, 2>(hp, pl, ns, st)
               : K == 3 ? launch_hist<float, 3>(hp, pl, ns, st) : launch_hist<float, 4>(hp, pl, ns, st);
        } else {
            rc = K == 1 ? launch_hist<double, 1>(hp, pl, ns, st) : K == 2 ? launch_hist<double, 2>(hp, pl, ns, st)
               : K == 3 ? launch_hist<double, 3>(hp, pl, ns, st) : launch_hist<double, 4>(hp, pl, ns, st);
        }
        if (rc) return rc;
    }
    if (hist_only) { hist_only->part = hp.part; hist_only->C = pl.C; return 0; }
    XC_REQUIRE(S <= 0x7fffffffL, "xc_bin_accumulate: too many slices");
    dim3 g2((unsigned)S, (unsigned)K);
    XC_REQUIRE((size_t)N * 8 <= 200 * 1024, "xc_bin_accumulate: N too large for the scan kernel");
    if ((size_t)N * 8 > 48 * 1024)
        XC_CUDA_OK(cudaFuncSetAttribute(k_reduce_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)N * 8)));
    k_reduce_scan<<<g2, 256, (size_t)N * 8, st>>>(hp.part, pl.C, K, N, scan_mode, decreasing, pdf, so);
    XC_LAUNCH_OK();
    return 0;
}
