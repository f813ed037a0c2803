// Shared helpers for libxcb200 (sm_100a).  Host-side error/launch bookkeeping
// and small device utilities used by every kernel file.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/xcb200.h"

namespace xc {

// ---- host bookkeeping ------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int  sm_count();

#define XC_REQUIRE(cond, ...)                                   \
    do { if (!(cond)) { xc::set_error(__VA_ARGS__); return 1; } } while (0)

#define XC_CUDA_OK(call)                                                       \
    do { cudaError_t e__ = (call);                                             \
         if (e__ != cudaSuccess) {                                             \
             xc::set_error("%s failed: %s (%s:%d)", #call,                     \
                           cudaGetErrorString(e__), __FILE__, __LINE__);       \
             return 1; } } while (0)

#define XC_LAUNCH_OK()                                                         \
    do { xc::count_launch();                                                   \
         cudaError_t e__ = cudaGetLastError();                                 \
         if (e__ != cudaSuccess) {                                             \
             xc::set_error("kernel launch failed: %s (%s:%d)",                 \
                           cudaGetErrorString(e__), __FILE__, __LINE__);       \
             return 1; } } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over a caller-provided workspace
struct Arena {
    char* base; size_t size; size_t off;
    Arena(void* p, size_t n) : base((char*)p), size(n), off(0) {}
    template <typename T> T* take(size_t count) {
        off = align_up(off, 256);
        T* r = (T*)(base + off);
        off += count * sizeof(T);
        return r;
    }
    bool ok() const { return off <= size; }
};

// ---- device helpers --------------------------------------------------------
#define XC_FULL 0xffffffffu

__device__ __forceinline__ double ld_as_f64(const void* p, long i, int is_f32) {
    return is_f32 ? (double)__ldg((const float*)p + i) : __ldg((const double*)p + i);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(XC_FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(XC_FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(XC_FULL, v, o));
    return v;
}

// Byte "tag" election in shared memory (used by the warp-private scatter-adds):
// several lanes may store different ids to the same byte at once and exactly one
// value sticks.  The accesses are RELAXED ATOMIC stores/loads at CTA scope -- they
// compile to ordinary STS.U8 / LDS.U8, but being morally strong they do not
// constitute a data race under the PTX memory model (this is the classic
// SDK-histogram "tagged write" idiom, stated in terms of the formal model).
__device__ __forceinline__ void tag_store(uint8_t* p, unsigned v) {
    asm volatile("st.relaxed.cta.shared.u8 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned tag_load(const uint8_t* p) {
    unsigned v;
    asm volatile("ld.relaxed.cta.shared.u8 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}

// np.interp for one x against ascending xp[0..n) / fp[0..n) with numpy's exact
// arithmetic (numpy/_core/src/multiarray/compiled_base.c, arr_interp): clamp to
// the end values, return fp[j] on an exact hit, otherwise
// slope*(x-xp[j])+fp[j] with separately rounded multiply and add.  `rev` walks
// the arrays backwards (np.interp(x, xp[::-1], fp[::-1]), core.py:1430).
__device__ __forceinline__ double np_interp(double x, const double* xp,
                                            const double* fp, int n, bool rev) {
    if (isnan(x)) return x;
    auto X = [&](int j) { return rev ? xp[n - 1 - j] : xp[j]; };
    auto F = [&](int j) { return rev ? fp[n - 1 - j] : fp[j]; };
    if (n == 1) return F(0);
    if (x > X(n - 1)) return F(n - 1);
    if (x < X(0)) return F(0);
    int lo = 0, hi = n;                       // largest j with X(j) <= x
    while (lo < hi) {
        int mid = lo + ((hi - lo) >> 1);
        if (x >= X(mid)) lo = mid + 1; else hi = mid;
    }
    int j = lo - 1;
    if (j < 0) return F(0);                   // only reachable with NaN in xp
    if (j >= n - 1) return F(n - 1);
    double xj = X(j), fj = F(j);
    if (xj == x) return fj;
    double xj1 = X(j + 1), fj1 = F(j + 1);
    double slope = __ddiv_rn(__dsub_rn(fj1, fj), __dsub_rn(xj1, xj));
    double r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xj)), fj);
    if (isnan(r)) {
        r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xj1)), fj1);
        if (isnan(r) && fj == fj1) r = fj;
    }
    return r;
}

// np.cumsum's order, one addition after the other; the loads and stores of sixteen elements are batched around the
// dependent chain of additions (element by element the loop paid a shared-memory round trip per addition: ~40
// cycles per element, 42 us of config 5's 90 us epilogue with its 2048 levels)
__device__ __forceinline__ void serial_cumsum(double* a, int N)
{
    double run = 0.0; int r = 0;
    for (; r + 16 <= N; r += 16) {
        double t[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) t[u] = a[r + u];
#pragma unroll
        for (int u = 0; u < 16; ++u) { run = __dadd_rn(run, t[u]); t[u] = run; }
#pragma unroll
        for (int u = 0; u < 16; ++u) a[r + u] = t[u];
    }
    for (; r < N; ++r) { run = __dadd_rn(run, a[r]); a[r] = run; }
}

}  // namespace xc
