// Decomposition of one fp64 term into its window and 96-bit shifted mantissa for the
// exponent-windowed accumulators of k_hist_keff<true> (hist_keff.cu), as pure functions of
// the term's two 32-bit words so that they compile for the host as well:
//   hkx_decompose_ref   the arithmetic of the default build, statement for statement
//   hkx_decompose_lean  the same values from 32-bit funnel shifts (-DXC_HKX_LEAN=1 build)
// tests/test_fixed_point_model.py compiles both for the CPU and compares them bit for bit.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HKX_HD __host__ __device__ __forceinline__
#else
#define HKX_HD static inline
#endif

struct HkxTerm { int w; uint32_t v0, v1, v2; };   // window index and the 96-bit value (v2:v1:v0)

// 0: accumulate `t`; 1: the term takes the fp64 side table (negative, non-finite, denormal, zero,
// above the top window)
HKX_HD int hkx_decompose_ref(int hi, uint32_t lo, int e_base, int nw, int wbits, HkxTerm& t)
{
    const int ex = (hi >> 20) & 0x7ff;
    int rel = ex - e_base;
    if (hi < 0 || ex == 0x7ff || ex == 0 || rel >= nw * wbits) return 1;
    unsigned long long m = ((unsigned long long)(uint32_t)((hi & 0xfffff) | 0x100000) << 32) | lo;
    if (rel < 0) { m = rel > -53 ? (m >> (-rel)) : 0ull; rel = 0; }
    const int w = (rel * 2731) >> 16;                        // rel / 24 for rel < 8192
    const int sh = rel - w * wbits;
    const unsigned long long v = m << sh;
    t.w = w; t.v0 = (uint32_t)v; t.v1 = (uint32_t)(v >> 32);
    t.v2 = sh > 11 ? (uint32_t)(m >> (64 - sh)) : 0u;
    return 0;
}

HKX_HD uint32_t hkx_funnel_l(uint32_t lo, uint32_t hi, int sh)   // upper word of (hi:lo) << sh, 0 <= sh < 32
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, sh);
#else
    return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
#endif
}

// Accepted terms are positive, normal, finite and below the top window: hi in
// [0x00100000, min(0x7ff, e_base + nw*wbits) << 20) -- one unsigned comparison.
HKX_HD uint32_t hkx_accept_span(int e_base, int nw, int wbits)
{
    const int top = e_base + nw * wbits;
    return ((uint32_t)(top < 0x7ff ? (top > 1 ? top : 1) : 0x7ff) << 20) - 0x00100000u;
}

HKX_HD int hkx_decompose_lean(int hi, uint32_t lo, int e_base, int nw, int wbits, HkxTerm& t)
{
    if ((uint32_t)hi - 0x00100000u >= hkx_accept_span(e_base, nw, wbits)) return 1;
    int rel = (hi >> 20) - e_base;                           // hi >= 0 here: no mask needed
    uint32_t mh = (uint32_t)((hi & 0xfffff) | 0x100000), ml = lo;
    if (rel < 0) {                                           // below the anchor: truncate (rare)
        const unsigned long long m = (((unsigned long long)mh << 32) | ml);
        const unsigned long long r = rel > -53 ? (m >> (-rel)) : 0ull;
        mh = (uint32_t)(r >> 32); ml = (uint32_t)r; rel = 0;
    }
    const int w = (rel * 2731) >> 16;
    const int sh = rel - w * wbits;                          // 0 <= sh < wbits <= 24
    t.w = w;
    t.v0 = ml << sh;
    t.v1 = hkx_funnel_l(ml, mh, sh);
    t.v2 = hkx_funnel_l(mh, 0u, sh);
    return 0;
}
