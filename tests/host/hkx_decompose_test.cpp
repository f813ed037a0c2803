// Host check of xcontour_b200/csrc/hkx_decompose.cuh: the lean decomposition returns the same
// (window, 96-bit value) as the reference statement for every term.  Built and run by
// tests/test_fixed_point_model.py::test_lean_term_decomposition_equals_the_default (g++).
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include "../../xcontour_b200/csrc/hkx_decompose.cuh"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t next() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

int main()
{
    const int nw = 6, wbits = 24;
    long n = 0, acc = 0, bad = 0;
    for (long it = 0; it < 20000000; ++it) {
        uint64_t bits = next();
        int e_base = -11 + (int)(next() % 2070);                 // anchors as the kernel can produce them (>= 1 - 12) and beyond the top
        if (it % 3 == 0) {                                   // exponents around the window range
            const int ex = e_base - 60 + (int)(next() % (nw * wbits + 70));
            if (ex > 0 && ex < 0x7ff) bits = (bits & 0x800FFFFFFFFFFFFFull) | ((uint64_t)ex << 52);
        }
        if (it % 7 == 0) bits &= ~(1ull << 63);
        const int hi = (int)(uint32_t)(bits >> 32); const uint32_t lo = (uint32_t)bits;
        HkxTerm a, b; memset(&a, 0, sizeof a); memset(&b, 0, sizeof b);
        const int ra = hkx_decompose_ref(hi, lo, e_base, nw, wbits, a);
        const int rb = hkx_decompose_lean(hi, lo, e_base, nw, wbits, b);
        ++n; acc += (ra == 0);
        if (ra != rb || (ra == 0 && (a.w != b.w || a.v0 != b.v0 || a.v1 != b.v1 || a.v2 != b.v2))) {
            if (bad++ < 5) printf("MISMATCH hi=%08x lo=%08x e_base=%d\n", (unsigned)hi, lo, e_base);
        }
    }
    printf("%ld terms, %ld accumulated, %ld mismatches\n", n, acc, bad);
    return bad != 0;
}
