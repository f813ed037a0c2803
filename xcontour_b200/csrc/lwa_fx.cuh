// Shared pieces of the two fixed-point LWA kernels (lwa_fx.cu: general weights, lwa_cols.cu: weights constant
// along a row): per-slice scales, the LUT over Q, value helpers, and the launchers lwa.cu dispatches to.
#pragma once
#include "common.cuh"
#include "internal.h"
#include <math_constants.h>

namespace xc {

constexpr int FX_SEG = 64;                 // row segments per column (threads = FX_SEG * TC)
constexpr int FX_LUT = 4096;               // buckets of the LUT over Q
constexpr int FX_TOTP = FX_SEG + 2;        // padded row of the totals table
constexpr int FX_PREP_NT = 256;
constexpr long FX_CHUNK = 1024;            // slices per fixed-point launch (bounds the per-slice LUT scratch)

// fixed-point scales of a slice: X_S = rn(w sS), X_V = rn(w (v - c) sV); iS = 1/sS, iV = 1/sV (powers of two)
struct FxScale { double c, sS, sV, iS, iV, pad0, pad1, pad2; };

// (Conversions: F2I / I2F with a 64-bit side cost ~2.3 cycles per warp instruction on the XU pipe,
// scripts/micro/cvt_bench.cu; the magic-number forms on the fp64 / integer pipes measure the same or worse and
// cost issue slots, which is what these kernels are short of, so the hardware conversions stay.)
// LUT bucket of a value: buckets 1 .. FX_LUT-2 divide [Q_first, Q_last) evenly, bucket 0 is everything below
// Q_first (no threshold lives there) and bucket FX_LUT-1 everything from Q_last on (only Q_last and its ties), so
// values outside the profile need no search and no special case.  NaN -> bucket 0.  Monotone in vf, and the SAME
// function places the profile's own values when the LUT is built (k_lwa_fx_prep).
__host__ __device__ __forceinline__ float fx_scale(double qmin, double qmax)
{
    return (qmax > qmin) ? (float)((double)(FX_LUT - 2) / (qmax - qmin)) : 0.0f;
}
__device__ __forceinline__ int fx_bucket(float vf, float qminf, float scalef)
{
    const float t = fminf(fmaxf(fmaf(vf - qminf, scalef, 1.0f), 0.0f), (float)(FX_LUT - 1));
    return __float2int_rz(t);
}
__device__ __forceinline__ long long fx_rn(double x) { return __double2ll_rn(x); }
__device__ __forceinline__ double fx_to_double(long long r) { return (double)r; }
// sign-adjusted value of a cell in fp32 (exact) / fp64 and its fp32 image for the LUT
__device__ __forceinline__ void fx_value(float qraw, float sgf, double, double& v, float& vf) { vf = sgf * qraw; v = (double)vf; }
__device__ __forceinline__ void fx_value(double qraw, float, double sg, double& v, float& vf) { v = sg * qraw; vf = (float)v; }

// ---- host-side launchers (each file owns its kernels) ----
// per-slice scales + LUT for slices [s0, s0 + ns) into fxs[ns], lutg[ns][FX_LUT]
int lwa_fx_prep_launch(long s0, long ns, int n_eq, const double* Qref, int increase, int32_t* sorted, int32_t* any_unsorted,
                       const double* rng, int rngC, const double* wmax_parts, int n_wmax, FxScale* fxs, uint32_t* lutg, void* stream);
// general weights ww[n_eq][n_x]; returns 1 when the planes do not fit shared memory
bool lwa_fx_fits(int n_eq);
int lwa_fx_launch(const void* q, int q_dtype, long s0, long ns, int n_eq, int n_x, const double* Qref, const double* ww,
                  int increase, int part, const int32_t* sorted, const FxScale* fxs, const uint32_t* lutg, double* out, void* stream);
// weights constant along a row (ww_row[n_eq]); out: fp64, or fp32 with out_f32
bool lwa_cols_fits(int n_eq, int qbytes);
int lwa_cols_launch(const void* q, int q_dtype, long s0, long ns, int n_eq, int n_x, const double* Qref, const double* ww_row,
                    int increase, int part, const int32_t* sorted, const FxScale* fxs, const uint32_t* lutg,
                    void* out, int out_f32, void* stream);

}  // namespace xc
