// Microbenchmark: throughput of the conversion instructions (XU pipe) that the
// fixed-point kernels would lean on, against magic-number equivalents on the
// fp64 / integer pipes.  Cycles per warp-instruction per SM, 32 warps per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cvt_bench cvt_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define UNROLL 8
template <int OP>
__global__ void k(const float* in, double* out, int iters) {
    const int t = threadIdx.x + blockIdx.x * blockDim.x;
    float f[UNROLL]; double d[UNROLL]; long long l[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { f[u] = in[t + u]; d[u] = (double)in[t + u + 8] * 1e6; l[u] = (long long)(in[t + u] * 1e9f); }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (OP == 0) { d[u] += (double)f[u]; f[u] = __int_as_float(__float_as_int(f[u]) + 1); }                 // F2F.F64.F32 (+DADD, IADD)
            if (OP == 1) { l[u] += __double2ll_rn(d[u]); d[u] = __longlong_as_double(__double_as_longlong(d[u]) + 1); } // F2I.S64.F64
            if (OP == 2) { d[u] += (double)l[u]; l[u] += 3; }                                                    // I2F.F64.S64
            if (OP == 3) { l[u] += __float2int_rz(f[u]); f[u] = __int_as_float(__float_as_int(f[u]) + 1); }        // F2I.TRUNC
            if (OP == 4) { d[u] += 1.0; f[u] = __int_as_float(__float_as_int(f[u]) + 1); }                         // baseline of OP 0: DADD + IADD
            if (OP == 5) { const double tt = d[u] + 6755399441055744.0; l[u] += __double_as_longlong(tt) - 0x4338000000000000ll;
                           d[u] = __longlong_as_double(__double_as_longlong(d[u]) + 1); }                          // magic double -> int64
            if (OP == 6) { const unsigned lo = (unsigned)l[u]; const int hi = (int)(l[u] >> 32);
                           const double dlo = __hiloint2double(0x43300000, lo) - 4503599627370496.0;
                           const double dhi = __hiloint2double(0x43300000, hi ^ 0x80000000) - 4503601774854144.0;
                           d[u] += fma(dhi, 4294967296.0, dlo); l[u] += 3; }                                      // magic int64 -> double
            if (OP == 7) { const unsigned uu = __float_as_uint(f[u]);
                           const unsigned hi = (((uu & 0x7fffffffu) >> 3) + 0x38000000u) | (uu & 0x80000000u);
                           d[u] += __hiloint2double(hi, uu << 29); f[u] = __int_as_float(__float_as_int(f[u]) + 1); } // bit-trick float -> double (normal numbers)
        }
    }
    double acc = 0; 
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += d[u] + (double)l[u] + f[u];
    out[t] = acc;
}
template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 1024, ctas = sms, iters = 4000;
    float* in; double* out; cudaMalloc(&in, (threads * ctas + 64) * 4); cudaMalloc(&out, threads * ctas * 8);
    float* h = new float[threads * ctas + 64]; for (int i = 0; i < threads * ctas + 64; ++i) h[i] = 1.0f + (i % 977) * 0.37f;
    cudaMemcpy(in, h, (threads * ctas + 64) * 4, cudaMemcpyHostToDevice);
    const double cyc = 1.965e6, n = (double)iters * UNROLL * 32;      // warp-ops per SM
    const char* names[8] = {"F2F.F64.F32 (+DADD+IADD)", "F2I.S64.F64 (+2 IADD)", "I2F.F64.S64 (+DADD+IADD)", "F2I.TRUNC (+IADD x2)",
                            "baseline DADD+IADD", "magic double->int64 (DADD + int)", "magic int64->double (2 DADD + DFMA + int)", "bit-trick float->double (+DADD+IADD)"};
    float ms[8];
    ms[0] = timeit([&] { k<0><<<ctas, threads>>>(in, out, iters); }); ms[1] = timeit([&] { k<1><<<ctas, threads>>>(in, out, iters); });
    ms[2] = timeit([&] { k<2><<<ctas, threads>>>(in, out, iters); }); ms[3] = timeit([&] { k<3><<<ctas, threads>>>(in, out, iters); });
    ms[4] = timeit([&] { k<4><<<ctas, threads>>>(in, out, iters); }); ms[5] = timeit([&] { k<5><<<ctas, threads>>>(in, out, iters); });
    ms[6] = timeit([&] { k<6><<<ctas, threads>>>(in, out, iters); }); ms[7] = timeit([&] { k<7><<<ctas, threads>>>(in, out, iters); });
    for (int o = 0; o < 8; ++o) printf("%-48s %.2f cycles per warp-op per SM\n", names[o], ms[o] * cyc / n);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    return 0;
}
