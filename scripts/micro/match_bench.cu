// Microbenchmark: throughput of MATCH.ANY, VOTE/ballot, shared-memory 128-bit RMW
// and the byte-tag round on sm_100a (cycles per warp-instruction per SM at 32 and
// 16 resident warps).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o match_bench match_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_match(const unsigned* keys, unsigned* out, int iters) {
    unsigned k = keys[threadIdx.x + blockIdx.x * blockDim.x], acc = 0;
    for (int i = 0; i < iters; ++i) { unsigned m = __match_any_sync(0xffffffffu, k); acc += m; k = k * 1664525u + (m & 7u); k &= 63u; }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
__global__ void k_ballot(const unsigned* keys, unsigned* out, int iters) {
    unsigned k = keys[threadIdx.x + blockIdx.x * blockDim.x], acc = 0;
    for (int i = 0; i < iters; ++i) { unsigned m = __ballot_sync(0xffffffffu, k & 1u); acc += m; k = k * 1664525u + (m & 7u); k &= 63u; }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
__global__ void k_rmw(const unsigned* keys, unsigned* out, int iters) {
    extern __shared__ double2 H[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* Hw = H + warp * 256;
    for (int i = lane; i < 256; i += 32) Hw[i] = make_double2(0, 0);
    unsigned k = keys[threadIdx.x + blockIdx.x * blockDim.x];
    for (int i = 0; i < iters; ++i) {
        unsigned b = (k + lane * 11u) & 255u;          // distinct slots per lane, conflict-free banks
        double2 t = Hw[b]; t.x += 1.0; t.y += 2.0; Hw[b] = t;
        __syncwarp();
        k = k * 1664525u + 1013904223u;
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = (unsigned)Hw[lane].x;
}
__global__ void k_rmw64(const unsigned* keys, unsigned* out, int iters) {
    extern __shared__ double2 H[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* Hw = reinterpret_cast<double*>(H + warp * 256);
    for (int i = lane; i < 512; i += 32) Hw[i] = 0.0;
    unsigned k = keys[threadIdx.x + blockIdx.x * blockDim.x];
    for (int i = 0; i < iters; ++i) {
        unsigned b = (k + lane * 11u) & 255u;
        Hw[b] += 1.0;
        __syncwarp();
        k = k * 1664525u + 1013904223u;
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = (unsigned)Hw[lane];
}
__global__ void k_lds_only(const unsigned* keys, unsigned* out, int iters) {
    extern __shared__ double2 H[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* Hw = H + warp * 256;
    for (int i = lane; i < 256; i += 32) Hw[i] = make_double2(i, 0);
    unsigned k = keys[threadIdx.x + blockIdx.x * blockDim.x]; double acc = 0;
    for (int i = 0; i < iters; ++i) {
        unsigned b = (k + lane * 11u) & 255u;
        double2 t = Hw[b]; acc += t.x + t.y;
        k = k * 1664525u + 1013904223u;
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = (unsigned)acc;
}
template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    const int iters = 20000; int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    unsigned *keys, *out; cudaMalloc(&keys, 1 << 24); cudaMalloc(&out, 1 << 24);
    unsigned* h = new unsigned[1 << 22]; for (int i = 0; i < (1 << 22); ++i) h[i] = (i * 2654435761u) >> 26;
    cudaMemcpy(keys, h, 1 << 24, cudaMemcpyHostToDevice);
    for (int wps : {4, 8, 16, 32}) {
        int threads = 32 * wps > 1024 ? 1024 : 32 * wps, ctas = sms * (32 * wps / threads);
        cudaFuncSetAttribute(k_rmw, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); cudaFuncSetAttribute(k_rmw64, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); cudaFuncSetAttribute(k_lds_only, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        float m1 = timeit([&] { k_match<<<ctas, threads>>>(keys, out, iters); });
        float m2 = timeit([&] { k_ballot<<<ctas, threads>>>(keys, out, iters); });
        float m3 = timeit([&] { k_rmw<<<ctas, threads, (threads / 32) * 256 * 16>>>(keys, out, iters); });
        double cyc = 1.965e6;   // cycles per ms at 1965 MHz
        float m4 = timeit([&] { k_rmw64<<<ctas, threads, (threads / 32) * 256 * 16>>>(keys, out, iters); });
        float m5 = timeit([&] { k_lds_only<<<ctas, threads, (threads / 32) * 256 * 16>>>(keys, out, iters); });
        printf("warps/SM %2d: MATCH.ANY %.1f cyc/warp-instr/SM | BALLOT %.1f | RMW128(LDS+2DADD+STS+syncwarp) %.1f | RMW64 %.1f | LDS128 only %.1f\n", wps,
               m1 * cyc / (double)(iters * wps), m2 * cyc / (double)(iters * wps), m3 * cyc / (double)(iters * wps),
               m4 * cyc / (double)(iters * wps), m5 * cyc / (double)(iters * wps));
    }
    return 0;
}
