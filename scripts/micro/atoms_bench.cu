// Microbenchmark: throughput of native shared-memory integer atomics (ATOMS.ADD.U32)
// on sm_100a, the building block of a fixed-point (exact, order-independent)
// alternative to the fp64 read-modify-write rounds of k_lwa_fast / k_hist_keff.
//   a) lane-private banks, random rows (no bank conflict, no address collision), no return
//   b) the same with the old value returned (needed for a carry into a high word)
//   c) 64-bit add as two ATOMS.ADD.U32 with carry (lo returns, hi does not)
//   d) random slots of a 361-entry table shared by the warp (bank conflicts + collisions)
//   e) d) restricted to a 32-slot window (heavy same-address collisions, smooth fields)
// Reported: cycles per warp-instruction (per 64-bit add for c) per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms_bench atoms_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ROWS 64
__global__ void k_a(const unsigned* keys, unsigned* out, int iters) {
    extern __shared__ unsigned T[];                       // [warps][ROWS][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned* Tw = T + warp * ROWS * 32;
    for (int i = lane; i < ROWS * 32; i += 32) Tw[i] = 0;
    __syncwarp();
    unsigned k = keys[threadIdx.x + blockIdx.x * blockDim.x];
    for (int i = 0; i < iters; ++i) {
        atomicAdd(&Tw[((k >> 8) % ROWS) * 32 + lane], k & 0xffu);
        k = k * 1664525u + 1013904223u;
    }
    __syncwarp();
    out[threadIdx.x + blockIdx.x * blockDim.x] = Tw[lane];
}
__global__ void k_b(const unsigned* keys, unsigned* out, int iters) {
    extern __shared__ unsigned T[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned* Tw = T + warp * ROWS * 32;
    for (int i = lane; i < ROWS * 32; i += 32) Tw[i] = 0;
    __syncwarp();
    unsigned k = keys[threadIdx.x + blockIdx.x * blockDim.x], acc = 0;
    for (int i = 0; i < iters; ++i) {
        acc += atomicAdd(&Tw[((k >> 8) % ROWS) * 32 + lane], k & 0xffu);
        k = k * 1664525u + 1013904223u;
    }
    __syncwarp();
    out[threadIdx.x + blockIdx.x * blockDim.x] = Tw[lane] + acc;
}
__global__ void k_c(const unsigned* keys, unsigned* out, int iters) {
    extern __shared__ unsigned T[];                       // [warps][ROWS/2][2][32]: lo words, hi words
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned* Tw = T + warp * ROWS * 32;
    for (int i = lane; i < ROWS * 32; i += 32) Tw[i] = 0;
    __syncwarp();
    unsigned k = keys[threadIdx.x + blockIdx.x * blockDim.x];
    for (int i = 0; i < iters; ++i) {
        const unsigned r = (k >> 8) % (ROWS / 2);
        const unsigned lo = k * 2654435761u, hi = k & 0xffu;
        const unsigned old = atomicAdd(&Tw[(2 * r) * 32 + lane], lo);
        const unsigned carry = (old + lo) < old ? 1u : 0u;
        atomicAdd(&Tw[(2 * r + 1) * 32 + lane], hi + carry);
        k = k * 1664525u + 1013904223u;
    }
    __syncwarp();
    out[threadIdx.x + blockIdx.x * blockDim.x] = Tw[lane] + Tw[32 + lane];
}
template <int WINDOW>
__global__ void k_d(const unsigned* keys, unsigned* out, int iters) {
    extern __shared__ unsigned T[];                       // [warps][384]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned* Tw = T + warp * 384;
    for (int i = lane; i < 384; i += 32) Tw[i] = 0;
    __syncwarp();
    unsigned k = keys[threadIdx.x + blockIdx.x * blockDim.x] * 2654435761u + threadIdx.x * 40503u;
    for (int i = 0; i < iters; ++i) {
        atomicAdd(&Tw[(k >> 10) % WINDOW], k & 0xffu);
        k = k * 1664525u + 1013904223u;
    }
    __syncwarp();
    out[threadIdx.x + blockIdx.x * blockDim.x] = Tw[lane];
}
// the incumbent: one conflict-free fp64-pair read-modify-write round (as in match_bench.cu)
__global__ void k_rmw(const unsigned* keys, unsigned* out, int iters) {
    extern __shared__ double2 H[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* Hw = H + warp * 128;
    for (int i = lane; i < 128; i += 32) Hw[i] = make_double2(0, 0);
    unsigned k = keys[threadIdx.x + blockIdx.x * blockDim.x];
    for (int i = 0; i < iters; ++i) {
        unsigned b = (k + lane * 11u) & 127u;
        double2 t = Hw[b]; t.x += 1.0; t.y += 2.0; Hw[b] = t;
        __syncwarp();
        k = k * 1664525u + 1013904223u;
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = (unsigned)Hw[lane].x;
}

// f) fp64 atomicAdd on shared memory (compiles to an ATOMS.CAST.SPIN.64 loop): random slots of a WINDOW-entry
//    table shared by ALL warps of the CTA (the alternative to warp-private histograms + lane de-duplication)
template <int WINDOW, int NACC>
__global__ void k_f(const unsigned* keys, unsigned* out, int iters) {
    __shared__ double T[2 * 384];
    for (int i = threadIdx.x; i < 2 * 384; i += blockDim.x) T[i] = 0.0;
    __syncthreads();
    unsigned k = keys[threadIdx.x + blockIdx.x * blockDim.x] * 2654435761u + threadIdx.x * 40503u;
    for (int i = 0; i < iters; ++i) {
        const unsigned b = (k >> 10) % WINDOW;
        atomicAdd(&T[b], 1.0 + (double)(k & 0xffu));
        if (NACC == 2) atomicAdd(&T[384 + b], 2.0);
        k = k * 1664525u + 1013904223u;
    }
    __syncthreads();
    out[threadIdx.x + blockIdx.x * blockDim.x] = (unsigned)T[threadIdx.x & 255];
}
template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    const int iters = 20000; int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned *keys, *out; cudaMalloc(&keys, 1 << 24); cudaMalloc(&out, 1 << 24);
    unsigned* h = new unsigned[1 << 22]; for (int i = 0; i < (1 << 22); ++i) h[i] = (i * 2654435761u) >> 7;
    cudaMemcpy(keys, h, 1 << 24, cudaMemcpyHostToDevice);
    const size_t smem = 200 * 1024;
    cudaFuncSetAttribute(k_a, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_b, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_c, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_d<361>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_d<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_d<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_rmw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const double cyc = 1.965e6;                           // cycles per ms at 1965 MHz
    for (int wps : {8, 16, 24}) {                          // ROWS*32*4 = 8 KB per warp -> at most 24 warps in 200 KB
        const int threads = 32 * wps, ctas = sms;
        const size_t sm = (size_t)wps * ROWS * 32 * 4;
        const float a = timeit([&] { k_a<<<ctas, threads, sm>>>(keys, out, iters); });
        const float b = timeit([&] { k_b<<<ctas, threads, sm>>>(keys, out, iters); });
        const float c = timeit([&] { k_c<<<ctas, threads, sm>>>(keys, out, iters); });
        const float d = timeit([&] { k_d<361><<<ctas, threads, sm>>>(keys, out, iters); });
        const float e = timeit([&] { k_d<32><<<ctas, threads, sm>>>(keys, out, iters); });
        const float f = timeit([&] { k_d<4><<<ctas, threads, sm>>>(keys, out, iters); });
        const float g = timeit([&] { k_rmw<<<ctas, threads, sm>>>(keys, out, iters); });
        const double n = (double)iters * wps;
        printf("warps/SM %2d: ATOMS.ADD private banks %.1f | with return %.1f | 64-bit add (2 ATOMS + carry) %.1f | "
               "random of 361 slots %.1f | random of 32 slots %.1f | random of 4 slots %.1f | fp64-pair RMW round %.1f  (cycles per warp-op per SM)\n",
               wps, a * cyc / n, b * cyc / n, c * cyc / n, d * cyc / n, e * cyc / n, f * cyc / n, g * cyc / n);
    }

    for (int wps : {16, 32, 64}) {
        const int threads = 32 * wps > 1024 ? 1024 : 32 * wps, ctas = sms * (32 * wps / threads);
        const int it2 = 4000;
        const float f1 = timeit([&] { k_f<361, 1><<<ctas, threads>>>(keys, out, it2); });
        const float f2 = timeit([&] { k_f<361, 2><<<ctas, threads>>>(keys, out, it2); });
        const float f3 = timeit([&] { k_f<32, 2><<<ctas, threads>>>(keys, out, it2); });
        const float f4 = timeit([&] { k_f<4, 2><<<ctas, threads>>>(keys, out, it2); });
        const double n = (double)it2 * wps;
        printf("warps/SM %2d (CTA-shared table): fp64 atomicAdd random of 361 slots %.1f | fp64 pair (2 atomicAdd) of 361 %.1f | pair of 32 slots %.1f | pair of 4 slots %.1f  (cycles per warp-op per SM)\n",
               wps, f1 * cyc / n, f2 * cyc / n, f3 * cyc / n, f4 * cyc / n);
    }
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    return 0;
}
