// Microbenchmark: cost of __match_any_sync vs ballot-on-bits ranking vs smem atomics, per warp.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_match(const int* in, int* out, int iters) {
    int v = in[threadIdx.x + blockIdx.x * blockDim.x];
    int acc = 0;
    for (int i = 0; i < iters; ++i) {
        unsigned m = __match_any_sync(0xffffffffu, v & 7);
        acc += __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
        v = v * 1664525 + 1013904223 + acc;
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
__global__ void k_ballot(const int* in, int* out, int iters) {
    int v = in[threadIdx.x + blockIdx.x * blockDim.x];
    int acc = 0;
    for (int i = 0; i < iters; ++i) {
        int li = v & 7;
        unsigned peers = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            bool p = (li >> j) & 1;
            unsigned b = __ballot_sync(0xffffffffu, p);
            peers &= p ? b : ~b;
        }
        acc += __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
        v = v * 1664525 + 1013904223 + acc;
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
__global__ void k_atom(const int* in, int* out, int iters) {
    __shared__ int cnt[8 * 32];
    if (threadIdx.x < 256) cnt[threadIdx.x] = 0;
    __syncthreads();
    int v = in[threadIdx.x + blockIdx.x * blockDim.x];
    int acc = 0;
    int w = threadIdx.x >> 5;
    for (int i = 0; i < iters; ++i) {
        acc += atomicAdd(&cnt[(w & 7) * 32 + (v & 7)], 1);
        v = v * 1664525 + 1013904223 + acc;
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
__global__ void k_base(const int* in, int* out, int iters) {
    int v = in[threadIdx.x + blockIdx.x * blockDim.x];
    int acc = 0;
    for (int i = 0; i < iters; ++i) {
        acc += __popc((unsigned)v & ((1u << (threadIdx.x & 31)) - 1u));
        v = v * 1664525 + 1013904223 + acc;
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
int main() {
    const int blocks = 148 * 4, threads = 256, iters = 4096;
    int *in, *out;
    cudaMalloc(&in, blocks * threads * 4); cudaMalloc(&out, blocks * threads * 4);
    int* h = new int[blocks * threads];
    for (int i = 0; i < blocks * threads; ++i) h[i] = rand();
    cudaMemcpy(in, h, blocks * threads * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    auto run = [&](const char* name, void (*k)(const int*, int*, int)) {
        k<<<blocks, threads>>>(in, out, 16); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<<<blocks, threads>>>(in, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        // warp-iterations per SM per cycle
        double warp_iters = (double)blocks * threads / 32 * iters;
        double cyc = ms * 1e-3 * clk * 1e3;
        printf("%-8s %.3f ms  -> %.2f SM-cycles per warp-iteration (throughput, 32 warps/SM resident)\n", name, ms, cyc * 148 / warp_iters);
    };
    run("base", k_base); run("match", k_match); run("ballot3", k_ballot); run("atoms", k_atom);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
