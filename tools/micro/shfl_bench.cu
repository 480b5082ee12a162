// Microbenchmark: do warp shuffles and shared-memory gathers share one data return path?
// Each kernel gathers 16 doubles per lane per iteration by a random 3-bit index (the cfg3 shape:
// 8 leaves, order 13, fp64): all from shared memory (LDS.64), all from registers of other lanes
// (2 x SHFL.IDX per double), or half and half.  If the mixed loop costs max(lds, shfl) instead of
// the sum, the two paths are independent and the random-point evaluator can split its gathers.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NC = 16;  // doubles per record

__device__ __forceinline__ double shfl_d(double v, int src) {
    int lo = __shfl_sync(0xffffffffu, __double2loint(v), src);
    int hi = __shfl_sync(0xffffffffu, __double2hiint(v), src);
    return __hiloint2double(hi, lo);
}

__constant__ double crec[8 * 17];

// shared-memory gathers + constant-bank gathers (LDC with a per-lane index)
template <int N_LDS, int N_LDC>
__global__ void __launch_bounds__(512) k_ldc(const int* in, double* out, int iters) {
    __shared__ double rec[8 * 17];
    for (int i = threadIdx.x; i < 8 * 17; i += blockDim.x) rec[i] = 1.0 + i;
    __syncthreads();
    int v = in[threadIdx.x + blockIdx.x * blockDim.x];
    double acc = 0.0;
    for (int i = 0; i < iters; ++i) {
        const int leaf = v & 7;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < N_LDS; ++k) s += rec[leaf * 17 + k];
#pragma unroll
        for (int k = 0; k < N_LDC; ++k) s += crec[leaf * 17 + 16 - k];
        acc += s;
        v = v * 1664525 + 1013904223 + (int)__double2loint(s);
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}

template <int N_LDS, int N_SHFL>
__global__ void __launch_bounds__(512) k_gather(const int* in, double* out, int iters) {
    __shared__ double rec[8 * 17];
    for (int i = threadIdx.x; i < 8 * 17; i += blockDim.x) rec[i] = 1.0 + i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double own[NC];  // lane l holds record (l & 7) in registers
#pragma unroll
    for (int k = 0; k < NC; ++k) own[k] = rec[(lane & 7) * 17 + k];
    int v = in[threadIdx.x + blockIdx.x * blockDim.x];
    double acc = 0.0;
    for (int i = 0; i < iters; ++i) {
        const int leaf = v & 7;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < N_LDS; ++k) s += rec[leaf * 17 + k];
#pragma unroll
        for (int k = 0; k < N_SHFL; ++k) s += shfl_d(own[NC - 1 - k], (lane & 24) | leaf);
        acc += s;
        v = v * 1664525 + 1013904223 + (int)__double2loint(s);
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}

// fp32 variant: 9 floats per record (the cfg2 shape), 15 leaves
template <int N_LDS, int N_SHFL>
__global__ void __launch_bounds__(512) k_gather32(const int* in, float* out, int iters) {
    __shared__ float rec[16 * 9];
    for (int i = threadIdx.x; i < 16 * 9; i += blockDim.x) rec[i] = 1.0f + i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float own[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) own[k] = rec[(lane & 15) * 9 + k];
    int v = in[threadIdx.x + blockIdx.x * blockDim.x];
    float acc = 0.f;
    for (int i = 0; i < iters; ++i) {
        const int leaf = v & 15;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < N_LDS; ++k) s += rec[leaf * 9 + k];
#pragma unroll
        for (int k = 0; k < N_SHFL; ++k) s += __shfl_sync(0xffffffffu, own[8 - k], (lane & 16) | leaf);
        acc += s;
        v = v * 1664525 + 1013904223 + __float_as_int(s);
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}

int main() {
    const int blocks = 148, threads = 512, iters = 8192;
    int* in; double* out;
    cudaMalloc(&in, blocks * threads * 4); cudaMalloc(&out, blocks * threads * 8);
    int* h = new int[blocks * threads];
    for (int i = 0; i < blocks * threads; ++i) h[i] = rand();
    cudaMemcpy(in, h, blocks * threads * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    auto report = [&](const char* name, float ms) {
        double warp_iters = (double)blocks * threads / 32 * iters;
        double cyc = ms * 1e-3 * clk * 1e3;
        printf("%-22s %.3f ms  -> %.2f SM-cycles per warp-iteration\n", name, ms, cyc * 148 / warp_iters);
    };
#define RUN(NAME, K, OUT)                                                                          \
    {                                                                                              \
        K<<<blocks, threads>>>(in, OUT, 16); cudaDeviceSynchronize();                              \
        cudaEventRecord(e0); K<<<blocks, threads>>>(in, OUT, iters); cudaEventRecord(e1);          \
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); report(NAME, ms);   \
    }
    RUN("f64 lds16 shfl0", (k_gather<16, 0>), out)
    RUN("f64 lds0 shfl16", (k_gather<0, 16>), out)
    RUN("f64 lds8 shfl8", (k_gather<8, 8>), out)
    RUN("f64 lds8 shfl0", (k_gather<8, 0>), out)
    RUN("f64 lds0 shfl8", (k_gather<0, 8>), out)
    RUN("f64 lds12 shfl4", (k_gather<12, 4>), out)
    {
        double hc[8 * 17];
        for (int i = 0; i < 8 * 17; ++i) hc[i] = 1.0 + i;
        cudaMemcpyToSymbol(crec, hc, sizeof(hc));
    }
    RUN("f64 lds0 ldc16", (k_ldc<0, 16>), out)
    RUN("f64 lds0 ldc4", (k_ldc<0, 4>), out)
    RUN("f64 lds12 ldc4", (k_ldc<12, 4>), out)
    RUN("f64 lds13 ldc3", (k_ldc<13, 3>), out)
    RUN("f64 lds14 ldc2", (k_ldc<14, 2>), out)
    RUN("f64 lds12 ldc0", (k_ldc<12, 0>), out)
    RUN("f32 lds9 shfl0", (k_gather32<9, 0>), (float*)out)
    RUN("f32 lds0 shfl9", (k_gather32<0, 9>), (float*)out)
    RUN("f32 lds5 shfl4", (k_gather32<5, 4>), (float*)out)
    RUN("f32 lds5 shfl0", (k_gather32<5, 0>), (float*)out)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
