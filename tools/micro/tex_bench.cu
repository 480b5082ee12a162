// Microbenchmark: does the TEXTURE path return data beside the shared-memory path?
// The cfg3 shape (8 records of 16 doubles, a random record per lane): N_LDS gathers from shared memory
// (LDS.64) and N_TEX gathers of the same words through a linear texture (tex1Dfetch<int2>, or
// <int4> = two doubles per fetch) per iteration.  If the mixed loop costs max(lds, tex) instead of
// the sum, the two return paths are independent and the gather-bound evaluator can split its
// coefficient fetches between them.  Also: an L1-cached global gather (LDG.64) for comparison.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tex_bench tex_bench.cu && ./tex_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int N_LDS, int N_TEX2, int N_TEX4, int N_LDG>
__global__ void __launch_bounds__(512) k(const int* in, double* out, int iters, cudaTextureObject_t t2,
                                         cudaTextureObject_t t4, const double* __restrict__ grec) {
    __shared__ double rec[8 * 17];
    for (int i = threadIdx.x; i < 8 * 17; i += blockDim.x) rec[i] = 1.0 + i;
    __syncthreads();
    int v = in[threadIdx.x + blockIdx.x * blockDim.x];
    double acc = 0.0;
    for (int i = 0; i < iters; ++i) {
        const int leaf = v & 7;
        double s = 0.0;
#pragma unroll
        for (int k2 = 0; k2 < N_LDS; ++k2) s += rec[leaf * 17 + k2];
#pragma unroll
        for (int k2 = 0; k2 < N_TEX2; ++k2) {
            const int2 w = tex1Dfetch<int2>(t2, leaf * 16 + 15 - k2);
            s += __hiloint2double(w.y, w.x);
        }
#pragma unroll
        for (int k2 = 0; k2 < N_TEX4; ++k2) {
            const int4 w = tex1Dfetch<int4>(t4, leaf * 8 + 7 - k2);
            s += __hiloint2double(w.y, w.x) + __hiloint2double(w.w, w.z);
        }
#pragma unroll
        for (int k2 = 0; k2 < N_LDG; ++k2) s += __ldg(grec + leaf * 16 + 15 - k2);
        acc += s;
        v = v * 1664525 + 1013904223 + (int)__double2loint(s);
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}

template <int A, int B, int C, int D>
void run(const char* name, const int* in, double* out, cudaTextureObject_t t2, cudaTextureObject_t t4, const double* g) {
    const int iters = 2000, grid = 148;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<A, B, C, D><<<grid, 512>>>(in, out, 10, t2, t4, g);
    cudaEventRecord(e0);
    k<A, B, C, D><<<grid, 512>>>(in, out, iters, t2, t4, g);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // SM cycles per warp-iteration (16 warps per SM share the paths)
    const double cyc = ms * 1e-3 * clk * 1e3 / iters / 16.0;
    printf("%-34s lds %2d tex2 %2d tex4 %2d ldg %2d : %7.3f ms, %6.1f SM-clk per warp-iteration (%s)\n", name, A, B, C, D, ms, cyc,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int* in; double* out; double* grec;
    cudaMalloc(&in, 148 * 512 * 4); cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&grec, 8 * 16 * 8);
    int h[148 * 512];
    for (int i = 0; i < 148 * 512; ++i) h[i] = i * 2654435761u;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    double hr[128];
    for (int i = 0; i < 128; ++i) hr[i] = 1.0 + i;
    cudaMemcpy(grec, hr, sizeof hr, cudaMemcpyHostToDevice);
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = grec;
    rd.res.linear.sizeInBytes = sizeof hr;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t t2 = 0, t4 = 0;
    rd.res.linear.desc = cudaCreateChannelDesc<int2>();
    cudaCreateTextureObject(&t2, &rd, &td, nullptr);
    rd.res.linear.desc = cudaCreateChannelDesc<int4>();
    cudaCreateTextureObject(&t4, &rd, &td, nullptr);
    printf("texture objects: %s\n", cudaGetErrorString(cudaGetLastError()));
    run<14, 0, 0, 0>("14 LDS.64 (today's cfg3 gather)", in, out, t2, t4, grec);
    run<8, 0, 0, 0>("8 LDS.64", in, out, t2, t4, grec);
    run<0, 14, 0, 0>("14 TEX int2", in, out, t2, t4, grec);
    run<0, 6, 0, 0>("6 TEX int2", in, out, t2, t4, grec);
    run<0, 0, 7, 0>("7 TEX int4 (14 doubles)", in, out, t2, t4, grec);
    run<0, 0, 3, 0>("3 TEX int4 (6 doubles)", in, out, t2, t4, grec);
    run<8, 6, 0, 0>("8 LDS.64 + 6 TEX int2", in, out, t2, t4, grec);
    run<8, 0, 3, 0>("8 LDS.64 + 3 TEX int4", in, out, t2, t4, grec);
    run<10, 0, 2, 0>("10 LDS.64 + 2 TEX int4", in, out, t2, t4, grec);
    run<0, 0, 0, 14>("14 LDG.64 (L1-cached)", in, out, t2, t4, grec);
    run<8, 0, 0, 6>("8 LDS.64 + 6 LDG.64", in, out, t2, t4, grec);
    return 0;
}
