// ai_bucket.cuh -- leaf-bucketed evaluation for RANDOM points on small tables at high order.
//
// Why: with random points every lane gathers its own leaf record, and a gather of (n+3) elements
// costs (n+3)*sizeof(T)/4 cycles of the 128 B/clk shared-memory return path per 32 points -- 33
// cycles at fp64 order 13, which caps that case at 60% of HBM (profiles/r1_ncu_cfg3_*).  Here each
// warp re-groups its tile of 32 x V points by leaf through a warp-private shared-memory buffer
// (ranks from shared atomics, ~4 cycles per warp-instruction for <= 8 leaves), hands each lane a
// chunk of up to C same-leaf points, and evaluates the chunk with ONE record held in registers.
// The arithmetic per point is the same eval_record(), so the output bits equal eval_kernel's.
// Tiles whose points all sit in one leaf (sorted inputs) skip the regrouping entirely.
#pragma once
#include "ai_kernels.cuh"

namespace ai {

constexpr int kBucketBlock = 256;          // threads per CTA (8 warps, each fully independent)
constexpr int kBucketV = 8;                // points per lane per tile  -> 256 points per warp tile
constexpr int kBucketMaxLeaves = 8;
// shared memory per warp: x[256] + y[256] + origin[256] (u8) + 3 small int arrays
template <typename T> __host__ __device__ constexpr int bucket_warp_bytes() {
    return 2 * 256 * (int)sizeof(T) + 256 + 3 * 16 * 4;
}

template <int BASIS, int ORDER, typename T>
__global__ void __launch_bounds__(kBucketBlock)
eval_bucket_kernel(const TableView tv, const T* __restrict__ x, T* __restrict__ y,
                   long long n, long long head, long long ntiles, int chunk) {
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned char* base = table_base<T, kSmem>(tv, smem);
    Lookup<T> lk;
    lk.init(tv, base);
    constexpr int R = rec_stride(BASIS, ORDER);
    const T* records = reinterpret_cast<const T*>(base + tv.off_records);
    const T inv_order = (T)tv.inv_order;
    const int L = tv.n_leaves;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = smem + ((tv.image_bytes + 15) & ~15) + warp * bucket_warp_bytes<T>();
    T* xs = reinterpret_cast<T*>(wbase);                                  // points grouped by leaf
    T* yb = xs + 256;                                                     // results, original order
    unsigned char* org = reinterpret_cast<unsigned char*>(yb + 256);      // original slot of xs[k]
    int* cnt = reinterpret_cast<int*>(org + 256);                         // [16] points per leaf
    int* start = cnt + 16;                                                // [16] first xs slot per leaf
    int* cbase = start + 16;                                              // [16] first chunk per leaf

    constexpr int V = kBucketV;
    constexpr int VN = Vec<T>::N;                  // points per 16-byte vector
    constexpr int NV = V / VN;                     // vectors per lane per tile (fp64: 4, fp32: 2)
    using VT = typename Vec<T>::type;
    const VT* __restrict__ xv = reinterpret_cast<const VT*>(x + head);
    VT* __restrict__ yv = reinterpret_cast<VT*>(y + head);
    const long long gwarp = (long long)blockIdx.x * (kBucketBlock / 32) + warp;
    const long long nwarps = (long long)gridDim.x * (kBucketBlock / 32);

    for (long long tile = gwarp; tile < ntiles; tile += nwarps) {
        const long long vbase = tile * (32 * NV);                         // in vectors
        // 1. this lane's V points: vector j covers tile slots (j*32 + lane)*VN .. +VN-1
        T px[V];
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const VT v = ld_stream(xv + vbase + j * 32 + lane);
            if constexpr (VN == 2) { px[2 * j] = v.x; px[2 * j + 1] = v.y; }
            else { px[4 * j] = v.x; px[4 * j + 1] = v.y; px[4 * j + 2] = v.z; px[4 * j + 3] = v.w; }
        }
        int li[V];
        lk.template leaves<V>(px, li);

        // (every lane must execute the shuffle: no short-circuit in front of it)
        const int leaf0 = __shfl_sync(0xffffffffu, li[0], 0);
        bool same = true;
#pragma unroll
        for (int s = 0; s < V; ++s) same = same & (li[s] == leaf0);
        if (__all_sync(0xffffffffu, same)) {
            // whole tile in one leaf: no regrouping
            Record<BASIS, ORDER, T> rec;
            rec.template load<1>(records + li[0] * R);
#pragma unroll
            for (int s = 0; s < V; ++s) px[s] = eval_record<BASIS, ORDER, T>(rec, px[s], inv_order);
        } else {
            // 2. rank every point inside its leaf (warp-private counters)
            if (lane < 16) cnt[lane] = 0;
            __syncwarp();
            int rk[V];
#pragma unroll
            for (int s = 0; s < V; ++s) rk[s] = atomicAdd(&cnt[li[s]], 1);
            __syncwarp();
            // 3. per-leaf start slot and first chunk (exclusive scans over <= 16 lanes)
            {
                const int nl = (lane < L) ? cnt[lane] : 0;
                const int nc = (nl + chunk - 1) / chunk;
                int a = nl, c = nc;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const int ua = __shfl_up_sync(0xffffffffu, a, o), uc = __shfl_up_sync(0xffffffffu, c, o);
                    if (lane >= o) { a += ua; c += uc; }
                }
                if (lane < 16) { start[lane] = a - nl; cbase[lane] = c - nc; }
            }
            __syncwarp();
            // 4. scatter the points into leaf order
#pragma unroll
            for (int s = 0; s < V; ++s) {
                const int k = start[li[s]] + rk[s];
                const int q = ((s / VN) * 32 + lane) * VN + (s % VN);     // original tile slot
                xs[k] = px[s];
                org[k] = (unsigned char)q;
            }
            __syncwarp();
            // 5. this lane's chunk: leaf = last one whose first chunk <= lane
            int leaf = 0;
            for (int l = 1; l < L; ++l) leaf += (cbase[l] <= lane) ? 1 : 0;
            const int nleaf = cnt[leaf];
            const int off = (lane - cbase[leaf]) * chunk;
            const int m = min(chunk, nleaf - off);                        // <= 0: idle lane
            if (m > 0) {
                Record<BASIS, ORDER, T> rec;
                rec.template load<1>(records + leaf * R);
                const int k0 = start[leaf] + off;
                for (int j = 0; j < m; ++j) {
                    const T v = eval_record<BASIS, ORDER, T>(rec, xs[k0 + j], inv_order);
                    yb[org[k0 + j]] = v;
                }
            }
            __syncwarp();
            // 6. results back in original order
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const VT v = reinterpret_cast<const VT*>(yb)[j * 32 + lane];
                if constexpr (VN == 2) { px[2 * j] = v.x; px[2 * j + 1] = v.y; }
                else { px[4 * j] = v.x; px[4 * j + 1] = v.y; px[4 * j + 2] = v.z; px[4 * j + 3] = v.w; }
            }
            __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            VT v;
            if constexpr (VN == 2) { v.x = px[2 * j]; v.y = px[2 * j + 1]; }
            else { v.x = px[4 * j]; v.y = px[4 * j + 1]; v.z = px[4 * j + 2]; v.w = px[4 * j + 3]; }
            st_stream(yv + vbase + j * 32 + lane, v);
        }
    }
    // everything outside the whole tiles: scalar
    auto f = [&](T xv) -> T {
        Record<BASIS, ORDER, T> r;
        r.template load<1>(records + lk.leaf(xv) * R);
        return eval_record<BASIS, ORDER, T>(r, xv, inv_order);
    };
    const long long gtid = (long long)blockIdx.x * kBucketBlock + threadIdx.x;
    const long long gstride = (long long)gridDim.x * kBucketBlock;
    for (long long k = gtid; k < head; k += gstride) y[k] = f(x[k]);
    for (long long k = head + ntiles * (32 * V) + gtid; k < n; k += gstride) y[k] = f(x[k]);
}

}  // namespace ai
