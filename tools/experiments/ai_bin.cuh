// ai_bin.cuh -- binned evaluation for high-order fp64 tables with at most eight records.
//
// Why: a random-point evaluation gathers one whole record per point through the 128 B/clk
// shared-memory return path; at fp64 order 13 that is 14 eight-byte words = 28 wavefronts per 32
// points against a budget of 22.8 clocks per 32 points at the HBM roofline, so eval_kernel stops at
// ~66% of HBM on BASELINE configs[2] however it is scheduled.  The only way
// under that ceiling is to move the 8-byte POINT to where a record already sits in registers
// instead of the 112-byte record to the point.
//
// How (per CTA tile of kBinThreads * NE points, no atomics):
//   1. every thread finds the record index of its NE points (same lookup as eval_kernel) and
//      counts them per record in a packed register histogram (8 x 4 bits), remembering each
//      point's rank among its own same-record points;
//   2. a warp scan (SHFL) of the histograms widened to 16-bit fields gives every lane its
//      exclusive offset per record; one warp turns the per-warp totals into per-warp bases, with
//      every record's bucket padded to a multiple of Q slots so that no evaluator thread straddles
//      two buckets;
//   3. x is scattered to slot = base[record] + lane offset + own rank in shared memory;
//   4. evaluator thread j owns slots [jQ, jQ+Q): ONE record load (registers), Q evaluations with
//      the same eval_record as eval_kernel (identical bits), results written back in place; Q is
//      odd, so the stride-Q reads and writes are bank-conflict free;
//   5. the owners gather their y from the slots they scattered to and store coalesced vectors.
// Shared-memory wavefronts per 32 points: ~2 (scan) + ~5 (scatter) + 2 + 2.5 (records) + 2 + ~5
// (gather) = 18-19 instead of 29-32.
//
// A tile whose threads each hold points of ONE record (sorted / clustered input) skips all of
// this and runs the register-resident fast path of eval_kernel.
#pragma once
#include "ai_kernels.cuh"

namespace ai {

#ifndef AI_BIN_THREADS
#define AI_BIN_THREADS 256
#endif
#ifndef AI_BIN_UNROLL
#define AI_BIN_UNROLL 5            // 16-byte vectors per thread per tile (10 points)
#endif
#ifndef AI_BIN_Q
#define AI_BIN_Q 11                // slots per evaluator thread; odd: conflict-free stride
#endif
#ifndef AI_BIN_MIN_BLOCKS
#define AI_BIN_MIN_BLOCKS 2
#endif
#ifndef AI_BIN_PREFETCH
#define AI_BIN_PREFETCH 1          // 1: tile k+1's x is loaded into registers while tile k is processed
#endif
constexpr int kBinThreads = AI_BIN_THREADS;
constexpr int kBinUnroll = AI_BIN_UNROLL;
constexpr int kBinQ = AI_BIN_Q;
constexpr int kBinMaxRecords = 8;
constexpr int kBinWarps = kBinThreads / 32;
constexpr int kBinNE = kBinUnroll * 2;                              // points per thread per tile
constexpr int kBinTilePoints = kBinThreads * kBinNE;
constexpr int kBinSlots = kBinThreads * kBinQ;                      // evaluator capacity
static_assert(kBinQ % 2 == 1, "Q must be odd (bank-conflict-free evaluator stride)");
static_assert(kBinNE <= 15, "per-thread counts live in 4-bit fields");
static_assert(kBinSlots >= kBinTilePoints + kBinMaxRecords * (kBinQ - 1), "evaluator capacity < tile + padding");
static_assert(kBinTilePoints + kBinMaxRecords * kBinQ < 65536, "slot numbers live in 16-bit fields");

// bytes of dynamic shared memory: [table image][slots][per-warp totals][per-warp bases][bucket starts]
__host__ __device__ constexpr int bin_smem_bytes(int image_bytes) {
    return image_bytes + kBinSlots * 8 + kBinWarps * 16 + kBinWarps * 16 + 16;
}

// ceil(t / Q) for 0 <= t <= tile points (checked at compile time below)
__host__ __device__ constexpr unsigned bin_ceil_div_q(unsigned t) { return ((t + kBinQ - 1) * (65536u / kBinQ + 1u)) >> 16; }
constexpr bool bin_div_ok() {
    for (unsigned t = 0; t <= (unsigned)kBinTilePoints; ++t)
        if (bin_ceil_div_q(t) != (t + kBinQ - 1) / kBinQ) return false;
    return true;
}
static_assert(bin_div_ok(), "ceil-division magic number invalid for this Q / tile");

template <int BASIS, int ORDER, int MODE>
__global__ void __launch_bounds__(kBinThreads, AI_BIN_MIN_BLOCKS)
eval_bin_kernel(const TableView tv, const double* __restrict__ x, double* __restrict__ y,
                long long n, long long head, long long nvec) {
    static_assert(MODE == kSmem || MODE == kSmemP || MODE == kSmemU, "packed shared-memory layouts only");
    using T = double;
    extern __shared__ __align__(16) unsigned char smem[];
    // stage the packed image (as table_base<T, kSmem>)
    {
        const int4* src = reinterpret_cast<const int4*>(tv.image);
        int4* dst = reinterpret_cast<int4*>(smem);
        for (int k = threadIdx.x; k < tv.image_bytes / 16; k += blockDim.x) dst[k] = __ldg(src + k);
    }
    T* const slots = reinterpret_cast<T*>(smem + tv.image_bytes);
    uint4* const wtot = reinterpret_cast<uint4*>(smem + tv.image_bytes + kBinSlots * 8);
    uint4* const wbase = wtot + kBinWarps;
    unsigned* const startq = reinterpret_cast<unsigned*>(wbase + kBinWarps);     // 4 words: starts / Q as bytes, evaluator count
    __syncthreads();

    Lookup<T> lk;
    lk.init(tv, smem);
    constexpr bool PK = mode_packed_hdr(MODE);
    constexpr int R = rec_stride_for<BASIS, ORDER, T, MODE>();
    const T* const records = reinterpret_cast<const T*>(smem + tv.off_records);
    const T inv_order = (T)tv.inv_order;
    auto load_record = [&](Record<BASIS, ORDER, T>& r, int idx) {
        AI_CHECK(idx >= 0 && idx < tv.n_records);
        const T* p = records + idx * R;
        if constexpr (PK) r.template load_packed<1>(p, tv.s0_hi, tv.s0_lo);
        else if constexpr (mode_uniform_hdr(MODE)) r.template load_uniform<1>(p, idx, (T)tv.u_w, (T)tv.u_lb, (T)tv.u_s);
        else r.template load<1>(p);
    };
    auto f = [&](T xv) -> T {
        Record<BASIS, ORDER, T> r;
        load_record(r, lk.leaf(xv));
        return eval_record<BASIS, ORDER, T>(r, xv, inv_order);
    };

    using V = double2;
    constexpr int U = kBinUnroll, NE = kBinNE, Q = kBinQ;
    constexpr long long kTile = (long long)kBinThreads * U;        // vectors per tile
    const V* __restrict__ xv = reinterpret_cast<const V*>(x + head);
    V* __restrict__ yv = reinterpret_cast<V*>(y + head);
    const long long full_tiles = nvec / kTile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    auto load_tile = [&](V (&buf)[U], long long tile) {
        const long long base = tile * kTile + threadIdx.x;
#pragma unroll
        for (int u = 0; u < U; ++u) buf[u] = ld_stream(xv + base + (long long)u * kBinThreads);
    };

    bool expect_uniform = true;              // CTA-uniform: derived from barrier results only
    long long tile = blockIdx.x;
    V in[U];
#if AI_BIN_PREFETCH
    if (tile < full_tiles) load_tile(in, tile);
#endif
    for (; tile < full_tiles; tile += gridDim.x) {
#if AI_BIN_PREFETCH
        V nxt[U];
        const long long next = tile + gridDim.x;
        if (next < full_tiles) load_tile(nxt, next);
#else
        load_tile(in, tile);                 // (other resident CTAs cover the latency)
#endif
        const long long base = tile * kTile + threadIdx.x;
        T xs[NE];
        int li[NE];
#pragma unroll
        for (int u = 0; u < U; ++u) { xs[2 * u] = in[u].x; xs[2 * u + 1] = in[u].y; }
        lk.leaves(xs, li);
        bool same = true;
#pragma unroll
        for (int e = 1; e < NE; ++e) same = same && (li[e] == li[0]);

        bool uniform = false;
        if (expect_uniform) {
            uniform = __syncthreads_and(same);
            expect_uniform = uniform;
        }
        unsigned slot[NE];
        if (!uniform) {
            // ---- 1. per-thread histogram (8 x 4 bits) and own ranks
            unsigned h = 0;
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                AI_CHECK(li[e] >= 0 && li[e] < kBinMaxRecords);
                const unsigned sh = (unsigned)li[e] * 4u;
                slot[e] = (h >> sh) & 15u;                       // own rank, for now
                h += 1u << sh;
            }
            // ---- 2. warp scan of the counts, two 16-bit fields per register (records 2r, 2r+1)
            unsigned c[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) c[r] = ((h >> (8 * r)) & 15u) | (((h >> (8 * r + 4)) & 15u) << 16);
            unsigned inc[4] = {c[0], c[1], c[2], c[3]};
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const unsigned v = __shfl_up_sync(0xffffffffu, inc[r], d);
                    if (lane >= d) inc[r] += v;
                }
            }
            if (lane == 31) wtot[warp] = make_uint4(inc[0], inc[1], inc[2], inc[3]);
            // barrier A doubles as the "has the input become per-thread uniform again?" vote
            const bool all_same = __syncthreads_and(same);
            if (all_same) {
                uniform = true;
                expect_uniform = true;
            } else {
                // ---- warp 0, lane b = record b: bucket starts (padded to multiples of Q) and warp bases
                if (warp == 0) {
                    const int b = lane & 7;
                    unsigned run = 0;                            // points of record b in warps before w
                    unsigned cnt[kBinWarps];
#pragma unroll
                    for (int w = 0; w < kBinWarps; ++w) {
                        const unsigned word = reinterpret_cast<const unsigned*>(wtot + w)[b >> 1];
                        cnt[w] = (word >> ((b & 1) * 16)) & 0xffffu;
                    }
                    unsigned total = 0;
#pragma unroll
                    for (int w = 0; w < kBinWarps; ++w) total += cnt[w];
                    const unsigned nq = bin_ceil_div_q(total);   // evaluator threads of this bucket
                    unsigned incq = nq;                          // inclusive scan over the 8 records
#pragma unroll
                    for (int d = 1; d < 8; d <<= 1) {
                        const unsigned v = __shfl_up_sync(0xffffffffu, incq, d);
                        if (b >= d) incq += v;
                    }
                    const unsigned sq = incq - nq;               // first evaluator thread of bucket b
                    if (lane < 8) {
#pragma unroll
                        for (int w = 0; w < kBinWarps; ++w) {
                            reinterpret_cast<unsigned short*>(wbase + w)[b] = (unsigned short)(sq * Q + run);
                            run += cnt[w];
                        }
                        reinterpret_cast<unsigned char*>(startq)[b] = (unsigned char)sq;
                        if (b == 7) startq[2] = incq;            // evaluator threads in use
                    }
                }
                __syncthreads();                                 // B
                // ---- 3. scatter: slot = warp base + lane offset (exclusive) + own rank
                const uint4 wb = wbase[warp];
                const unsigned p[4] = {wb.x + inc[0] - c[0], wb.y + inc[1] - c[1], wb.z + inc[2] - c[2], wb.w + inc[3] - c[3]};
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    const unsigned b = (unsigned)li[e];
                    const unsigned lo = (b & 2u) ? p[1] : p[0], hi = (b & 2u) ? p[3] : p[2];
                    const unsigned w = (b & 4u) ? hi : lo;
                    slot[e] += (w >> ((b & 1u) * 16u)) & 0xffffu;
                    AI_CHECK(slot[e] < (unsigned)kBinSlots);
                    slots[slot[e]] = xs[e];
                }
                __syncthreads();                                 // C
                // ---- 4. evaluate Q consecutive slots with one register-resident record
                if (threadIdx.x < startq[2]) {
                    const unsigned s4a = startq[0], s4b = startq[1];
                    // bucket = #{b : start[b] <= tid} - 1 (starts are non-decreasing bytes, start[0] = 0)
                    const unsigned t4 = threadIdx.x * 0x01010101u;
                    const unsigned le = __vcmpleu4(s4a, t4), le2 = __vcmpleu4(s4b, t4);
                    const int b = (__popc(le) + __popc(le2)) / 8 - 1;
                    Record<BASIS, ORDER, T> rec;
                    load_record(rec, b);
                    T* const mine = slots + threadIdx.x * Q;
#pragma unroll
                    for (int e = 0; e < Q; ++e) mine[e] = eval_record<BASIS, ORDER, T>(rec, mine[e], inv_order);
                }
                __syncthreads();                                 // D
                // ---- 5. gather y from the slots this thread scattered to
#pragma unroll
                for (int e = 0; e < NE; ++e) xs[e] = slots[slot[e]];
            }
        }
        if (uniform) {
            Record<BASIS, ORDER, T> rec;
            load_record(rec, li[0]);
#pragma unroll
            for (int e = 0; e < NE; ++e) xs[e] = eval_record<BASIS, ORDER, T>(rec, xs[e], inv_order);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            V out;
            out.x = xs[2 * u]; out.y = xs[2 * u + 1];
            st_stream(yv + base + (long long)u * kBinThreads, out);
        }
#if AI_BIN_PREFETCH
#pragma unroll
        for (int u = 0; u < U; ++u) in[u] = nxt[u];
#endif
    }
    // remainder vectors (less than one tile), spread over the whole grid; scalar head and tail
    const long long gtid = (long long)blockIdx.x * kBinThreads + threadIdx.x;
    const long long gstride = (long long)gridDim.x * kBinThreads;
    for (long long v = full_tiles * kTile + gtid; v < nvec; v += gstride) {
        V iv = ld_stream(xv + v), out;
        out.x = f(iv.x); out.y = f(iv.y);
        st_stream(yv + v, out);
    }
    for (long long k = gtid; k < head; k += gstride) y[k] = f(x[k]);
    for (long long k = head + nvec * 2 + gtid; k < n; k += gstride) y[k] = f(x[k]);
}

}  // namespace ai
