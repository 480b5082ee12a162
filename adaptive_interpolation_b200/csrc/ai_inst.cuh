// ai_inst.cuh -- explicit instantiation of eval_kernel / eval_sum_kernel / eval_multi_kernel for
// ONE (basis, dtype) and a range of orders, plus the order switch.  Included by the twelve
// ai_inst_*.cu units (orders 0..kSplitOrder and kSplitOrder+1..16 per basis and dtype) so that
// they compile in parallel.
#pragma once
#include "ai_kernels.cuh"

#include <mutex>

namespace ai {
std::mutex& smem_attr_mutex();          // ai_api.cu: serialises the get-then-set below across threads
namespace {

constexpr int kMaxOrder = 16;

// Opt a kernel into more than 48 KB of dynamic shared memory.  The limit is per kernel FUNCTION, not
// per table: only ever raise it, so a later, smaller table of the same (basis, order, dtype, mode)
// cannot take the room of an earlier, larger one away.
template <typename K>
cudaError_t raise_smem_limit(K kernel, int smem_bytes) {
    if (smem_bytes <= 48 * 1024) return cudaSuccess;
    std::lock_guard<std::mutex> lock(smem_attr_mutex());
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    if (fa.maxDynamicSharedSizeBytes >= smem_bytes) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
}

template <int BASIS, typename T, int ORDER, int MODE>
struct Unit {
    static cudaError_t eval(const LaunchCfg& c, const TableView& tv, const T* x, T* y,
                            long long n, long long head, long long nvec) {
        SelectTable<T, MODE> sel;
        if constexpr (MODE == kSelect) {
            if (!c.select) return cudaErrorInvalidValue;
            sel.mid = (T)c.select[0];
            for (int k = 0; k < kMaxRecord; ++k) {
                sel.rec[0][k] = (T)c.select[1 + k];
                sel.rec[1][k] = (T)c.select[1 + kMaxRecord + k];
            }
        }
        eval_kernel<BASIS, ORDER, T, MODE><<<c.grid, kBlock, c.smem_bytes, c.stream>>>(tv, sel, x, y, n, head, nvec);
        return cudaGetLastError();
    }
    static cudaError_t sum(const LaunchCfg& c, const TableView& tv, const T* x, double* r, long long n) {
        if constexpr (mode_bank(MODE) || MODE == kSelect) {
            return cudaErrorInvalidValue;
        } else {
            eval_sum_kernel<BASIS, ORDER, T, MODE><<<c.grid, kBlock, c.smem_bytes, c.stream>>>(tv, x, r, n);
            return cudaGetLastError();
        }
    }
    static cudaError_t multi(const LaunchCfg& c, const MultiArgs& a, const T* x, long long n, long long head,
                             long long nvec) {
        constexpr int HM = mode_packed_hdr(MODE) ? 1 : (mode_uniform_hdr(MODE) ? 2 : 0);
        cudaError_t e = raise_smem_limit(eval_multi_kernel<BASIS, ORDER, T, HM>, c.smem_bytes);
        if (e != cudaSuccess) return e;
        eval_multi_kernel<BASIS, ORDER, T, HM><<<c.grid, kBlock, c.smem_bytes, c.stream>>>(a, x, n, head, nvec);
        return cudaGetLastError();
    }
    static cudaError_t prepare(int smem_bytes, int* blocks_per_sm) {
        cudaError_t e = raise_smem_limit(eval_kernel<BASIS, ORDER, T, MODE>, smem_bytes);
        if (e != cudaSuccess) return e;
        if constexpr (MODE == kSmem || MODE == kSmemP || MODE == kSmemU || MODE == kGlobal || MODE == kGlobalC) {
            e = raise_smem_limit(eval_sum_kernel<BASIS, ORDER, T, MODE>, smem_bytes);
            if (e != cudaSuccess) return e;
        }
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, eval_kernel<BASIS, ORDER, T, MODE>,
                                                             kBlock, smem_bytes);
    }
};

// compile-time unrolled switch over (order, mode)
template <int BASIS, typename T, int ORDER, int HI>
struct Switch {
    template <typename... A> static cudaError_t eval(int order, int mode, A&&... a) {
        if (order == ORDER) {
            if (mode == kSmem) return Unit<BASIS, T, ORDER, kSmem>::eval(a...);
            if (mode == kSmemBank) return Unit<BASIS, T, ORDER, kSmemBank>::eval(a...);
            if (mode == kGlobal) return Unit<BASIS, T, ORDER, kGlobal>::eval(a...);
            if (mode == kSmemBankN) return Unit<BASIS, T, ORDER, kSmemBankN>::eval(a...);
            if (mode == kSelect) return Unit<BASIS, T, ORDER, kSelect>::eval(a...);
            if (mode == kSmemV) return Unit<BASIS, T, ORDER, kSmemV>::eval(a...);
            if (mode == kGlobalC) return Unit<BASIS, T, ORDER, kGlobalC>::eval(a...);
            if (mode == kSmemVC) return Unit<BASIS, T, ORDER, kSmemVC>::eval(a...);
            if constexpr (BASIS != kMono) {          // the monomial record has no header to pack
                if (mode == kSmemP) return Unit<BASIS, T, ORDER, kSmemP>::eval(a...);
                if (mode == kSmemBankP) return Unit<BASIS, T, ORDER, kSmemBankP>::eval(a...);
                if (mode == kSmemU) return Unit<BASIS, T, ORDER, kSmemU>::eval(a...);
                if constexpr (sizeof(T) == 8) {      // fp32 tables beyond 32 cells keep first[] + header
                    if (mode == kSmemBankU) return Unit<BASIS, T, ORDER, kSmemBankU>::eval(a...);
                }
            }
            return cudaErrorInvalidValue;
        }
        if constexpr (ORDER < HI) return Switch<BASIS, T, ORDER + 1, HI>::eval(order, mode, a...);
        else return cudaErrorInvalidValue;
    }
    template <typename... A> static cudaError_t sum(int order, int mode, A&&... a) {
        if (order == ORDER) {
            if (mode == kGlobal) return Unit<BASIS, T, ORDER, kGlobal>::sum(a...);
            if (mode == kGlobalC) return Unit<BASIS, T, ORDER, kGlobalC>::sum(a...);
            if constexpr (BASIS != kMono) {
                if (mode == kSmemP) return Unit<BASIS, T, ORDER, kSmemP>::sum(a...);
                if (mode == kSmemU) return Unit<BASIS, T, ORDER, kSmemU>::sum(a...);
            }
            return Unit<BASIS, T, ORDER, kSmem>::sum(a...);
        }
        if constexpr (ORDER < HI) return Switch<BASIS, T, ORDER + 1, HI>::sum(order, mode, a...);
        else return cudaErrorInvalidValue;
    }
    template <typename... A> static cudaError_t multi(int order, int packed, A&&... a) {
        if (order == ORDER) {
            if constexpr (BASIS != kMono) {          // packed: the tables' common header form (0, 1, 2)
                if (packed == 1) return Unit<BASIS, T, ORDER, kSmemP>::multi(a...);
                if (packed == 2) return Unit<BASIS, T, ORDER, kSmemU>::multi(a...);
            }
            return packed ? cudaErrorInvalidValue : Unit<BASIS, T, ORDER, kSmem>::multi(a...);
        }
        if constexpr (ORDER < HI) return Switch<BASIS, T, ORDER + 1, HI>::multi(order, packed, a...);
        else return cudaErrorInvalidValue;
    }
    static cudaError_t prepare(int order, int mode, int smem_bytes, int* bps) {
        if (order == ORDER) {
            if (mode == kSmem) return Unit<BASIS, T, ORDER, kSmem>::prepare(smem_bytes, bps);
            if (mode == kSmemBank) return Unit<BASIS, T, ORDER, kSmemBank>::prepare(smem_bytes, bps);
            if (mode == kGlobal) return Unit<BASIS, T, ORDER, kGlobal>::prepare(smem_bytes, bps);
            if (mode == kSmemBankN) return Unit<BASIS, T, ORDER, kSmemBankN>::prepare(smem_bytes, bps);
            if (mode == kSelect) return Unit<BASIS, T, ORDER, kSelect>::prepare(smem_bytes, bps);
            if (mode == kSmemV) return Unit<BASIS, T, ORDER, kSmemV>::prepare(smem_bytes, bps);
            if (mode == kGlobalC) return Unit<BASIS, T, ORDER, kGlobalC>::prepare(smem_bytes, bps);
            if (mode == kSmemVC) return Unit<BASIS, T, ORDER, kSmemVC>::prepare(smem_bytes, bps);
            if constexpr (BASIS != kMono) {
                if (mode == kSmemP) return Unit<BASIS, T, ORDER, kSmemP>::prepare(smem_bytes, bps);
                if (mode == kSmemBankP) return Unit<BASIS, T, ORDER, kSmemBankP>::prepare(smem_bytes, bps);
                if (mode == kSmemU) return Unit<BASIS, T, ORDER, kSmemU>::prepare(smem_bytes, bps);
                if constexpr (sizeof(T) == 8) {
                    if (mode == kSmemBankU) return Unit<BASIS, T, ORDER, kSmemBankU>::prepare(smem_bytes, bps);
                }
            }
            return cudaErrorInvalidValue;
        }
        if constexpr (ORDER < HI) return Switch<BASIS, T, ORDER + 1, HI>::prepare(order, mode, smem_bytes, bps);
        else return cudaErrorInvalidValue;
    }
};

}  // namespace
}  // namespace ai

#define AI_DEFINE_UNIT(NAME, BASIS, T, LO, HI)                                                    \
    namespace ai {                                                                                \
    static_assert(HI <= kMaxOrder, "order range");                                                \
    cudaError_t launch_eval_##NAME(int order, int mode, const LaunchCfg& c, const TableView& tv,  \
                                   const T* x, T* y, long long n, long long head, long long nvec) { \
        return Switch<BASIS, T, LO, HI>::eval(order, mode, c, tv, x, y, n, head, nvec);           \
    }                                                                                             \
    cudaError_t launch_sum_##NAME(int order, int mode, const LaunchCfg& c, const TableView& tv,   \
                                  const T* x, double* r, long long n) {                           \
        return Switch<BASIS, T, LO, HI>::sum(order, mode, c, tv, x, r, n);                        \
    }                                                                                             \
    cudaError_t launch_multi_##NAME(int order, int packed, const LaunchCfg& c, const MultiArgs& a, \
                                    const T* x, long long n, long long head, long long nvec) {    \
        return Switch<BASIS, T, LO, HI>::multi(order, packed, c, a, x, n, head, nvec);            \
    }                                                                                             \
    cudaError_t prepare_##NAME(int order, int mode, int smem_bytes, int* bps) {                   \
        return Switch<BASIS, T, LO, HI>::prepare(order, mode, smem_bytes, bps);                   \
    }                                                                                             \
    }
