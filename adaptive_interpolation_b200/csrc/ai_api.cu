// ai_api.cu -- the C ABI of include/ai_b200.h: table construction (host), kernel dispatch,
// the host-buffer pipeline, and the measurement helpers.
//
// Reference interfaces replaced (see include/ai_b200.h for the full mapping):
//   adaptive_interpolation/generate.py:408-502   build_code / run / run_single / run_multi
//   run_test.py:242-283                          ctypes call of the generated eval(...)
//   adaptive_interpolation/approximator.py:191-243  tree flattening + uniform-grid map
#include "../../include/ai_b200.h"
#include "ai_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

namespace ai {
// one lock for every get-then-set of a kernel's dynamic shared memory limit (ai_inst.cuh, index_entry)
std::mutex& smem_attr_mutex() {
    static std::mutex mu;
    return mu;
}
}  // namespace ai

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define AI_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return fail(AI_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                \
    } while (0)

// cudaGetDeviceProperties costs milliseconds; tables can be created per fitter node (fit.py)
struct DeviceInfo { bool known = false; int major = 0, minor = 0, sm_count = 0; size_t smem_optin = 0; };
int device_info(int device, DeviceInfo* out) {
    static std::mutex mu;
    static DeviceInfo cache[64];
    std::lock_guard<std::mutex> lock(mu);
    if (device < 0 || device >= 64) return fail(AI_ERR_INVALID, "device %d out of range", device);
    if (!cache[device].known) {
        cudaDeviceProp prop;
        cudaError_t e = cudaGetDeviceProperties(&prop, device);
        if (e != cudaSuccess) return fail(AI_ERR_CUDA, "cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
        cache[device].known = true;
        cache[device].major = prop.major; cache[device].minor = prop.minor;
        cache[device].sm_count = prop.multiProcessorCount;
        cache[device].smem_optin = (size_t)prop.sharedMemPerBlockOptin;
    }
    *out = cache[device];
    return AI_OK;
}

// No C++ exception may cross the extern "C" boundary (a huge n_leaves would otherwise end the
// process with std::bad_alloc): every entry point that allocates runs inside guarded().
template <typename F>
int guarded(F&& body) {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        return fail(AI_ERR_UNSUPPORTED, "out of host memory");
    } catch (const std::exception& e) {
        return fail(AI_ERR_INVALID, "internal error: %s", e.what());
    } catch (...) {
        return fail(AI_ERR_INVALID, "internal error");
    }
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

constexpr int kGridMaxExact = 16384;   // cells allowed when the grid alone resolves the leaf
constexpr int kGridMaxSearch = 4096;   // cells when a search follows anyway
// Tables far beyond shared memory (the reference's missing order-3 / 2.2e-14 fits have 1e4-1e5
// leaves, .MISSING_LARGE_BLOBS) read everything through L2 anyway: there the grid may grow with
// the table (one L2 access instead of a long search), up to 4 MiB of 32-bit entries.
constexpr int64_t kLargeTableLeaves = 4096;
constexpr int64_t kGridMaxLarge = 1 << 20;
constexpr int64_t kMaxLeaves = 1 << 24;

// Test hook: AI_B200_MAX_CELLS=<n> caps both budgets, forcing the search path on tables whose
// breakpoints would otherwise sit on cell edges (tests/test_gpu_parity.py).
int grid_budget(int dflt) {
    const char* e = std::getenv("AI_B200_MAX_CELLS");
    if (!e || !*e) return dflt;
    const long v = std::strtol(e, nullptr, 10);
    return (v >= 2 && v < dflt) ? (int)v : dflt;    // a domain straddling 0 needs >= 2 cells
}
constexpr int kHostSlots = 4;          // host pipeline depth (slots actually used: AI_B200_HOST_SLOTS)

// Fork-join pool of host threads for the bounce copies of the staged host pipeline (pageable caller
// buffers).  run() hands every participant (the workers and the caller) one contiguous share of every
// job and returns when ALL workers have reported for this batch, so no thread can still be touching a
// buffer of batch g when batch g + 1 is set up.
class CopyPool {
public:
    struct Job { char* dst; const char* src; size_t bytes; };
    explicit CopyPool(int participants) : n_(std::max(1, participants)) {
        for (int w = 0; w + 1 < n_; ++w) th_.emplace_back([this, w] { worker(w); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_start_.notify_all();
        for (auto& t : th_) t.join();
    }
    CopyPool(const CopyPool&) = delete;
    CopyPool& operator=(const CopyPool&) = delete;
    void run(const Job* jobs, int nj) {
        if (nj <= 0) return;
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (int j = 0; j < nj; ++j) jobs_[j] = jobs[j];
            njobs_ = nj; finished_ = 0; ++gen_;
        }
        cv_start_.notify_all();
        share(jobs, nj, n_ - 1, n_);
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return finished_ == (int)th_.size(); });
    }

private:
    static void share(const Job* jobs, int nj, int who, int n) {
        for (int j = 0; j < nj; ++j) {
            const size_t per = ((jobs[j].bytes + (size_t)n - 1) / (size_t)n + 4095) & ~(size_t)4095;   // page-aligned shares
            const size_t lo = std::min(jobs[j].bytes, per * (size_t)who), hi = std::min(jobs[j].bytes, lo + per);
            if (hi > lo) stream_copy(jobs[j].dst + lo, jobs[j].src + lo, hi - lo);
        }
    }
    // memcpy with non-temporal stores: the destination (a bounce buffer the DMA engine reads next, or
    // the caller's y) is not read again by this core, so the write-allocate traffic of ordinary stores
    // is wasted host memory bandwidth -- and host memory bandwidth is what bounds this path
    static void stream_copy(char* dst, const char* src, size_t bytes) {
#if defined(__x86_64__)
        const size_t head = std::min(bytes, (size_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15));
        if (head) { std::memcpy(dst, src, head); dst += head; src += head; bytes -= head; }
        const size_t blocks = bytes / 64;
        for (size_t b = 0; b < blocks; ++b) {
            const __m128i v0 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 0);
            const __m128i v1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 1);
            const __m128i v2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 2);
            const __m128i v3 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 3);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 0, v0);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 1, v1);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 2, v2);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 3, v3);
            src += 64; dst += 64;
        }
        _mm_sfence();
        if (bytes & 63) std::memcpy(dst, src, bytes & 63);
#else
        std::memcpy(dst, src, bytes);
#endif
    }
    void worker(int who) {
        uint64_t seen = 0;
        for (;;) {
            Job local[2];
            int nj = 0;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_start_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_; nj = njobs_;
                for (int j = 0; j < nj; ++j) local[j] = jobs_[j];
            }
            share(local, nj, who, n_);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (++finished_ == (int)th_.size()) cv_done_.notify_one();
            }
        }
    }
    const int n_;
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_start_, cv_done_;
    uint64_t gen_ = 0;
    int finished_ = 0, njobs_ = 0;
    bool stop_ = false;
    Job jobs_[2] = {};
};

struct HostPipe {
    bool ready = false;
    int64_t chunk = 0;                 // elements per chunk
    int slots = 3;
    void* dx[kHostSlots] = {nullptr, nullptr, nullptr, nullptr};
    void* dy[kHostSlots] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaEvent_t in_done[kHostSlots] = {}, k_done[kHostSlots] = {}, out_done[kHostSlots] = {};
    cudaEvent_t k_start[kHostSlots] = {};
    // staged path (pageable caller buffers): pinned bounce buffers per slot and the copy threads
    void* hin[kHostSlots] = {nullptr, nullptr, nullptr, nullptr};
    void* hout[kHostSlots] = {nullptr, nullptr, nullptr, nullptr};
    size_t elem = 0;
    std::unique_ptr<CopyPool> pool;
    std::mutex mu;
};

}  // namespace

struct ai_table {
    int basis = 0, order = 0, dtype = 0, device = 0;
    int64_t L = 0;
    int G = 0, p = 0, K = 0;
    long long g0 = 0;
    int rstride = 0;
    std::vector<double> breaks;        // L+1, exact widening of the table dtype
    std::vector<double> coef;          // L*(order+1)
    void* d_image = nullptr;
    ai::TableView view{};
    int blocks_per_sm = 1, sm_count = 1;
    int smem_resident = 1;
    int rank_mode = 0;
    int mode = ai::kSmem;              // table residency: kSmem / kSmemBank / kGlobal / kSmemBankN
    int copies = 1;                    // bank modes: copies per element
    int packed = 0;                    // 1: records carry the one-element header (modes kSmemP / kSmemBankP)
    int cell_records = 0;              // 1: one record per coarse cell (view.rank_mode 2)
    int uniform_hdr = 0;               // 1: records carry no header, computed from the cell (modes kSmemU / kSmemBankU)
    int rank16 = 0;                    // 1: the first[] section holds rank words (view.rank_mode bit 2)
    int compact = 0;                   // 1: kGlobalC image (headerless 32-byte records per coarse cell)
    int path = 0;                      // 1: path lookup (view.rank_mode bit 4)
    int eval_smem = 0;                 // dynamic shared memory of the eval kernel in this mode
    int aux_smem = 0;                  // ... of the index / sum kernels (packed layouts only)
    std::vector<double> select;        // kSelect: [mid, rec0[kMaxRecord], rec1[kMaxRecord]] (exact widening)
    HostPipe pipe;
};

namespace {

// Build the exact power-of-two grid for sorted breakpoints b[0..L].
//   cell(x) = floor(x * 2^p);  first[g] = leaf holding the left edge of cell g0+g.
// K = 0 when every interior breakpoint is a cell edge (then first[] alone is the answer).
struct GridPlan {
    int p = 0, K = 0, G = 0;
    long long g0 = 0;
    std::vector<uint32_t> first;
    // rank mode (see Lookup::rank_guess): coarse cell = fine cell / q, leaf = rank in cont
    int rank_mode = 0;
    unsigned div_mul = 0, div_shift = 0, div_add = 0;
    unsigned cont = 0;
    int coarse_ok = 0;                   // div_* map a fine cell to its coarse cell, coarse_leaf is valid
    std::vector<uint32_t> coarse_leaf;   // leaf of every coarse cell
    std::vector<uint32_t> rank16;        // Lookup::rank16_guess words (empty: not available)
    // wide division (kGlobalC, grids of up to 2^20 fine cells): coarse = ((fine - g0) * wide_mul) >> wide_shift (64-bit)
    int wide_ok = 0;
    unsigned wide_mul = 0, wide_shift = 0;
    std::vector<uint32_t> wide_coarse_leaf;
    // path lookup (Lookup::path_leaf): replaces grid + search for trees that are a path below level Lc
    int path_ok = 0;
    int path_p = 0, path_D = 0, path_Lc = 0;          // scale 2^p; V has D bits; 2^Lc rows
    long long path_u0 = 0, path_u1 = 0;               // first / last fine cell (absolute)
    unsigned long long path_mul = 0;                  // 0: q = 1
    unsigned path_shift = 0;
    std::vector<unsigned long long> path_target;      // per row
    std::vector<uint32_t> path_rows;                  // 2^Lc rows of D - Lc + 1 leaf ordinals
};

long long gcd_ll(long long a, long long b) {
    while (b) { const long long t = a % b; a = b; b = t; }
    return a < 0 ? -a : a;
}

// Coarse cells: when every breakpoint sits on a multiple of q fine cells (q = gcd of the breakpoint
// offsets), coarse cell = fine cell / q by an exact multiply-shift.  Two uses, both verified
// exhaustively over the G fine cells (on any mismatch they stay off):
//   * records stored once per coarse cell (coarse_ok): the cell index addresses the record;
//   * with <= 32 coarse cells, the leaf as the rank of the cell in a bitmap (rank_mode).
void plan_rank(const std::vector<double>& b, GridPlan* gp) {
    if (gp->K != 0 || gp->p < 0 || gp->path_ok) return;
    const int64_t L = (int64_t)b.size() - 1;
    const int G = gp->G;
    long long q = 0;
    for (int64_t k = 1; k < L; ++k) {
        const long long off = (long long)std::ldexp(b[k], gp->p) - gp->g0;     // exact integer
        q = gcd_ll(q, off);
    }
    if (q == 0) q = G;                                   // single leaf
    const long long Gc = (G + q - 1) / q;
    if (Gc > (1 << 21)) return;
    // 64-bit product form for large grids: floor(v / q) = (v * M) >> s, checked for every fine cell
    for (int sft = 0; sft <= 40 && !gp->wide_ok; ++sft) {
        const unsigned long long M = ((1ull << sft) + (unsigned long long)q - 1) / (unsigned long long)q;
        if (M == 0 || M >= (1ull << 32)) continue;
        bool ok = true;
        for (int v = 0; v < G && ok; ++v)
            ok = (unsigned)(((unsigned long long)(unsigned)v * M) >> sft) == (unsigned)(v / q);
        if (ok) { gp->wide_ok = 1; gp->wide_mul = (unsigned)M; gp->wide_shift = (unsigned)sft; }
    }
    if (gp->wide_ok) {
        gp->wide_coarse_leaf.resize((size_t)Gc);
        for (long long c = 0; c < Gc; ++c) gp->wide_coarse_leaf[(size_t)c] = gp->first[(size_t)(c * q)];
        for (int g = 0; g < G && gp->wide_ok; ++g)
            if (gp->wide_coarse_leaf[(size_t)(g / q)] != gp->first[g]) gp->wide_ok = 0;     // a breakpoint off the coarse grid
    }
    unsigned mul = 0, sh = 0;
    bool found = false;
    for (int S = 31; S >= 0 && !found; --S) {
        const unsigned long long M = ((1ull << S) + (unsigned long long)q - 1) / (unsigned long long)q;
        if (M == 0 || M >= (1ull << 32) || (unsigned long long)(G - 1) * M >= (1ull << 32)) continue;
        if ((unsigned long long)(G - 1) * M + ((1ull << S) - 1) >= (1ull << 32)) continue;
        bool ok = true;
        for (int v = 0; v < G && ok; ++v) ok = (((unsigned)v * (unsigned)M) >> S) == (unsigned)(v / q);
        if (ok) { mul = (unsigned)M; sh = (unsigned)S; found = true; }
    }
    if (!found) return;
    const unsigned add = (unsigned)(0u - (unsigned)gp->g0 * mul);      // -g0 * mul  (mod 2^32)
    std::vector<uint32_t> coarse_leaf((size_t)Gc);
    for (long long c = 0; c < Gc; ++c) coarse_leaf[(size_t)c] = gp->first[(size_t)(c * q)];
    for (int g = 0; g < G; ++g) {                        // exhaustive check of the device formula
        const unsigned gabs = (unsigned)(gp->g0 + g);    // what Lookup::cell_abs returns
        const unsigned c = (gabs * mul + add) >> sh;
        if (c >= (unsigned)Gc || coarse_leaf[c] != gp->first[g]) return;
    }
    gp->coarse_ok = 1; gp->div_mul = mul; gp->div_shift = sh; gp->div_add = add;
    gp->coarse_leaf.swap(coarse_leaf);
    if (Gc > 32 && Gc <= 1024) {
        // rank words: (continuation bits in all earlier words) << 16 | 16 continuation bits
        std::vector<uint32_t> words((size_t)((Gc + 15) / 16), 0);
        unsigned before = 0;
        for (size_t w = 0; w < words.size(); ++w) {
            unsigned bits = 0;
            for (int j = 0; j < 16; ++j) {
                const long long c = (long long)w * 16 + j;
                if (c >= 1 && c < Gc && gp->coarse_leaf[(size_t)c] == gp->coarse_leaf[(size_t)(c - 1)]) bits |= (1u << j);
            }
            words[w] = (before << 16) | bits;
            before += (unsigned)__builtin_popcount(bits);
        }
        bool ok = before < 65536u;
        for (int g = 0; g < G && ok; ++g) {              // the device formula, for every fine cell
            const unsigned gabs = (unsigned)(gp->g0 + g);
            const unsigned c = (gabs * mul + add) >> sh;
            const unsigned w = words[c >> 4];
            const unsigned upto = (2u << (c & 15u)) - 1u;
            const int leaf = (int)c - (int)(w >> 16) - __builtin_popcount(w & upto & 0xffffu);
            ok = leaf == (int)gp->first[g];
        }
        if (ok) gp->rank16.swap(words);
    }
    if (Gc > 32) return;
    unsigned cont = 0;
    for (long long c = 1; c < Gc; ++c)
        if (gp->coarse_leaf[(size_t)c] == gp->coarse_leaf[(size_t)(c - 1)]) cont |= (1u << c);
    for (int g = 0; g < G; ++g) {
        const unsigned gabs = (unsigned)(gp->g0 + g);
        const unsigned c = (gabs * mul + add) >> sh;
        const unsigned below = (c >= 31) ? ~0u : ((2u << c) - 1u);
        const int leaf = (int)c - __builtin_popcount(cont & below);
        if (leaf != (int)gp->first[g]) return;
    }
    gp->rank_mode = 1; gp->cont = cont;
}

bool cells_for(const std::vector<double>& b, int p, long long* g0, long long* glast) {
    const double lo = std::ldexp(b.front(), p), hi = std::ldexp(b.back(), p);
    if (!(std::fabs(lo) < 9.0e15) || !(std::fabs(hi) < 9.0e15)) return false;
    *g0 = (long long)std::floor(lo);
    long long gl = (long long)std::ceil(hi) - 1;     // last cell holding a point < ub
    if (gl < *g0) gl = *g0;
    *glast = gl;
    return true;
}

// The PATH lookup (ai_kernels.cuh, Lookup::path_leaf) for tables whose exact grid is far too large
// (deep bisection towards a discontinuity: BASELINE config 4 has leaf widths from 1.25 down to
// 10 * 2^-43).  Conditions, all checked here on the exact breakpoints:
//   * every breakpoint (both ends included) is an integer after scaling by 2^p, p <= 60, and the
//     fine cell ids stay below 2^31 (fp32) / 2^52 (fp64);
//   * the domain is W = q * 2^D fine cells wide (q odd) and every breakpoint offset is a multiple of
//     q, so V = offset / q is a D-bit position; every leaf is a dyadic node of that D-level trie;
//   * below some level Lc <= 6 every node of the trie is a PATH: at each split at least one child
//     is a leaf.  Then, with `target` any position inside the deepest leaf of the row's path,
//     the first bit in which V differs from target names the sibling leaf that holds V.
// The device formula (multiply-high division included) is then replayed on both sides of every
// breakpoint; the division is monotone, so that proves it for every fine cell.
constexpr int kPathMaxRowsLog2 = 6;
void plan_path(const std::vector<double>& b, int dtype, GridPlan* gp) {
    if (std::getenv("AI_B200_NO_PATH")) return;
    const int64_t L = (int64_t)b.size() - 1;
    if (L < 3 || L > 65535) return;
    const double lim = (dtype == AI_DTYPE_F32) ? 2147483648.0 : 4503599627370496.0;
    int p = -1;
    for (int q = 0; q <= 60 && p < 0; ++q) {
        bool all = true;
        for (int64_t k = 0; k <= L && all; ++k) {
            const double v = std::ldexp(b[k], q);
            all = (std::fabs(v) < lim) && (v == std::floor(v));
        }
        if (all) p = q;
    }
    if (p < 0) return;
    std::vector<long long> off((size_t)L + 1);
    const long long u0 = (long long)std::ldexp(b[0], p);
    for (int64_t k = 0; k <= L; ++k) off[(size_t)k] = (long long)std::ldexp(b[k], p) - u0;
    const long long W = off[(size_t)L];
    if (W <= 0) return;
    long long q = W;
    int D = 0;
    while ((q & 1) == 0) { q >>= 1; ++D; }
    const int wordbits = (dtype == AI_DTYPE_F32) ? 32 : 64;
    if (D < 1 || D >= wordbits - 1) return;
    std::vector<unsigned long long> V((size_t)L + 1);
    for (int64_t k = 0; k <= L; ++k) {
        if (off[(size_t)k] % q) return;
        V[(size_t)k] = (unsigned long long)(off[(size_t)k] / q);
    }
    for (int64_t i = 0; i < L; ++i) {                       // every leaf a dyadic node
        const unsigned long long w = V[(size_t)i + 1] - V[(size_t)i];
        if (w == 0 || (w & (w - 1)) || (V[(size_t)i] & (w - 1))) return;
    }
    // leaves inside [lo, hi): indices [i0, i1)
    auto leaf_at = [&](unsigned long long v) {              // leaf holding position v
        return (int64_t)(std::upper_bound(V.begin(), V.end(), v) - V.begin()) - 1;
    };
    auto is_leaf = [&](unsigned long long lo, unsigned long long hi) {   // node [lo, hi) inside ONE leaf?
        const int64_t i = leaf_at(lo);
        return V[(size_t)i + 1] >= hi;
    };
    for (int Lc = 0; Lc <= std::min(D, kPathMaxRowsLog2); ++Lc) {
        const int rowlen = D - Lc + 1;
        const size_t rows = (size_t)1 << Lc;
        std::vector<unsigned long long> target(rows);
        std::vector<uint32_t> tab(rows * (size_t)rowlen);
        bool ok = true;
        for (size_t r = 0; r < rows && ok; ++r) {
            unsigned long long lo = (unsigned long long)r << (D - Lc), hi = lo + (1ull << (D - Lc));
            int level = Lc;                                 // [lo, hi) is the path node of this level
            uint32_t* row = &tab[r * (size_t)rowlen];
            while (!is_leaf(lo, hi)) {
                const unsigned long long mid = lo + ((hi - lo) >> 1);
                const bool lleaf = is_leaf(lo, mid), rleaf = is_leaf(mid, hi);
                if (!lleaf && !rleaf) { ok = false; break; }
                // the path continues into the child that is NOT a leaf (either one if both are);
                // positions that leave the path at bit `level` sit in the other child
                if (rleaf) { row[level - Lc] = (uint32_t)leaf_at(mid); hi = mid; }
                else       { row[level - Lc] = (uint32_t)leaf_at(lo);  lo = mid; }
                ++level;
            }
            if (!ok) break;
            for (int z = level; z <= D; ++z) row[z - Lc] = (uint32_t)leaf_at(lo);
            target[r] = lo;
        }
        if (!ok) continue;
        // exact division by q:  floor(u / q) = mulhi(u, M) >> s
        unsigned long long M = 0;
        unsigned sh = 0;
        auto divide = [&](unsigned long long u) -> unsigned long long {
            if (!M) return u;
            if (wordbits == 32) return (unsigned long long)((unsigned)(((unsigned long long)(unsigned)u * (unsigned)M) >> 32) >> sh);
            return (unsigned long long)(((unsigned __int128)u * M) >> 64) >> sh;
        };
        auto device_leaf = [&](unsigned long long u) -> int64_t {
            const unsigned long long v = divide(u);
            const unsigned long long row = v >> (D - Lc);
            if (row >= rows) return -1;
            const unsigned long long x = v ^ target[(size_t)row];
            int clz = wordbits;
            if (x) clz = (wordbits == 32) ? __builtin_clz((unsigned)x) : __builtin_clzll(x);
            const int e = clz - (wordbits - D) - Lc;
            if (e < 0 || e >= rowlen) return -1;
            return (int64_t)tab[(size_t)row * rowlen + e];
        };
        auto verify = [&]() {
            for (int64_t k = 0; k < L; ++k) {
                const unsigned long long a = (unsigned long long)off[(size_t)k], e = (unsigned long long)off[(size_t)k + 1] - 1;
                if (divide(a) != V[(size_t)k] || divide(e) != V[(size_t)k + 1] - 1) return false;
                if (device_leaf(a) != k || device_leaf(e) != k) return false;
                const unsigned long long m = a + ((e - a) >> 1);
                if (device_leaf(m) != k) return false;
            }
            return true;
        };
        bool found = (q == 1);
        if (!found) {
            for (unsigned s = 0; s < (unsigned)wordbits - 1 && !found; ++s) {
                const unsigned __int128 one = (unsigned __int128)1 << (wordbits + s);
                const unsigned __int128 m = (one + (unsigned __int128)q - 1) / (unsigned __int128)q;
                if (m >> wordbits) break;                   // M must fit the word
                M = (unsigned long long)m; sh = s;
                found = verify();
            }
            if (!found) { M = 0; sh = 0; }
        } else {
            found = verify();
        }
        if (!found) return;
        gp->path_ok = 1; gp->path_p = p; gp->path_D = D; gp->path_Lc = Lc;
        gp->path_u0 = u0; gp->path_u1 = u0 + W - 1;
        gp->path_mul = M; gp->path_shift = sh;
        gp->path_target.swap(target);
        gp->path_rows.swap(tab);
        return;
    }
}

int plan_grid(const std::vector<double>& b, int dtype, GridPlan* out, int64_t exact_budget_override = 0) {
    const int64_t L = (int64_t)b.size() - 1;
    const long long rep_limit = (dtype == AI_DTYPE_F32) ? (1LL << 24) : (1LL << 30);
    auto fits = [&](int p, int gmax, long long* g0, long long* gl) {
        if (!cells_for(b, p, g0, gl)) return false;
        if (*gl - *g0 + 1 > gmax) return false;
        return std::llabs(*g0) < rep_limit && std::llabs(*gl) < rep_limit;
    };
    // smallest p >= 0 at which all interior breakpoints are integers after scaling
    int p_exact = -1;
    for (int p = 0; p <= 60; ++p) {
        bool all = true;
        for (int64_t k = 1; k < L && all; ++k) {
            const double v = std::ldexp(b[k], p);
            all = (std::fabs(v) < 9.0e15) && (v == std::floor(v));
        }
        if (all) { p_exact = p; break; }
    }
    long long g0 = 0, gl = 0;
    int p = 0;
    bool exact = false;
    int budget_exact = kGridMaxExact, budget_search = kGridMaxSearch;
    if (L > kLargeTableLeaves) {
        budget_exact = (int)std::min<int64_t>(kGridMaxLarge, std::max<int64_t>(kGridMaxExact, 4 * L));
        budget_search = (int)std::min<int64_t>(kGridMaxLarge, 4 * L);
    }
    if (exact_budget_override > 0) budget_exact = (int)std::min<int64_t>(exact_budget_override, 1 << 22);
    if (p_exact >= 0 && fits(p_exact, grid_budget(budget_exact), &g0, &gl)) {
        p = p_exact;
        exact = true;
    } else {
        // finest grid within the budget; p may be negative for very wide domains
        bool found = false;
        for (int q = 60; q >= -1000; --q) {
            if (fits(q, grid_budget(budget_search), &g0, &gl)) { p = q; found = true; break; }
        }
        if (!found) return fail(AI_ERR_TABLE, "cannot build a lookup grid for domain [%g, %g]", b.front(), b.back());
    }
    // first[] and the search depth K of the grid with scale 2^pp covering cells [c0, c1]
    auto fill = [&](int pp, long long c0, long long c1, std::vector<uint32_t>* first, int* Kout) {
        const int G = (int)(c1 - c0 + 1);
        first->assign(G, 0);
        // A negative p scales DOWN and can underflow for denormal-range x; widen the search window
        // by one cell to the left so the result stays exact regardless.
        const int margin = (pp < 0) ? 1 : 0;
        int64_t maxspan = 0;
        for (int g = 0; g < G; ++g) {
            const double left = std::ldexp((double)(c0 + g - margin), -pp);
            const double right = std::ldexp((double)(c0 + g + 1), -pp);
            // leaf holding `left`: #{k: b[k] <= left} - 1, clamped
            int64_t f = (std::upper_bound(b.begin(), b.end(), left) - b.begin()) - 1;
            f = std::min<int64_t>(std::max<int64_t>(f, 0), L - 1);
            // last leaf with a point < right
            int64_t l = (std::lower_bound(b.begin(), b.end(), right) - b.begin()) - 1;
            l = std::min<int64_t>(std::max<int64_t>(l, 0), L - 1);
            if (g == G - 1) l = L - 1;
            if (g == 0) f = 0;
            (*first)[g] = (uint32_t)f;
            maxspan = std::max(maxspan, l - f);
        }
        int K = 0;
        while ((1LL << K) - 1 < maxspan) ++K;
        *Kout = K;
    };
    int K = 0;
    fill(p, g0, gl, &out->first, &K);
    if (exact && K != 0) return fail(AI_ERR_TABLE, "internal: exact grid has span %lld", (long long)((1LL << K) - 1));
    // When a search follows anyway, a grid of at most 128 byte entries (32 words: every lane of a
    // warp reads a different bank or the same word) is read in ONE shared-memory wavefront, where
    // random reads of the 4096-cell table cost about 4.5 (ncu, config 4).  Take the coarse grid if
    // the extra search steps cost less than that (a step is 1 wavefront in fp32, 2 in fp64).
    if (!exact && K > 0 && L <= 256 && (gl - g0 + 1) > 128 && !std::getenv("AI_B200_NO_COARSE_GRID")) {
        long long c0 = 0, c1 = 0;
        for (int q = p - 1; q >= -1000; --q) {
            if (!fits(q, grid_budget(128), &c0, &c1)) continue;
            std::vector<uint32_t> first2;
            int K2 = 0;
            fill(q, c0, c1, &first2, &K2);
            const double step = (dtype == AI_DTYPE_F32) ? 1.0 : 2.0;
            if (K2 * step + 1.0 < K * step + 4.5) { p = q; g0 = c0; gl = c1; K = K2; out->first.swap(first2); }
            break;
        }
    }
    out->p = p; out->G = (int)(gl - g0 + 1); out->g0 = g0;
    out->K = K;
    // a search would follow: the path lookup instead, when the tree is a path below some level
    if (!exact && K > 0) plan_path(b, dtype, out);
    if (out->path_ok) {
        out->p = out->path_p; out->G = 1 << out->path_Lc; out->g0 = out->path_u0;
        out->K = 0;
        out->first.clear();
    }
    if (!std::getenv("AI_B200_NO_RANK")) plan_rank(b, out);     // (env: test hook for the table path)
    return AI_OK;
}

int round_up16(size_t v) { return (int)((v + 15) & ~(size_t)15); }

// The one-element record header of modes kSmemP / kSmemBankP (see ai_kernels.cuh): possible when,
// for EVERY leaf, a_i has 9 (fp32) / 32 (fp64) trailing zero bits and 2/(b_i-a_i) differs from
// S0 = max_i 2/(b_i-a_i) only by j_i <= 255 in the exponent field.  Checked on the exact values
// the unpacked records would hold, so unpacking reproduces them bit for bit.
template <typename T> struct HeaderPack {
    bool ok = false;
    unsigned s0_hi = 0, s0_lo = 0;
    std::vector<T> word;               // per leaf: (bits of a) | j
};

template <typename T>
HeaderPack<T> plan_header_pack(const ai_table* t) {
    HeaderPack<T> hp;
    if (t->basis == ai::kMono || std::getenv("AI_B200_NO_PACKED_HDR")) return hp;
    const int64_t L = t->L;
    std::vector<T> a(L), s(L);
    T smax = 0;
    for (int64_t i = 0; i < L; ++i) {
        a[i] = (T)t->breaks[i];
        const T b = (T)t->breaks[i + 1];
        s[i] = (T)2 / (b - a[i]);
        if (!(s[i] > 0) || !std::isfinite(s[i])) return hp;
        smax = std::max(smax, s[i]);
    }
    hp.word.resize(L);
    if constexpr (sizeof(T) == 4) {
        uint32_t s0;
        std::memcpy(&s0, &smax, 4);
        for (int64_t i = 0; i < L; ++i) {
            uint32_t ab, sb;
            std::memcpy(&ab, &a[i], 4);
            std::memcpy(&sb, &s[i], 4);
            const uint32_t d = s0 - sb;
            if ((ab & 0x1FFu) || (d & 0x7FFFFFu) || (d >> 23) > 255u || (sb >> 23) == 0u) return hp;
            const uint32_t w = ab | (d >> 23);
            // the device formula, checked here on the host
            const uint32_t ua = w & 0xFFFFFE00u, us = w * 0xFF800000u + s0;
            if (ua != ab || us != sb) return hp;
            std::memcpy(&hp.word[i], &w, 4);
        }
        hp.s0_hi = s0;
    } else {
        uint64_t s0;
        std::memcpy(&s0, &smax, 8);
        for (int64_t i = 0; i < L; ++i) {
            uint64_t ab, sb;
            std::memcpy(&ab, &a[i], 8);
            std::memcpy(&sb, &s[i], 8);
            const uint64_t d = s0 - sb;
            if ((ab & 0xFFFFFFFFull) || (d & ((1ull << 52) - 1)) || (d >> 52) > 255ull || ((sb >> 52) & 0x7FF) == 0) return hp;
            const uint64_t w = ab | (d >> 52);
            const uint32_t hi = (uint32_t)(w >> 32), lo = (uint32_t)w;
            const uint64_t ua = (uint64_t)hi << 32;
            const uint64_t us = ((uint64_t)((uint32_t)(s0 >> 32) - (lo << 20)) << 32) | (uint32_t)s0;
            if (ua != ab || us != sb) return hp;
            std::memcpy(&hp.word[i], &w, 8);
        }
        hp.s0_hi = (uint32_t)(s0 >> 32);
        hp.s0_lo = (uint32_t)s0;
    }
    hp.ok = true;
    return hp;
}

// The computed record header of modes kSmemU / kSmemBankU: fp64 tables with one leaf per coarse
// cell whose stored header is reproduced exactly by a_c = fma(c, w, lb) and one constant s.
struct UniformHeader {
    bool ok = false;
    double w = 0, lb = 0, s = 0;
};

template <typename T>
UniformHeader plan_uniform_header(const ai_table* t, const GridPlan& gp) {
    UniformHeader u;
    const int64_t L = t->L;
    // fp32: only while the table stays packed (<= 32 cells); its bank-mode tables keep first[] + header
    if (t->basis == ai::kMono || !gp.coarse_ok || L <= 2 || (int64_t)gp.coarse_leaf.size() != L ||
        (sizeof(T) == 4 && L > 32) || std::getenv("AI_B200_NO_UNIFORM_HDR"))
        return u;
    const T lb = (T)t->breaks[0], w = (T)t->breaks[1] - (T)t->breaks[0];
    const T s = (T)2 / ((T)t->breaks[1] - (T)t->breaks[0]);
    for (int64_t c = 0; c < L; ++c) {
        if (gp.coarse_leaf[(size_t)c] != (uint32_t)c) return u;
        const T a = (T)t->breaks[c], b = (T)t->breaks[c + 1];
        if (std::fma((T)c, w, lb) != a || (T)2 / (b - a) != s) return u;     // bit for bit, in table arithmetic
    }
    u.lb = lb; u.w = w; u.s = s;
    u.ok = true;
    return u;
}

template <typename T>
int build_image(ai_table* t, const GridPlan& gp, bool global_layout, std::vector<unsigned char>* image,
                const HeaderPack<T>* pack = nullptr, bool want_cell = true, const UniformHeader* uni = nullptr) {
    const int n = t->order;
    const int64_t L = t->L;
    const bool uniform = uni && uni->ok && !global_layout && want_cell;
    const bool packed = pack && pack->ok && !global_layout && !uniform;
    // Records once per COARSE CELL instead of once per leaf: the cell index the lookup computes
    // anyway addresses the record; neither the rank nor the first[] table is needed on the
    // evaluation path.  Offered while it at most doubles the record array (near-uniform trees);
    // the caller keeps it only if the gathers stay conflict-free (plan_table).
    const int64_t Gc = (int64_t)gp.coarse_leaf.size();
    // Measured (A/B on one box, random points): up to 32 cells it replaces the bitmap rank and always
    // pays (cfg2 93 -> 97% of HBM); beyond 32 cells it replaces the first[] load, which pays for
    // fp64 tables (64-leaf o7 87 -> 94%, 128-leaf o6 84 -> 93%) but not for the HBM-bound fp32
    // order-3 tables (62 / 69 / 44 leaves: 93.5 / 98.8 / 93 -> 89%), so fp32 keeps first[] there.
    const bool cell_records = want_cell && gp.coarse_ok && Gc >= L && Gc <= std::max<int64_t>(2 * L, 32)
                              && (Gc <= 32 || sizeof(T) == 8 || global_layout) && !std::getenv("AI_B200_NO_CELL_RECORDS");
    const int64_t NR = cell_records ? Gc : L;            // records stored
    const int R = global_layout ? ai::rec_stride_global(t->basis, n, (int)sizeof(T))
                                : (uniform ? ai::rec_stride(ai::kMono, n)
                                           : (packed ? ai::rec_stride_packed(n) : ai::rec_stride(t->basis, n)));
    const int hdr = (t->basis == ai::kMono || uniform) ? 0 : (packed ? 1 : 2);
    const int pad = (gp.K > 0) ? (1 << gp.K) : 0;
    const int off_first = 0;
    int fshift = (L <= 256) ? 2 : (L <= 65536 ? 1 : 0);   // 8-, 16- or 32-bit entries
    // the leaf as a rank from memory (Lookup::rank16_guess) where a large first[] table would be read
    // with bank conflicts: leaf-indexed records, no search
    // Measured (A/B on one box, random points): against a 1 KB+ first[] table the rank words win
    // (429-leaf fp64 table 80.0 -> 81.4% of HBM); against the 320-640 byte tables of the fp32 order-3
    // tables the extra integer work loses (69 leaves: 98.9 -> 93.7%), so those keep first[].
    size_t rank16_min = 1024;                            // (tuning hook: AI_B200_RANK16_MIN_BYTES)
    if (const char* e = std::getenv("AI_B200_RANK16_MIN_BYTES")) rank16_min = (size_t)std::max(0L, std::atol(e));
    const bool rank16 = !gp.rank16.empty() && !cell_records && gp.K == 0 && !gp.rank_mode
                        && (size_t)gp.G * (size_t)(4 >> fshift) > rank16_min;
    if (rank16) fshift = 3;
    // path lookup: first[] = one target word per row (32-bit for float, 64-bit for double), then the
    // rows of leaf ordinals (bytes up to 256 leaves, else 16-bit)
    const bool path = gp.path_ok != 0;
    const int path_toff = path ? round_up16(gp.path_target.size() * sizeof(T)) : 0;
    const bool path_t16 = L > 256;
    const int off_breaks = path ? path_toff + round_up16(gp.path_rows.size() * (path_t16 ? 2 : 1))
                         : rank16 ? round_up16(gp.rank16.size() * 4) : round_up16((size_t)gp.G * (size_t)(4 >> fshift));
    const int off_records = off_breaks + round_up16((size_t)(L + 1 + pad) * sizeof(T));
    const size_t total = (size_t)off_records + (size_t)round_up16((size_t)NR * R * sizeof(T));
    image->assign(total, 0);
    if (path) {
        for (size_t r = 0; r < gp.path_target.size(); ++r) {
            if constexpr (sizeof(T) == 4) {
                const uint32_t w = (uint32_t)gp.path_target[r];
                std::memcpy(image->data() + off_first + 4 * r, &w, 4);
            } else {
                std::memcpy(image->data() + off_first + 8 * r, &gp.path_target[r], 8);
            }
        }
        unsigned char* rows = image->data() + off_first + path_toff;
        for (size_t e = 0; e < gp.path_rows.size(); ++e) {
            if (path_t16) { const uint16_t w = (uint16_t)gp.path_rows[e]; std::memcpy(rows + 2 * e, &w, 2); }
            else rows[e] = (unsigned char)gp.path_rows[e];
        }
    } else if (rank16) {
        std::memcpy(image->data() + off_first, gp.rank16.data(), gp.rank16.size() * sizeof(uint32_t));
    } else if (fshift == 2) {
        for (int g = 0; g < gp.G; ++g) (*image)[off_first + g] = (unsigned char)gp.first[g];
    } else if (fshift == 1) {
        uint16_t* f16 = reinterpret_cast<uint16_t*>(image->data() + off_first);
        for (int g = 0; g < gp.G; ++g) f16[g] = (uint16_t)gp.first[g];
    } else {
        std::memcpy(image->data() + off_first, gp.first.data(), gp.first.size() * sizeof(uint32_t));
    }
    t->view.first_shift = fshift;
    t->path = path ? 1 : 0;
    t->view.path_toff = path_toff;
    t->view.path_t16 = path_t16 ? 1 : 0;
    T* br = reinterpret_cast<T*>(image->data() + off_breaks);
    for (int64_t k = 0; k <= L; ++k) br[k] = (T)t->breaks[k];
    for (int k = 0; k < pad; ++k) br[L + 1 + k] = std::numeric_limits<T>::infinity();
    T* rec = reinterpret_cast<T*>(image->data() + off_records);
    for (int64_t ri = 0; ri < NR; ++ri) {
        const int64_t i = cell_records ? (int64_t)gp.coarse_leaf[(size_t)ri] : ri;      // the leaf stored here
        T* r = rec + ri * R;
        if (hdr == 2) {
            const T a = (T)t->breaks[i], b = (T)t->breaks[i + 1];
            r[0] = a;
            r[1] = (T)2 / (b - a);       // in T arithmetic: generate.py:83 `(2./(b - a))`
        } else if (hdr == 1) {
            r[0] = pack->word[i];
        }
        for (int j = 0; j <= n; ++j) r[hdr + j] = (T)t->coef[i * (n + 1) + j];
    }
    t->rstride = R;
    t->packed = packed ? 1 : 0;
    t->uniform_hdr = uniform ? 1 : 0;
    t->view.u_w = uniform ? uni->w : 0.0;
    t->view.u_lb = uniform ? uni->lb : 0.0;
    t->view.u_s = uniform ? uni->s : 0.0;
    t->view.s0_hi = packed ? pack->s0_hi : 0;
    t->view.s0_lo = packed ? pack->s0_lo : 0;
    t->view.image_bytes = (int)total;
    t->view.n_records_elems = (int)(NR * R);
    t->view.n_records = (int)NR;
    t->cell_records = cell_records ? 1 : 0;
    t->rank16 = rank16 ? 1 : 0;
    t->view.n_breaks_elems = (int)(L + 1 + pad);
    t->view.off_first = off_first;
    t->view.off_breaks = off_breaks;
    t->view.off_records = off_records;
    return AI_OK;
}

// The kGlobalC image of a dyadic table too large for shared memory (see ai_kernels.cuh):
//   [first[] by coarse cell][breaks][4-bit level per coarse cell][records: c0..cn per coarse cell, 32-byte padded]
// Possible when the EXACT grid exists (every breakpoint on a cell edge), every leaf spans 2^k
// coarse cells starting at a multiple of 2^k, and the kernel's header arithmetic reproduces the
// stored [a, 2/(b-a)] of every leaf bit for bit:  a = fma(first cell, w, lb),  2/(b-a) = S1 * 2^-k.
constexpr int64_t kCompactMaxCells = 1 << 17;       // 64 KB of levels in shared memory
template <typename T>
bool build_compact_image(ai_table* t, const GridPlan& gp, std::vector<unsigned char>* image) {
    const std::vector<uint32_t>& coarse_leaf = gp.wide_coarse_leaf;
    const int64_t L = t->L, Gc = (int64_t)coarse_leaf.size();
    const int n = t->order;
    if (gp.K != 0 || !gp.wide_ok || Gc < L || Gc > kCompactMaxCells) return false;
    if (Gc * (int64_t)ai::rec_stride_compact(n, (int)sizeof(T)) * (int64_t)sizeof(T) > ((int64_t)1 << 28)) return false;
    const bool hdr = t->basis != ai::kMono;
    // coarse cell width: the narrowest leaf is one cell wide
    std::vector<int64_t> first_cell((size_t)L, -1), n_cells((size_t)L, 0);
    for (int64_t c = 0; c < Gc; ++c) {
        const int64_t i = coarse_leaf[(size_t)c];
        if (i < 0 || i >= L) return false;
        if (first_cell[(size_t)i] < 0) first_cell[(size_t)i] = c;
        if (c != first_cell[(size_t)i] + n_cells[(size_t)i]) return false;          // cells of a leaf are contiguous
        ++n_cells[(size_t)i];
    }
    const T lb = (T)t->breaks[0];
    T w = 0;
    for (int64_t i = 0; i < L; ++i) {
        if (n_cells[(size_t)i] == 1) { w = (T)t->breaks[i + 1] - (T)t->breaks[i]; break; }
    }
    if (!(w > 0)) return false;
    const T s1 = (T)2 / w;
    std::vector<unsigned char> level((size_t)Gc, 0);
    for (int64_t i = 0; i < L; ++i) {
        const int64_t m = n_cells[(size_t)i], c0 = first_cell[(size_t)i];
        if (m < 1 || (m & (m - 1)) || (c0 & (m - 1))) return false;                  // 2^k cells, aligned
        int k = 0;
        while ((1LL << k) < m) ++k;
        if (k > 15) return false;
        if (hdr) {
            const T a = (T)t->breaks[i], b = (T)t->breaks[i + 1];
            const T s = (T)2 / (b - a);
            if (std::fma((T)(int)c0, w, lb) != a) return false;                      // the device formula, bit for bit
            if constexpr (sizeof(T) == 4) {
                uint32_t sb, s1b;
                std::memcpy(&sb, &s, 4); std::memcpy(&s1b, &s1, 4);
                if (s1b - ((uint32_t)k << 23) != sb || (sb >> 23) == 0) return false;
            } else {
                uint64_t sb, s1b;
                std::memcpy(&sb, &s, 8); std::memcpy(&s1b, &s1, 8);
                const uint64_t dev = ((uint64_t)((uint32_t)(s1b >> 32) - ((uint32_t)k << 20)) << 32) | (uint32_t)s1b;
                if (dev != sb || ((sb >> 52) & 0x7FF) == 0) return false;
            }
        }
        for (int64_t c = c0; c < c0 + m; ++c) level[(size_t)c] = (unsigned char)k;
    }
    const int R = ai::rec_stride_compact(n, (int)sizeof(T));
    const int fshift = (L <= 65536) ? 1 : 0;                     // first[] by coarse cell: 16- or 32-bit
    const int off_first = 0;
    const int off_breaks = round_up16((size_t)Gc * (size_t)(4 >> fshift));
    const int off_aux = off_breaks + round_up16((size_t)(L + 1) * sizeof(T));
    const int aux_bytes = hdr ? round_up16((size_t)(Gc + 1) / 2) : 0;
    const size_t off_records = ((size_t)off_aux + (size_t)aux_bytes + 31) & ~(size_t)31;   // 32-byte aligned records
    const size_t total = off_records + (size_t)Gc * R * sizeof(T);
    image->assign((total + 15) & ~(size_t)15, 0);
    if (fshift == 1) {
        uint16_t* f16 = reinterpret_cast<uint16_t*>(image->data() + off_first);
        for (int64_t c = 0; c < Gc; ++c) f16[c] = (uint16_t)coarse_leaf[(size_t)c];
    } else {
        std::memcpy(image->data() + off_first, coarse_leaf.data(), (size_t)Gc * sizeof(uint32_t));
    }
    T* br = reinterpret_cast<T*>(image->data() + off_breaks);
    for (int64_t k = 0; k <= L; ++k) br[k] = (T)t->breaks[k];
    if (hdr) {
        unsigned char* nib = image->data() + off_aux;
        for (int64_t c = 0; c < Gc; ++c) nib[c >> 1] |= (unsigned char)(level[(size_t)c] << ((c & 1) * 4));
    }
    T* rec = reinterpret_cast<T*>(image->data() + off_records);
    for (int64_t c = 0; c < Gc; ++c) {
        const int64_t i = coarse_leaf[(size_t)c];
        for (int j = 0; j <= n; ++j) rec[c * R + j] = (T)t->coef[(size_t)i * (n + 1) + j];
    }
    t->rstride = R;
    t->packed = 0; t->uniform_hdr = 0; t->rank16 = 0; t->path = 0;
    t->cell_records = 1;
    t->compact = 1;
    t->view.first_shift = fshift;
    t->view.u_w = (double)w; t->view.u_lb = (double)lb; t->view.u_s = (double)s1;
    if constexpr (sizeof(T) == 4) {
        uint32_t s1b; std::memcpy(&s1b, &s1, 4);
        t->view.s0_hi = s1b; t->view.s0_lo = 0;
    } else {
        uint64_t s1b; std::memcpy(&s1b, &s1, 8);
        t->view.s0_hi = (uint32_t)(s1b >> 32); t->view.s0_lo = (uint32_t)s1b;
    }
    t->view.image_bytes = (int)image->size();
    t->view.n_records_elems = (int)(Gc * R);
    t->view.n_records = (int)Gc;
    t->view.n_breaks_elems = (int)(L + 1);
    t->view.off_first = off_first;
    t->view.off_breaks = off_breaks;
    t->view.off_records = (int)off_records;
    t->view.off_aux = off_aux;
    t->view.aux_bytes = aux_bytes;
    t->view.wide_mul = gp.wide_mul;
    t->view.wide_shift = gp.wide_shift;
    return true;
}

// what build_image / build_compact_image write into the table, to put it back when a layout is
// built but not taken
struct SavedPlan {
    int rstride, packed, uniform_hdr, rank16, cell_records, compact, path;
    ai::TableView view;
    explicit SavedPlan(const ai_table* t)
        : rstride(t->rstride), packed(t->packed), uniform_hdr(t->uniform_hdr), rank16(t->rank16),
          cell_records(t->cell_records), compact(t->compact), path(t->path), view(t->view) {}
    void restore(ai_table* t) const {
        t->rstride = rstride; t->packed = packed; t->uniform_hdr = uniform_hdr; t->rank16 = rank16;
        t->cell_records = cell_records; t->compact = compact; t->path = path; t->view = view;
    }
};

int prepare_dispatch(int basis, int dtype, int order, int mode, int smem, int* bps) {
    cudaError_t e = cudaErrorInvalidValue;
    if (dtype == AI_DTYPE_F32) {
        if (basis == ai::kCheb) e = ai::prepare_cheb_f32(order, mode, smem, bps);
        else if (basis == ai::kLeg) e = ai::prepare_leg_f32(order, mode, smem, bps);
        else e = ai::prepare_mono_f32(order, mode, smem, bps);
    } else {
        if (basis == ai::kCheb) e = ai::prepare_cheb_f64(order, mode, smem, bps);
        else if (basis == ai::kLeg) e = ai::prepare_leg_f64(order, mode, smem, bps);
        else e = ai::prepare_mono_f64(order, mode, smem, bps);
    }
    if (e != cudaSuccess) return fail(AI_ERR_CUDA, "kernel prepare failed: %s", cudaGetErrorString(e));
    return AI_OK;
}

// Everything about a table that is decided on the HOST, without touching CUDA: validation, the
// lookup grid, the record layout and the residency mode for a device with `smem_cap` bytes of
// opt-in shared memory per CTA.  Fills `t` and the packed image.  (Exposed as ai_table_plan so
// that the CPU-only test tier covers this logic.)
int plan_table(ai_table* t, size_t smem_cap, std::vector<unsigned char>* image_out) {
    // validate the flattened table
    const int64_t L = t->L;
    for (int64_t k = 0; k <= L; ++k)
        if (!std::isfinite(t->breaks[k])) return fail(AI_ERR_TABLE, "breakpoint %lld is not finite", (long long)k);
    for (int64_t k = 0; k < L; ++k)
        if (!(t->breaks[k] < t->breaks[k + 1]))
            return fail(AI_ERR_TABLE, "breakpoints not strictly increasing at %lld (%.17g, %.17g)",
                        (long long)k, t->breaks[k], t->breaks[k + 1]);
    if (L > kMaxLeaves) return fail(AI_ERR_UNSUPPORTED, "more than %lld leaves (%lld) not supported", (long long)kMaxLeaves, (long long)L);
    if ((int64_t)L * (t->order + 4) * 8 > (int64_t)1 << 30)
        return fail(AI_ERR_UNSUPPORTED, "table image would exceed 1 GiB (%lld leaves, order %d)", (long long)L, t->order);

    GridPlan gp;
    int rc = plan_grid(t->breaks, t->dtype, &gp);
    if (rc) return rc;
    t->G = gp.G; t->p = gp.p; t->K = gp.K; t->g0 = gp.g0;

    if (const char* e = std::getenv("AI_B200_SMEM_CAP_KB")) {            // test / tuning hook
        const long kb = std::atol(e);
        if (kb >= 16 && (size_t)kb * 1024 < smem_cap) smem_cap = (size_t)kb * 1024;
    }

    // (the kernels also hold a few bytes of static shared memory: keep a margin below the opt-in limit
    // so that a table planned at the very edge cannot fail to launch)
    if (smem_cap > 1024) smem_cap -= 256;

    // Build the shared-memory image first -- with the one-word record header when every leaf
    // qualifies (bit-identical values, one shared-memory wavefront less per 32 points, two in
    // fp64, and a shorter record) -- and if it cannot be staged rebuild it in the global layout
    // (16-byte padded records, two-element header).
    const char* force = std::getenv("AI_B200_FORCE_MODE");                // test hook
    const int forced = force ? std::atoi(force) : -1;
    std::vector<unsigned char>& image = *image_out;
    HeaderPack<float> hp32;
    HeaderPack<double> hp64;
    if (t->dtype == AI_DTYPE_F32) hp32 = plan_header_pack<float>(t); else hp64 = plan_header_pack<double>(t);
    const UniformHeader uni = (t->dtype == AI_DTYPE_F32) ? plan_uniform_header<float>(t, gp)
                                                         : plan_uniform_header<double>(t, gp);
    const size_t esz = (t->dtype == AI_DTYPE_F32) ? 4 : 8;
    const int full_copies = (t->dtype == AI_DTYPE_F32) ? 32 : 16;
    int mode = ai::kSmem, copies = 1, bcopies = 1, vc_smem = 0;
    bool go_global = false;
    // First with one record per coarse cell (when the grid allows it); kept only if the gathers
    // stay conflict-free -- packed with one record per bank, or fully replicated.  Otherwise
    // planned again with one record per leaf.
    // (does the one-record-per-leaf image fit in shared memory?  Then the cell layout must not push
    // the table out to the global path just because it is up to twice as large.)
    rc = (t->dtype == AI_DTYPE_F32) ? build_image<float>(t, gp, false, &image, &hp32, false, &uni)
                                    : build_image<double>(t, gp, false, &image, &hp64, false, &uni);
    if (rc) return rc;
    const bool leaf_image_fits = (size_t)t->view.image_bytes <= smem_cap;
    // Too large for shared memory: the compact cell-record image when the table is dyadic (exact grid
    // allowed to grow to 2^20 cells for this purpose only).
    bool compact = false;
    if (((!leaf_image_fits && forced < 0 && !std::getenv("AI_B200_NO_COMPACT")) || forced == ai::kGlobalC)
        && !std::getenv("AI_B200_NO_CELL_RECORDS")) {
        GridPlan gpx;
        if (plan_grid(t->breaks, t->dtype, &gpx, kGridMaxLarge) == AI_OK) {
            compact = (t->dtype == AI_DTYPE_F32) ? build_compact_image<float>(t, gpx, &image)
                                                 : build_compact_image<double>(t, gpx, &image);
            if (compact) {
                gp = gpx;
                t->G = gp.G; t->p = gp.p; t->K = gp.K; t->g0 = gp.g0;
                mode = ai::kGlobalC;
            }
        }
        if (!compact) {       // (build_image state of the leaf layout is rebuilt by the attempts below)
            t->compact = 0;
        }
    }
    for (int attempt = 0; attempt < 2 && !compact; ++attempt) {
    const bool want_cell = (attempt == 0);
    rc = (t->dtype == AI_DTYPE_F32) ? build_image<float>(t, gp, false, &image, &hp32, want_cell, &uni)
                                    : build_image<double>(t, gp, false, &image, &hp64, want_cell, &uni);
    if (rc) return rc;
    go_global = (forced == ai::kGlobal) || ((size_t)t->view.image_bytes > smem_cap);
    if (go_global) {
        // (cell-indexed records also here: the record gather then needs no first[] access before it)
        rc = (t->dtype == AI_DTYPE_F32) ? build_image<float>(t, gp, true, &image, nullptr, want_cell)
                                        : build_image<double>(t, gp, true, &image, nullptr, want_cell);
        if (rc) return rc;
    }
    const int64_t NR = t->view.n_records;                // records a warp may have to tell apart

    // ---- table residency -------------------------------------------------------------------
    // shared-memory bytes of the bank layouts with `cp` copies: [grid table][breaks x cp][records x cp]
    auto bank_layout = [&](int cp, int* soff_b, int* soff_r) -> size_t {
        const size_t ob = (size_t)t->view.off_breaks;
        const size_t orr = ob + (size_t)round_up16((size_t)t->view.n_breaks_elems * esz * cp);
        if (soff_b) *soff_b = (int)ob;
        if (soff_r) *soff_r = (int)orr;
        return orr + (size_t)round_up16((size_t)t->view.n_records_elems * esz * cp);
    };
    // kSmemV: [grid / rank words][breaks x bcp][16-byte padded leaf records x 8], 128-bit gathers
    auto vec_layout = [&](int bcp) -> size_t {
        const size_t Rg = (size_t)ai::rec_stride_global(t->basis, t->order, (int)esz);
        return (size_t)t->view.off_breaks + (size_t)round_up16((size_t)t->view.n_breaks_elems * esz * bcp)
               + (size_t)round_up16((size_t)L * Rg * esz * ai::kVecCopies);
    };
    // leaves a warp can address without two of them sharing a bank (odd record stride)
    const int64_t conflict_free_leaves = full_copies;
    mode = ai::kSmem; copies = 1; bcopies = 1;
    if (go_global) {
        mode = ai::kGlobal;
    } else if (NR > conflict_free_leaves || forced == ai::kSmemBank || forced == ai::kSmemBankN || forced == ai::kSmemV) {
        const bool vec_ok = !t->cell_records && !std::getenv("AI_B200_NO_VEC_MODE") && forced != ai::kSmemBankN
                            && (forced == ai::kSmemV || forced < 0);
        if (bank_layout(full_copies, nullptr, nullptr) <= smem_cap && forced != ai::kSmemBankN && forced != ai::kSmemV) {
            mode = ai::kSmemBank; copies = full_copies;
        } else if (vec_ok && vec_layout(1) <= smem_cap) {
            // full per-bank replication does not fit: 8 copies of 16-byte padded records, 128-bit loads
            // (conflict-free per quarter-warp); the breakpoints per bank too when a search reads them
            mode = ai::kSmemV; copies = ai::kVecCopies;
            bcopies = (gp.K > 0 && vec_layout(full_copies) <= smem_cap) ? full_copies : 1;
        } else if (gp.K == 0 || forced == ai::kSmemBankN) {
            // partial replication pays only when no search follows (measured: the 429-leaf 64o5
            // table 39% -> 44% of HBM on random points; config 4's 220-leaf K=6 table gets slower)
            for (int cp = full_copies / 2; cp >= 2; cp /= 2)
                if (bank_layout(cp, nullptr, nullptr) <= smem_cap) { mode = ai::kSmemBankN; copies = cp; break; }
        }
    }
    if (!go_global && forced == ai::kSmem) { mode = ai::kSmem; copies = 1; }
    // (NR == L: one leaf per cell, the cell layout costs nothing and is kept in every mode)
    const bool layout_ok = !t->cell_records || NR == L ||
                           (mode == ai::kGlobal && (!leaf_image_fits || forced == ai::kGlobal)) ||
                           mode == ai::kSmemBank || (mode == ai::kSmem && NR <= conflict_free_leaves);
    if (mode == ai::kSmemBank || mode == ai::kSmemBankN) bcopies = copies;
    // only these have computed-header kernels (the bank form for fp64 only)
    const bool header_ok = !t->uniform_hdr || mode == ai::kSmem || (mode == ai::kSmemBank && t->dtype == AI_DTYPE_F64);
    if (layout_ok && header_ok) break;
    }   // attempts
    auto bank_layout = [&](int cp, int* soff_b, int* soff_r) -> size_t {
        const size_t ob = (size_t)t->view.off_breaks;
        const size_t orr = ob + (size_t)round_up16((size_t)t->view.n_breaks_elems * esz * cp);
        if (soff_b) *soff_b = (int)ob;
        if (soff_r) *soff_r = (int)orr;
        return orr + (size_t)round_up16((size_t)t->view.n_records_elems * esz * cp);
    };
    // Eight-copy layout of a DYADIC table: the compact cell records (no header words, no rank lookup)
    // when their eight copies fit -- 429-leaf 64o5: 3 loads per point instead of 4 + a rank word.
    if ((mode == ai::kSmemV && forced < 0 && !std::getenv("AI_B200_NO_COMPACT") && !std::getenv("AI_B200_NO_CELL_RECORDS"))
        || (forced == ai::kSmemVC && !go_global && !std::getenv("AI_B200_NO_CELL_RECORDS"))) {
        GridPlan gpx;
        std::vector<unsigned char> image_c;
        const SavedPlan saved(t);
        if (plan_grid(t->breaks, t->dtype, &gpx, kGridMaxLarge) == AI_OK) {
            const bool ok = (t->dtype == AI_DTYPE_F32) ? build_compact_image<float>(t, gpx, &image_c)
                                                       : build_compact_image<double>(t, gpx, &image_c);
            const int nv = ai::rec_stride_vc(t->order, (int)esz) * (int)esz / 16;
            const size_t need = (size_t)t->view.aux_bytes + (size_t)t->view.n_records * nv * 16 * ai::kVecCopies;
            if (ok && need <= smem_cap) {
                image.swap(image_c);
                gp = gpx;
                t->G = gp.G; t->p = gp.p; t->K = gp.K; t->g0 = gp.g0;
                mode = ai::kSmemVC; compact = true;
                copies = nv;                                                   // chunks per shared-memory record
                bcopies = ai::rec_stride_compact(t->order, (int)esz) * (int)esz / 16;   // chunks per global record
                t->view.soff_records = t->view.aux_bytes;
                vc_smem = (int)need;
            } else {
                saved.restore(t);
            }
        }
    }
    if (mode == ai::kSmemV) {
        // the records in the 16-byte padded global layout ([a, 2/(b-a), c0..cn, pad]), one per leaf
        rc = (t->dtype == AI_DTYPE_F32) ? build_image<float>(t, gp, true, &image, nullptr, false)
                                        : build_image<double>(t, gp, true, &image, nullptr, false);
        if (rc) return rc;
        if (t->cell_records) return fail(AI_ERR_TABLE, "internal: kSmemV planned with cell-indexed records");
    }
    // two leaves: both records ride in the kernel parameters, element-wise select instead of a gather
    // (the image stays: the reduction and multi-table kernels read it)
    if (!compact && !go_global && ((L == 2 && forced < 0) || (L <= 2 && forced == ai::kSelect))) {
        mode = ai::kSelect; copies = 1;
        const int hdr = (t->basis == ai::kMono) ? 0 : 2;
        t->select.assign(1 + 2 * ai::kMaxRecord, 0.0);
        t->select[0] = (L == 2) ? t->breaks[1] : std::numeric_limits<double>::infinity();
        for (int i = 0; i < 2; ++i) {
            const int64_t leaf = std::min<int64_t>(i, L - 1);
            double* rec = &t->select[1 + i * ai::kMaxRecord];
            if (hdr) {           // the values build_image stores: a and 2/(b-a) in table-dtype arithmetic
                if (t->dtype == AI_DTYPE_F32) {
                    const float a = (float)t->breaks[leaf], b = (float)t->breaks[leaf + 1];
                    rec[0] = a; rec[1] = 2.0f / (b - a);
                } else {
                    const double a = t->breaks[leaf], b = t->breaks[leaf + 1];
                    rec[0] = a; rec[1] = 2.0 / (b - a);
                }
            }
            for (int k = 0; k <= t->order; ++k) rec[hdr + k] = t->coef[(size_t)leaf * (t->order + 1) + k];
        }
    }
    if (t->packed && mode == ai::kSmem) mode = ai::kSmemP;
    if (t->packed && mode == ai::kSmemBank) mode = ai::kSmemBankP;
    if (t->uniform_hdr && mode == ai::kSmem) mode = ai::kSmemU;
    if (t->uniform_hdr && mode == ai::kSmemBank) mode = ai::kSmemBankU;
    t->mode = mode;
    t->copies = copies;
    t->view.copies = copies;
    t->view.bcopies = bcopies;
    t->smem_resident = (mode != ai::kGlobal && mode != ai::kGlobalC);
    t->eval_smem = 0;
    if (mode == ai::kGlobalC) t->eval_smem = t->view.aux_bytes;
    if (mode == ai::kSmemVC) t->eval_smem = vc_smem;
    if (mode == ai::kSmem || mode == ai::kSmemP || mode == ai::kSmemU) t->eval_smem = t->view.image_bytes;
    if (mode == ai::kSmemV) {
        t->view.soff_breaks = t->view.off_breaks;
        t->view.soff_records = t->view.off_breaks + round_up16((size_t)t->view.n_breaks_elems * esz * bcopies);
        t->eval_smem = t->view.soff_records + round_up16((size_t)t->view.n_records_elems * esz * ai::kVecCopies);
    } else if (ai::mode_bank(mode)) {
        t->eval_smem = (int)bank_layout(copies, &t->view.soff_breaks, &t->view.soff_records);
    }
    t->aux_smem = (mode == ai::kGlobal) ? 0
                  : ((mode == ai::kGlobalC || mode == ai::kSmemVC) ? t->view.aux_bytes : t->view.image_bytes);
    t->view.image = nullptr;
    t->view.n_leaves = (int)L;
    t->view.g0 = (int)gp.g0;
    t->view.kstep = gp.K > 0 ? (1 << (gp.K - 1)) : 0;
    t->view.rank_mode = (gp.rank_mode ? 1 : 0) | (t->cell_records ? 2 : 0) | (t->rank16 ? 4 : 0) | (compact ? 8 : 0)
                        | (t->path ? 16 : 0);
    if (t->path) {
        const int wordbits = (t->dtype == AI_DTYPE_F32) ? 32 : 64;
        t->view.path_mul = gp.path_mul;
        t->view.path_shift = gp.path_shift;
        t->view.path_cshift = (unsigned)(gp.path_D - gp.path_Lc);
        t->view.path_rowlen = gp.path_D - gp.path_Lc + 1;
        t->view.path_zbias = (wordbits - gp.path_D) + gp.path_Lc;
    }
    t->view.div_mul = gp.div_mul;
    t->view.div_shift = gp.div_shift;
    t->view.div_add = gp.div_add;
    t->view.cont_mask = gp.cont;
    t->rank_mode = gp.rank_mode;
    t->view.scale = std::ldexp(1.0, gp.p);
    t->view.lo = (double)gp.g0;
    t->view.hi = (double)(gp.g0 + gp.G - 1);
    if (t->path) {                       // the fine cells of the path lookup (not the row count G)
        t->view.hi = (double)gp.path_u1;
        t->view.g0 = (t->dtype == AI_DTYPE_F32) ? (int)gp.path_u0 : 0;
    }
    // adapt.py:107 `*(1./n)`: 1./n in double, then rounded to the table dtype by the multiply
    const double inv = t->order > 0 ? 1.0 / (double)t->order : 0.0;
    t->view.inv_order = (t->dtype == AI_DTYPE_F32) ? (double)(float)inv : inv;
    return AI_OK;
}

// plan on the host, then stage the image on the table's device and prepare its kernels
int finish_table(ai_table* t) {
    DeviceGuard guard(t->device);
    if (!guard.ok) return fail(AI_ERR_CUDA, "cannot select device %d", t->device);
    DeviceInfo prop;
    int rc = device_info(t->device, &prop);
    if (rc) return rc;
    if (prop.major < 10)
        return fail(AI_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only",
                    t->device, prop.major, prop.minor);
    t->sm_count = prop.sm_count;
    std::vector<unsigned char> image;
    rc = plan_table(t, prop.smem_optin, &image);
    if (rc) return rc;
    AI_CUDA(cudaMalloc(&t->d_image, image.size()));
    AI_CUDA(cudaMemcpy(t->d_image, image.data(), image.size(), cudaMemcpyHostToDevice));
    t->view.image = static_cast<const unsigned char*>(t->d_image);

    int bps = 0;
    rc = prepare_dispatch(t->basis, t->dtype, t->order, t->mode, t->eval_smem, &bps);
    if (rc) return rc;
    if (bps < 1) return fail(AI_ERR_CUDA, "kernel does not fit on an SM with %d B of shared memory", t->eval_smem);
    if (t->mode == ai::kSmemVC) {
        int dummy = 0;      // its reduction runs the kGlobalC kernel on the same image
        rc = prepare_dispatch(t->basis, t->dtype, t->order, ai::kGlobalC, t->aux_smem, &dummy);
        if (rc) return rc;
    }
    if (t->mode != ai::kGlobal && t->mode != ai::kSmemV && t->mode != ai::kGlobalC && t->mode != ai::kSmemVC
        && t->aux_smem > 48 * 1024) {
        int dummy = 0;      // opt the reduction kernel (packed layout) into large shared memory too
        rc = prepare_dispatch(t->basis, t->dtype, t->order,
                              t->uniform_hdr ? ai::kSmemU : (t->packed ? ai::kSmemP : ai::kSmem), t->aux_smem, &dummy);
        if (rc) return rc;
    }
    t->blocks_per_sm = bps;
    return AI_OK;
}

// before anything is allocated for the table
int check_leaf_count(int64_t n_leaves, int order) {
    if (n_leaves < 1) return fail(AI_ERR_TABLE, "a table needs at least one leaf (got %lld)", (long long)n_leaves);
    if (n_leaves > kMaxLeaves)
        return fail(AI_ERR_UNSUPPORTED, "more than %lld leaves (%lld) not supported", (long long)kMaxLeaves, (long long)n_leaves);
    if (n_leaves * (int64_t)(order + 4) * 8 > (int64_t)1 << 30)
        return fail(AI_ERR_UNSUPPORTED, "table image would exceed 1 GiB (%lld leaves, order %d)", (long long)n_leaves, order);
    return AI_OK;
}

int check_shape(int basis, int order, int dtype) {
    if (basis < 0 || basis > 2) return fail(AI_ERR_INVALID, "basis %d not in {0 chebyshev, 1 legendre, 2 monomial}", basis);
    if (order < 0 || order > AI_MAX_ORDER) return fail(AI_ERR_INVALID, "order %d not in [0, %d]", order, AI_MAX_ORDER);
    if (dtype != AI_DTYPE_F32 && dtype != AI_DTYPE_F64) return fail(AI_ERR_INVALID, "dtype %d not in {0 f32, 1 f64}", dtype);
    return AI_OK;
}

int check_common(int basis, int order, int dtype, int device) {
    int rc = check_shape(basis, order, dtype);
    if (rc) return rc;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return fail(AI_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(AI_ERR_INVALID, "device %d not in [0, %d)", device, count);
    return AI_OK;
}

template <typename T>
int launch_eval_t(const ai_table* t, const T* x, T* y, int64_t n, cudaStream_t stream) {
    if (n == 0) return AI_OK;
    constexpr int VN = ai::Vec<T>::N;
    const uintptr_t ax = reinterpret_cast<uintptr_t>(x), ay = reinterpret_cast<uintptr_t>(y);
    if ((ax % sizeof(T)) || (ay % sizeof(T))) return fail(AI_ERR_INVALID, "x / y not aligned to the element size");
    // scalar head up to the first 16-byte boundary of y, then 16-byte vectors, then a scalar tail.  When
    // x and y can never be 16-byte aligned together (x is a view that starts one element into its
    // buffer, say) the stores stay 128-bit and the kernel reads the tile's x element by element; routing
    // the whole call through the scalar loop instead cost a factor of six (profiles/r3_misaligned_*.txt).
    long long head = (long long)(((16 - (ay % 16)) % 16) / sizeof(T));
    if (head > n) head = n;
    const long long nvec = (n - head) / VN;
    ai::TableView view = t->view;
    view.x_elem_loads = ((ax % 16) != (ay % 16)) ? 1 : 0;
    ai::LaunchCfg cfg;
    const long long tile_elems = (long long)ai::kBlock * ai::Unroll<T>::value * VN;
    const long long want = (n + tile_elems - 1) / tile_elems;
    cfg.grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)t->sm_count * t->blocks_per_sm));
    cfg.smem_bytes = t->eval_smem;
    cfg.stream = stream;
    cfg.select = t->select.empty() ? nullptr : t->select.data();
    cudaError_t e;
    const int m = t->mode;
    if constexpr (sizeof(T) == 4) {
        if (t->basis == ai::kCheb) e = ai::launch_eval_cheb_f32(t->order, m, cfg, view, x, y, n, head, nvec);
        else if (t->basis == ai::kLeg) e = ai::launch_eval_leg_f32(t->order, m, cfg, view, x, y, n, head, nvec);
        else e = ai::launch_eval_mono_f32(t->order, m, cfg, view, x, y, n, head, nvec);
    } else {
        if (t->basis == ai::kCheb) e = ai::launch_eval_cheb_f64(t->order, m, cfg, view, x, y, n, head, nvec);
        else if (t->basis == ai::kLeg) e = ai::launch_eval_leg_f64(t->order, m, cfg, view, x, y, n, head, nvec);
        else e = ai::launch_eval_mono_f64(t->order, m, cfg, view, x, y, n, head, nvec);
    }
    if (e != cudaSuccess) return fail(AI_ERR_CUDA, "eval kernel launch failed: %s", cudaGetErrorString(e));
    return AI_OK;
}

template <typename T>
int eval_entry(const ai_table* t, const T* x, T* y, int64_t n, void* stream) {
    if (!t) return fail(AI_ERR_INVALID, "table handle is NULL");
    if (n < 0) return fail(AI_ERR_INVALID, "n = %lld is negative", (long long)n);
    if ((sizeof(T) == 4) != (t->dtype == AI_DTYPE_F32))
        return fail(AI_ERR_DTYPE, "table dtype is %s but the %s entry point was called",
                    t->dtype == AI_DTYPE_F32 ? "f32" : "f64", sizeof(T) == 4 ? "f32" : "f64");
    if (n > 0 && (!x || !y)) return fail(AI_ERR_INVALID, "x or y is NULL");
    DeviceGuard guard(t->device);
    if (!guard.ok) return fail(AI_ERR_CUDA, "cannot select device %d", t->device);
    return launch_eval_t<T>(t, x, y, n, static_cast<cudaStream_t>(stream));
}

// Several tables, one pass over x.  Fused when every table can be staged
// packed into shared memory next to the others and all share basis / order; otherwise the same
// result comes from one single-table launch per table on the same stream.
template <typename T>
int multi_entry(const ai_table* const* tabs, int m, const T* x, T* const* y, int64_t n, void* stream, int* fused_out) {
    if (fused_out) *fused_out = 0;
    if (!tabs || !y) return fail(AI_ERR_INVALID, "tables or y is NULL");
    if (m < 1 || m > ai::kMaxMulti) return fail(AI_ERR_INVALID, "m = %d not in [1, %d]", m, ai::kMaxMulti);
    if (n < 0) return fail(AI_ERR_INVALID, "n = %lld is negative", (long long)n);
    for (int j = 0; j < m; ++j) {
        if (!tabs[j]) return fail(AI_ERR_INVALID, "table %d is NULL", j);
        if ((sizeof(T) == 4) != (tabs[j]->dtype == AI_DTYPE_F32))
            return fail(AI_ERR_DTYPE, "table %d: dtype does not match the entry point", j);
        if (tabs[j]->device != tabs[0]->device)
            return fail(AI_ERR_INVALID, "table %d lives on device %d, table 0 on device %d", j, tabs[j]->device, tabs[0]->device);
        if (n > 0 && !y[j]) return fail(AI_ERR_INVALID, "y[%d] is NULL", j);
        for (int i = 0; i < j; ++i)
            if (n > 0 && y[i] == y[j]) return fail(AI_ERR_INVALID, "y[%d] and y[%d] are the same buffer", i, j);
    }
    if (n == 0) return AI_OK;
    if (!x) return fail(AI_ERR_INVALID, "x is NULL");
    const ai_table* t0 = tabs[0];
    DeviceGuard guard(t0->device);
    if (!guard.ok) return fail(AI_ERR_CUDA, "cannot select device %d", t0->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);

    DeviceInfo prop;
    int rc = device_info(t0->device, &prop);
    if (rc) return rc;
    const uintptr_t ax = reinterpret_cast<uintptr_t>(x);
    // Fusing saves (m - 1) reads of x but keeps every table's record gathers; it pays while a record is
    // short (measured, profiles/r1_multi_table.txt: fp32 / fp64 order 6 gain 1.07-1.41x, fp64 order 13
    // on random points 0.95x), so long records take the per-table launches.  Test hook:
    // AI_B200_FORCE_FUSED_MULTI=1 fuses whenever the shapes allow it.
    // Two-leaf tables lose too (fp32 order 13, m = 2: 0.86x random, 0.97x sorted): alone they run
    // the select kernel, fused they would gather from shared memory.
    const bool forced_fuse = std::getenv("AI_B200_FORCE_FUSED_MULTI") != nullptr;
    bool worth_fusing = (size_t)(t0->order + 1) * sizeof(T) <= 64;
    for (int j = 0; j < m; ++j) worth_fusing = worth_fusing && tabs[j]->mode != ai::kSelect;
    bool fuse = (m >= 2) && (worth_fusing || forced_fuse) && !std::getenv("AI_B200_NO_FUSED_MULTI") && (ax % sizeof(T)) == 0;
    size_t smem = 0;
    ai::MultiArgs args{};
    for (int j = 0; j < m && fuse; ++j) {
        const ai_table* t = tabs[j];
        fuse = t->basis == t0->basis && t->order == t0->order && t->packed == t0->packed
               && t->uniform_hdr == t0->uniform_hdr
               && (t->mode == ai::kSmem || t->mode == ai::kSmemP || t->mode == ai::kSmemU || t->mode == ai::kSelect)
               && (reinterpret_cast<uintptr_t>(y[j]) % 16) == (ax % 16);
        args.tv[j] = t->view;
        args.y[j] = y[j];
        args.soff[j] = (int)smem;
        smem += (size_t)t->view.image_bytes;          // multiple of 16
    }
    fuse = fuse && smem <= prop.smem_optin;
    if (!fuse) {
        for (int j = 0; j < m; ++j) {
            rc = launch_eval_t<T>(tabs[j], x, y[j], n, s);
            if (rc) return rc;
        }
        return AI_OK;
    }
    args.m = m;
    const int hdr_form = t0->uniform_hdr ? 2 : t0->packed;
    constexpr int VN = ai::Vec<T>::N;
    long long head = (long long)(((16 - (ax % 16)) % 16) / sizeof(T));
    if (head > n) head = n;
    const long long nvec = (n - head) / VN;
    ai::LaunchCfg cfg;
    const long long tile_elems = (long long)ai::kBlock * ai::UnrollMulti<T>::value * VN;
    const long long want = (n + tile_elems - 1) / tile_elems;
    cfg.grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)t0->sm_count));
    cfg.smem_bytes = (int)smem;
    cfg.stream = s;
    cfg.select = nullptr;
    cudaError_t e;
    if constexpr (sizeof(T) == 4) {
        if (t0->basis == ai::kCheb) e = ai::launch_multi_cheb_f32(t0->order, hdr_form, cfg, args, x, n, head, nvec);
        else if (t0->basis == ai::kLeg) e = ai::launch_multi_leg_f32(t0->order, hdr_form, cfg, args, x, n, head, nvec);
        else e = ai::launch_multi_mono_f32(t0->order, hdr_form, cfg, args, x, n, head, nvec);
    } else {
        if (t0->basis == ai::kCheb) e = ai::launch_multi_cheb_f64(t0->order, hdr_form, cfg, args, x, n, head, nvec);
        else if (t0->basis == ai::kLeg) e = ai::launch_multi_leg_f64(t0->order, hdr_form, cfg, args, x, n, head, nvec);
        else e = ai::launch_multi_mono_f64(t0->order, hdr_form, cfg, args, x, n, head, nvec);
    }
    if (e != cudaSuccess) return fail(AI_ERR_CUDA, "multi-table kernel launch failed: %s", cudaGetErrorString(e));
    if (fused_out) *fused_out = 1;
    return AI_OK;
}

template <typename T>
int index_entry(const ai_table* t, const T* x, int32_t* idx, int64_t n, void* stream) {
    if (!t) return fail(AI_ERR_INVALID, "table handle is NULL");
    if (n < 0) return fail(AI_ERR_INVALID, "n = %lld is negative", (long long)n);
    if ((sizeof(T) == 4) != (t->dtype == AI_DTYPE_F32)) return fail(AI_ERR_DTYPE, "table dtype mismatch");
    if (n == 0) return AI_OK;
    if (!x || !idx) return fail(AI_ERR_INVALID, "x or idx is NULL");
    DeviceGuard guard(t->device);
    if (!guard.ok) return fail(AI_ERR_CUDA, "cannot select device %d", t->device);
    const long long want = (n + ai::kBlock - 1) / ai::kBlock;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)t->sm_count * 2));
    if (t->mode == ai::kGlobal || t->mode == ai::kGlobalC || t->mode == ai::kSmemVC) {
        ai::index_kernel<T, ai::kGlobal><<<grid, ai::kBlock, 0, static_cast<cudaStream_t>(stream)>>>(t->view, x, idx, n);
    } else {
        if (t->aux_smem > 48 * 1024) {
            std::lock_guard<std::mutex> lock(ai::smem_attr_mutex());
            cudaFuncAttributes fa;
            AI_CUDA(cudaFuncGetAttributes(&fa, ai::index_kernel<T, ai::kSmem>));
            if (fa.maxDynamicSharedSizeBytes < t->aux_smem)
                AI_CUDA(cudaFuncSetAttribute(ai::index_kernel<T, ai::kSmem>, cudaFuncAttributeMaxDynamicSharedMemorySize, t->aux_smem));
        }
        ai::index_kernel<T, ai::kSmem><<<grid, ai::kBlock, t->aux_smem, static_cast<cudaStream_t>(stream)>>>(t->view, x, idx, n);
    }
    AI_CUDA(cudaGetLastError());
    return AI_OK;
}

template <typename T>
int sum_entry(const ai_table* t, const T* x, double* result, int64_t n, void* stream) {
    if (!t) return fail(AI_ERR_INVALID, "table handle is NULL");
    if (n < 0) return fail(AI_ERR_INVALID, "n = %lld is negative", (long long)n);
    if ((sizeof(T) == 4) != (t->dtype == AI_DTYPE_F32)) return fail(AI_ERR_DTYPE, "table dtype mismatch");
    if (!result || (n > 0 && !x)) return fail(AI_ERR_INVALID, "x or result is NULL");
    DeviceGuard guard(t->device);
    if (!guard.ok) return fail(AI_ERR_CUDA, "cannot select device %d", t->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    AI_CUDA(cudaMemsetAsync(result, 0, sizeof(double), s));
    if (n == 0) return AI_OK;
    ai::LaunchCfg cfg;
    const long long want = (n + ai::kBlock - 1) / ai::kBlock;
    cfg.grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)t->sm_count * t->blocks_per_sm));
    cfg.smem_bytes = t->aux_smem;
    cfg.stream = s;
    cfg.select = nullptr;
    cudaError_t e;
    // (a kSmemV table's image holds global-layout records: its reduction runs the L2-resident kernel)
    const bool from_global = t->mode == ai::kGlobal || t->mode == ai::kSmemV;
    if (from_global) cfg.smem_bytes = 0;
    const int m = from_global ? ai::kGlobal
                              : ((t->mode == ai::kGlobalC || t->mode == ai::kSmemVC) ? ai::kGlobalC
                                 : (t->uniform_hdr ? ai::kSmemU : (t->packed ? ai::kSmemP : ai::kSmem)));
    if constexpr (sizeof(T) == 4) {
        if (t->basis == ai::kCheb) e = ai::launch_sum_cheb_f32(t->order, m, cfg, t->view, x, result, n);
        else if (t->basis == ai::kLeg) e = ai::launch_sum_leg_f32(t->order, m, cfg, t->view, x, result, n);
        else e = ai::launch_sum_mono_f32(t->order, m, cfg, t->view, x, result, n);
    } else {
        if (t->basis == ai::kCheb) e = ai::launch_sum_cheb_f64(t->order, m, cfg, t->view, x, result, n);
        else if (t->basis == ai::kLeg) e = ai::launch_sum_leg_f64(t->order, m, cfg, t->view, x, result, n);
        else e = ai::launch_sum_mono_f64(t->order, m, cfg, t->view, x, result, n);
    }
    if (e != cudaSuccess) return fail(AI_ERR_CUDA, "sum kernel launch failed: %s", cudaGetErrorString(e));
    return AI_OK;
}

// ------------------------------------------------------------------ host-buffer pipeline
void pipe_destroy(ai_table* t);

int pipe_init_body(ai_table* t, size_t elem);

int pipe_init(ai_table* t, size_t elem) {
    if (t->pipe.ready) return AI_OK;
    const int rc = pipe_init_body(t, elem);
    if (rc) pipe_destroy(t);           // a partial pipeline (some buffers / events) is released, not leaked
    return rc;
}

int pipe_init_body(ai_table* t, size_t elem) {
    HostPipe& p = t->pipe;
    long mib = 32;                                            // MiB per direction per slot
    if (const char* e = std::getenv("AI_B200_HOST_CHUNK_MB")) { const long v = std::atol(e); if (v >= 1 && v <= 1024) mib = v; }
    if (const char* e = std::getenv("AI_B200_HOST_SLOTS")) { const long v = std::atol(e); if (v >= 2 && v <= kHostSlots) p.slots = (int)v; }
    p.chunk = (int64_t)(mib << 20) / (int64_t)elem;
    p.elem = elem;
    for (int s = 0; s < p.slots; ++s) {
        AI_CUDA(cudaMalloc(&p.dx[s], (size_t)p.chunk * elem));
        AI_CUDA(cudaMalloc(&p.dy[s], (size_t)p.chunk * elem));
        AI_CUDA(cudaEventCreateWithFlags(&p.in_done[s], cudaEventDisableTiming));
        AI_CUDA(cudaEventCreateWithFlags(&p.out_done[s], cudaEventDisableTiming));
        AI_CUDA(cudaEventCreate(&p.k_start[s]));
        AI_CUDA(cudaEventCreate(&p.k_done[s]));
    }
    AI_CUDA(cudaStreamCreateWithFlags(&p.s_in, cudaStreamNonBlocking));
    AI_CUDA(cudaStreamCreateWithFlags(&p.s_k, cudaStreamNonBlocking));
    AI_CUDA(cudaStreamCreateWithFlags(&p.s_out, cudaStreamNonBlocking));
    p.ready = true;
    return AI_OK;
}

void pipe_destroy(ai_table* t) {
    HostPipe& p = t->pipe;
    for (int s = 0; s < kHostSlots; ++s) {
        if (p.dx[s]) cudaFree(p.dx[s]);
        if (p.dy[s]) cudaFree(p.dy[s]);
        if (p.in_done[s]) cudaEventDestroy(p.in_done[s]);
        if (p.out_done[s]) cudaEventDestroy(p.out_done[s]);
        if (p.k_start[s]) cudaEventDestroy(p.k_start[s]);
        if (p.k_done[s]) cudaEventDestroy(p.k_done[s]);
        if (p.hin[s]) cudaFreeHost(p.hin[s]);
        if (p.hout[s]) cudaFreeHost(p.hout[s]);
        p.hin[s] = p.hout[s] = nullptr;
        p.dx[s] = p.dy[s] = nullptr;
        p.in_done[s] = p.out_done[s] = p.k_start[s] = p.k_done[s] = nullptr;
    }
    if (p.s_in) cudaStreamDestroy(p.s_in);
    if (p.s_k) cudaStreamDestroy(p.s_k);
    if (p.s_out) cudaStreamDestroy(p.s_out);
    p.s_in = p.s_k = p.s_out = nullptr;
    p.pool.reset();
    p.ready = false;
}

// One pass of the 3-stage pipeline over n host elements: chunk c is copied in on s_in, evaluated on
// s_k and copied out on s_out, `slots` chunks in flight.  with_kernel = false moves the same bytes
// with no kernel between the copies: the plain-copy ceiling of this host / link, measured through
// the very same code path (ai_measure_host_roundtrip).
template <typename T>
int host_pipeline_pass(ai_table* t, const T* x, T* y, int64_t n, bool with_kernel, double* ms_total) {
    HostPipe& p = t->pipe;
    const int64_t nchunks = (n + p.chunk - 1) / p.chunk;
    for (int64_t c = 0; c < nchunks; ++c) {
        const int s = (int)(c % p.slots);
        const int64_t off = c * p.chunk, m = std::min<int64_t>(p.chunk, n - off);
        if (c >= p.slots) {
            // slot reuse: its previous D2H must have drained before we overwrite dx/dy
            AI_CUDA(cudaEventSynchronize(p.out_done[s]));
            float ms = 0.f;
            AI_CUDA(cudaEventElapsedTime(&ms, p.k_start[s], p.k_done[s]));
            *ms_total += ms;
        }
        AI_CUDA(cudaMemcpyAsync(p.dx[s], x + off, (size_t)m * sizeof(T), cudaMemcpyHostToDevice, p.s_in));
        AI_CUDA(cudaEventRecord(p.in_done[s], p.s_in));
        AI_CUDA(cudaStreamWaitEvent(p.s_k, p.in_done[s], 0));
        AI_CUDA(cudaEventRecord(p.k_start[s], p.s_k));
        if (with_kernel) {
            const int rc = launch_eval_t<T>(t, static_cast<const T*>(p.dx[s]), static_cast<T*>(p.dy[s]), m, p.s_k);
            if (rc) return rc;
        }
        AI_CUDA(cudaEventRecord(p.k_done[s], p.s_k));
        AI_CUDA(cudaStreamWaitEvent(p.s_out, p.k_done[s], 0));
        AI_CUDA(cudaMemcpyAsync(y + off, p.dy[s], (size_t)m * sizeof(T), cudaMemcpyDeviceToHost, p.s_out));
        AI_CUDA(cudaEventRecord(p.out_done[s], p.s_out));
    }
    AI_CUDA(cudaStreamSynchronize(p.s_out));
    const int64_t first_pending = std::max<int64_t>(0, nchunks - p.slots);
    for (int64_t c = first_pending; c < nchunks; ++c) {
        const int s = (int)(c % p.slots);
        float ms = 0.f;
        AI_CUDA(cudaEventElapsedTime(&ms, p.k_start[s], p.k_done[s]));
        *ms_total += ms;
    }
    return AI_OK;
}

// Is a caller's host buffer something cudaMemcpyAsync can DMA from / to directly (page-locked,
// managed, or in fact device memory)?  Ordinary (pageable) memory is not: the driver then stages every
// copy through its own buffer, synchronously and from one thread -- 6.6 GB/s per direction against 44
// for pinned buffers on this pool's hosts (tools/host_pageable_bench.py).
bool directly_copyable(const void* ptr) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged || a.type == cudaMemoryTypeDevice;
}

int host_threads_default() {
    if (const char* e = std::getenv("AI_B200_HOST_THREADS")) { const long v = std::atol(e); if (v >= 1 && v <= 256) return (int)v; }
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min<unsigned>(16u, std::max<unsigned>(1u, hw));
}

// The pipeline for PAGEABLE caller buffers (what a numpy caller of the reference-style binding passes):
// the library's own page-locked bounce buffers, filled / drained by a pool of host threads, so that
// the DMA engines run asynchronously at their pinned rate while the host copies of the neighbouring
// chunks proceed.  Iteration c: [copy x chunk c -> hin  ||  copy hout -> y of chunk c - lag] on all
// threads, then enqueue chunk c (H2D, kernel, D2H); the GPU works on chunk c - 1 meanwhile.
template <typename T>
int host_pipeline_pass_staged(ai_table* t, const T* x, T* y, int64_t n, bool with_kernel, double* ms_total,
                              bool stage_in, bool stage_out) {
    HostPipe& p = t->pipe;
    const size_t cbytes = (size_t)p.chunk * sizeof(T);
    for (int s = 0; s < p.slots; ++s) {
        if (stage_in && !p.hin[s]) AI_CUDA(cudaHostAlloc(&p.hin[s], cbytes, cudaHostAllocDefault));
        if (stage_out && !p.hout[s]) AI_CUDA(cudaHostAlloc(&p.hout[s], cbytes, cudaHostAllocDefault));
    }
    if (!p.pool) {
        try {
            p.pool.reset(new CopyPool(host_threads_default()));
        } catch (const std::exception& e) {          // no exception may cross the C ABI
            return fail(AI_ERR_UNSUPPORTED, "cannot start the host copy threads: %s", e.what());
        }
    }
    const int64_t nchunks = (n + p.chunk - 1) / p.chunk;
    const int lag = p.slots - 1;                          // < slots: a slot is drained before it is refilled
    for (int64_t c = 0; c < nchunks + lag; ++c) {
        CopyPool::Job jobs[2];
        int nj = 0;
        const int s = (int)(c % p.slots);
        const int64_t off = c * p.chunk, m = (c < nchunks) ? std::min<int64_t>(p.chunk, n - off) : 0;
        if (c < nchunks) {
            if (c >= p.slots) {
                // slot reuse: everything of chunk c - slots has drained (its copy to y ran at iteration c - 1)
                AI_CUDA(cudaEventSynchronize(p.out_done[s]));
                float ms = 0.f;
                AI_CUDA(cudaEventElapsedTime(&ms, p.k_start[s], p.k_done[s]));
                *ms_total += ms;
            }
            if (stage_in) jobs[nj++] = {static_cast<char*>(p.hin[s]), reinterpret_cast<const char*>(x + off), (size_t)m * sizeof(T)};
        }
        const int64_t d = c - lag;
        if (stage_out && d >= 0 && d < nchunks) {
            const int sd = (int)(d % p.slots);
            const int64_t offd = d * p.chunk, md = std::min<int64_t>(p.chunk, n - offd);
            AI_CUDA(cudaEventSynchronize(p.out_done[sd]));
            jobs[nj++] = {reinterpret_cast<char*>(y + offd), static_cast<const char*>(p.hout[sd]), (size_t)md * sizeof(T)};
        }
        p.pool->run(jobs, nj);
        if (c < nchunks) {
            const void* src = stage_in ? p.hin[s] : static_cast<const void*>(x + off);
            void* dst = stage_out ? p.hout[s] : static_cast<void*>(y + off);
            AI_CUDA(cudaMemcpyAsync(p.dx[s], src, (size_t)m * sizeof(T), cudaMemcpyHostToDevice, p.s_in));
            AI_CUDA(cudaEventRecord(p.in_done[s], p.s_in));
            AI_CUDA(cudaStreamWaitEvent(p.s_k, p.in_done[s], 0));
            AI_CUDA(cudaEventRecord(p.k_start[s], p.s_k));
            if (with_kernel) {
                const int rc = launch_eval_t<T>(t, static_cast<const T*>(p.dx[s]), static_cast<T*>(p.dy[s]), m, p.s_k);
                if (rc) return rc;
            }
            AI_CUDA(cudaEventRecord(p.k_done[s], p.s_k));
            AI_CUDA(cudaStreamWaitEvent(p.s_out, p.k_done[s], 0));
            AI_CUDA(cudaMemcpyAsync(dst, p.dy[s], (size_t)m * sizeof(T), cudaMemcpyDeviceToHost, p.s_out));
            AI_CUDA(cudaEventRecord(p.out_done[s], p.s_out));
        }
    }
    AI_CUDA(cudaStreamSynchronize(p.s_out));
    const int64_t first_pending = std::max<int64_t>(0, nchunks - p.slots);
    for (int64_t c = first_pending; c < nchunks; ++c) {
        const int s = (int)(c % p.slots);
        float ms = 0.f;
        AI_CUDA(cudaEventElapsedTime(&ms, p.k_start[s], p.k_done[s]));
        *ms_total += ms;
    }
    return AI_OK;
}

template <typename T>
int host_entry(const ai_table* ct, const T* x, T* y, int64_t n, double* kernel_ms, bool with_kernel = true) {
    if (kernel_ms) *kernel_ms = 0.0;
    if (!ct) return fail(AI_ERR_INVALID, "table handle is NULL");
    if (n < 0) return fail(AI_ERR_INVALID, "n = %lld is negative", (long long)n);
    if ((sizeof(T) == 4) != (ct->dtype == AI_DTYPE_F32)) return fail(AI_ERR_DTYPE, "table dtype mismatch");
    if (n == 0) return AI_OK;
    if (!x || !y) return fail(AI_ERR_INVALID, "x or y is NULL");
    ai_table* t = const_cast<ai_table*>(ct);          // the pipeline scratch is mutable state
    DeviceGuard guard(t->device);
    if (!guard.ok) return fail(AI_ERR_CUDA, "cannot select device %d", t->device);
    std::lock_guard<std::mutex> lock(t->pipe.mu);
    int rc = pipe_init(t, sizeof(T));
    if (rc) return rc;
    double ms_total = 0.0;
    // pageable caller buffers go through the library's pinned bounce buffers (AI_B200_HOST_STAGING=0: never)
    const char* stg = std::getenv("AI_B200_HOST_STAGING");
    // (not for small calls: below 1 MiB the driver's own staging costs microseconds, while the first staged
    // call allocates the bounce buffers and starts the copy threads)
    const bool staging = !(stg && stg[0] == '0') && (size_t)n * sizeof(T) >= ((size_t)1 << 20);
    const bool stage_in = staging && !directly_copyable(x), stage_out = staging && !directly_copyable(y);
    if (stage_in || stage_out) rc = host_pipeline_pass_staged<T>(t, x, y, n, with_kernel, &ms_total, stage_in, stage_out);
    else rc = host_pipeline_pass<T>(t, x, y, n, with_kernel, &ms_total);
    if (rc) {
        // no copy into (or out of) the caller's buffers may still be in flight when we return
        const std::string msg = g_err;
        HostPipe& p = t->pipe;
        cudaStreamSynchronize(p.s_in); cudaStreamSynchronize(p.s_k); cudaStreamSynchronize(p.s_out);
        g_err = msg;
        return rc;
    }
    if (kernel_ms) *kernel_ms = ms_total;
    return AI_OK;
}

// ------------------------------------------------------------------ measurement kernels
template <typename T>
__global__ void __launch_bounds__(256) fma_chain_kernel(T* out, int iters, T a, T b) {
    T v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (T)(threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = ai::fma_(v[k], a, b);
    }
    T s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    if (s == (T)123456789) out[0] = s;      // never true; keeps the chain alive
}

template <typename V>
__global__ void __launch_bounds__(ai::kBlock) copy_kernel(const V* __restrict__ x, V* __restrict__ y, long long nvec) {
    constexpr long long kTile = (long long)ai::kBlock * ai::kUnroll;
    const long long full = nvec / kTile;
    for (long long tile = blockIdx.x; tile < full; tile += gridDim.x) {
        const long long base = tile * kTile + threadIdx.x;
        V in[ai::kUnroll];
#pragma unroll
        for (int u = 0; u < ai::kUnroll; ++u) in[u] = __ldcs(x + base + (long long)u * ai::kBlock);
#pragma unroll
        for (int u = 0; u < ai::kUnroll; ++u) __stcs(y + base + (long long)u * ai::kBlock, in[u]);
    }
    for (long long v = full * kTile + (long long)blockIdx.x * ai::kBlock + threadIdx.x; v < nvec;
         v += (long long)gridDim.x * ai::kBlock)
        __stcs(y + v, __ldcs(x + v));
}

}  // namespace

// =================================================================================== C ABI
extern "C" {

const char* ai_last_error(void) { return g_err.c_str(); }
const char* ai_version(void) { return "ai_b200 0.1.0 sm_100a"; }

int ai_device_count(int* count) {
    if (!count) return fail(AI_ERR_INVALID, "count is NULL");
    AI_CUDA(cudaGetDeviceCount(count));
    return AI_OK;
}

int ai_device_sm_arch(int device, int* major, int* minor, int* sm_count) {
    cudaDeviceProp prop;
    AI_CUDA(cudaGetDeviceProperties(&prop, device));
    if (major) *major = prop.major;
    if (minor) *minor = prop.minor;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    return AI_OK;
}

static ai_table* table_from_flat(int basis, int order, int dtype, int64_t n_leaves, const void* breaks,
                                 const void* coef, int device) {
    ai_table* t = new ai_table;
    t->basis = basis; t->order = order; t->dtype = dtype; t->device = device; t->L = n_leaves;
    t->breaks.resize(n_leaves + 1);
    t->coef.resize((size_t)n_leaves * (order + 1));
    if (dtype == AI_DTYPE_F32) {
        const float* b = static_cast<const float*>(breaks);
        const float* c = static_cast<const float*>(coef);
        for (int64_t k = 0; k <= n_leaves; ++k) t->breaks[k] = b[k];
        for (size_t k = 0; k < t->coef.size(); ++k) t->coef[k] = c[k];
    } else {
        const double* b = static_cast<const double*>(breaks);
        const double* c = static_cast<const double*>(coef);
        for (int64_t k = 0; k <= n_leaves; ++k) t->breaks[k] = b[k];
        for (size_t k = 0; k < t->coef.size(); ++k) t->coef[k] = c[k];
    }
    return t;
}

int ai_table_create(int basis, int order, int dtype, int64_t n_leaves, const void* breaks,
                    const void* coef, int device, ai_table_t** out) {
    if (!out) return fail(AI_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int rc = check_common(basis, order, dtype, device);
    if (rc) return rc;
    rc = check_leaf_count(n_leaves, order);
    if (rc) return rc;
    if (!breaks || !coef) return fail(AI_ERR_INVALID, "breaks or coef is NULL");
    return guarded([&]() -> int {
        ai_table* t = table_from_flat(basis, order, dtype, n_leaves, breaks, coef, device);
        const int rc2 = finish_table(t);
        if (rc2) { ai_table_destroy(t); return rc2; }
        *out = t;
        return AI_OK;
    });
}

int ai_table_plan(int basis, int order, int dtype, int64_t n_leaves, const void* breaks, const void* coef,
                  int64_t smem_cap_bytes, ai_table_info_t* info) {
    if (!info) return fail(AI_ERR_INVALID, "info is NULL");
    int rc = check_shape(basis, order, dtype);
    if (rc) return rc;
    rc = check_leaf_count(n_leaves, order);
    if (rc) return rc;
    if (!breaks || !coef) return fail(AI_ERR_INVALID, "breaks or coef is NULL");
    if (smem_cap_bytes <= 0) smem_cap_bytes = 227 * 1024;        // sm_100: opt-in shared memory per CTA
    return guarded([&]() -> int {
        std::unique_ptr<ai_table> t(table_from_flat(basis, order, dtype, n_leaves, breaks, coef, -1));
        std::vector<unsigned char> image;
        int rc2 = plan_table(t.get(), (size_t)smem_cap_bytes, &image);
        if (rc2 == AI_OK) {
            t->blocks_per_sm = 0; t->sm_count = 0;
            rc2 = ai_table_get_info(t.get(), info);
        }
        return rc2;
    });
}

int ai_table_create_from_tree(int basis, int order, int dtype, int num_levels, int layout,
                              const void* tree_1d, int64_t n_elems, int device, ai_table_t** out) {
    if (!out) return fail(AI_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int rc = check_common(basis, order, dtype, device);
    if (rc) return rc;
    if (!tree_1d || n_elems <= 0) return fail(AI_ERR_INVALID, "tree_1d is NULL or empty");
    if (layout != AI_TREE_LAYOUT_STANDARD && layout != AI_TREE_LAYOUT_TRIM)
        return fail(AI_ERR_INVALID, "unknown tree layout %d", layout);
    if (num_levels < 1) return fail(AI_ERR_INVALID, "num_levels %d < 1", num_levels);
    if (n_elems / 3 > 2 * kMaxLeaves) return fail(AI_ERR_UNSUPPORTED, "tree_1d of %lld elements is too large", (long long)n_elems);
    return guarded([&]() -> int {
        auto at = [&](int64_t k) -> double {
            return dtype == AI_DTYPE_F32 ? (double)static_cast<const float*>(tree_1d)[k]
                                         : static_cast<const double*>(tree_1d)[k];
        };
        const int n = order;
        const int stride = n + 6;
        std::vector<double> a, b, coef;
        if (layout == AI_TREE_LAYOUT_TRIM) {
            // "trim_data" (approximator.py:202-206, 232-233): nodes in preorder, a leaf is
            // [mid, self, self, c0..cn, a, b] (both links equal), an inner node is
            // [mid, left NODE index, right NODE index] (never equal).  The reference's own tree_1d walk
            // reads those node indices as element offsets (approximator.py:275-290), so the links
            // cannot be followed; but preorder visits the leaves left to right, which is all the
            // flattening needs.  The links are still checked: left child = next node, right child
            // = the node after the left subtree.
            std::vector<int64_t> pending_right;          // right-child node indices still expected
            int64_t node = 0, o = 0;
            while (o < n_elems) {
                if (o + 3 > n_elems) return fail(AI_ERR_TABLE, "trim_data tree_1d ends inside node %lld", (long long)node);
                if (!pending_right.empty() && pending_right.back() == node) pending_right.pop_back();
                const double l = at(o + 1), r = at(o + 2);
                if (l == r) {                            // leaf
                    if (o + stride > n_elems) return fail(AI_ERR_TABLE, "trim_data tree_1d ends inside leaf node %lld", (long long)node);
                    for (int j = 0; j <= n; ++j) coef.push_back(at(o + 3 + j));
                    a.push_back(at(o + n + 4));
                    b.push_back(at(o + n + 5));
                    o += stride;
                } else {
                    if ((int64_t)l != node + 1) return fail(AI_ERR_TABLE, "trim_data node %lld: left child %g is not the next node", (long long)node, l);
                    pending_right.push_back((int64_t)r);
                    o += 3;
                }
                ++node;
            }
            if (!pending_right.empty()) return fail(AI_ERR_TABLE, "trim_data tree_1d: right child %lld never reached", (long long)pending_right.back());
            if ((int64_t)a.size() * 2 - 1 != node) return fail(AI_ERR_TABLE, "trim_data tree_1d is not a full binary tree (%lld nodes, %lld leaves)", (long long)node, (long long)a.size());
        } else {
            // Iterative traversal over element offsets, following the child links the way
            // Approximator.evaluate does (approximator.py:274-290): children are ELEMENT offsets at
            // [n+4], [n+5], leaves point to themselves (approximator.py:198-199).
            if (n_elems % stride) return fail(AI_ERR_TABLE, "tree_1d length %lld is not a multiple of the node stride %d",
                                              (long long)n_elems, stride);
            const int64_t n_nodes = n_elems / stride;
            struct Frame { int64_t off; int state; };
            std::vector<Frame> stack;
            stack.push_back({0, 0});
            int64_t visited = 0;
            while (!stack.empty()) {
                Frame& fr = stack.back();
                const int64_t o = fr.off;
                if (o < 0 || o + stride > n_elems || (o % stride)) return fail(AI_ERR_TABLE, "child offset %lld out of range", (long long)o);
                const int64_t lo = (int64_t)at(o + n + 4), ro = (int64_t)at(o + n + 5);
                const bool leaf = (lo == o) || (ro == o);
                if (leaf) {
                    a.push_back(at(o + n + 2));
                    b.push_back(at(o + n + 3));
                    for (int j = 0; j <= n; ++j) coef.push_back(at(o + 1 + j));
                    stack.pop_back();
                } else if (fr.state == 0) {
                    fr.state = 1;
                    stack.push_back({lo, 0});
                } else if (fr.state == 1) {
                    fr.state = 2;
                    stack.push_back({ro, 0});
                } else {
                    stack.pop_back();
                }
                if (++visited > 4 * n_nodes + 8) return fail(AI_ERR_TABLE, "tree_1d child links form a cycle");
                if ((int64_t)stack.size() > n_nodes + 1) return fail(AI_ERR_TABLE, "tree_1d deeper than its node count");
            }
        }
        const int64_t L = (int64_t)a.size();
        if (L < 1) return fail(AI_ERR_TABLE, "tree has no leaves");
        for (int64_t k = 0; k + 1 < L; ++k)
            if (b[k] != a[k + 1]) return fail(AI_ERR_TABLE, "leaves %lld and %lld are not contiguous (%.17g != %.17g)",
                                              (long long)k, (long long)k + 1, b[k], a[k + 1]);
        int rc2 = check_leaf_count(L, order);
        if (rc2) return rc2;
        ai_table* t = new ai_table;
        t->basis = basis; t->order = order; t->dtype = dtype; t->device = device; t->L = L;
        t->breaks = a;
        t->breaks.push_back(b.back());
        t->coef = coef;
        rc2 = finish_table(t);
        if (rc2) { ai_table_destroy(t); return rc2; }
        *out = t;
        return AI_OK;
    });
}

int ai_table_destroy(ai_table_t* t) {
    if (!t) return AI_OK;
    DeviceGuard guard(t->device);
    pipe_destroy(t);
    if (t->d_image) cudaFree(t->d_image);
    delete t;
    return AI_OK;
}

int ai_table_get_info(const ai_table_t* t, ai_table_info_t* info) {
    if (!t || !info) return fail(AI_ERR_INVALID, "table or info is NULL");
    info->basis = t->basis; info->order = t->order; info->dtype = t->dtype; info->device = t->device;
    info->n_leaves = t->L;
    info->grid_cells = t->G; info->grid_log2_scale = t->p; info->search_steps = t->K;
    info->record_stride = t->rstride;
    info->image_bytes = t->view.image_bytes;
    info->smem_bytes = t->eval_smem;
    info->residency = (t->mode == ai::kSmemP || t->mode == ai::kSmemU) ? ai::kSmem
                      : ((t->mode == ai::kSmemBankP || t->mode == ai::kSmemBankU) ? ai::kSmemBank : t->mode);
    info->header_packed = t->uniform_hdr ? 2 : t->packed;
    info->cell_records = t->cell_records;
    info->bank_copies = (t->mode == ai::kSmemVC) ? ai::kVecCopies : t->copies;
    info->smem_resident = t->smem_resident;
    info->lookup_mode = t->path ? 4 : (t->rank_mode ? 2 : (t->K > 0 ? 1 : (t->rank16 ? 3 : 0)));
    info->block_threads = ai::kBlock; info->blocks_per_sm = t->blocks_per_sm; info->sm_count = t->sm_count;
    info->lower_bound = t->breaks.front(); info->upper_bound = t->breaks.back();
    return AI_OK;
}

int ai_table_get_flat(const ai_table_t* t, void* breaks, void* coef) {
    if (!t) return fail(AI_ERR_INVALID, "table is NULL");
    if (t->dtype == AI_DTYPE_F32) {
        if (breaks) for (size_t k = 0; k < t->breaks.size(); ++k) static_cast<float*>(breaks)[k] = (float)t->breaks[k];
        if (coef) for (size_t k = 0; k < t->coef.size(); ++k) static_cast<float*>(coef)[k] = (float)t->coef[k];
    } else {
        if (breaks) std::memcpy(breaks, t->breaks.data(), t->breaks.size() * sizeof(double));
        if (coef) std::memcpy(coef, t->coef.data(), t->coef.size() * sizeof(double));
    }
    return AI_OK;
}

int ai_eval_f32(const ai_table_t* t, const float* x, float* y, int64_t n, void* s) { return eval_entry<float>(t, x, y, n, s); }
int ai_eval_f64(const ai_table_t* t, const double* x, double* y, int64_t n, void* s) { return eval_entry<double>(t, x, y, n, s); }
int ai_eval_index_f32(const ai_table_t* t, const float* x, int32_t* i, int64_t n, void* s) { return index_entry<float>(t, x, i, n, s); }
int ai_eval_index_f64(const ai_table_t* t, const double* x, int32_t* i, int64_t n, void* s) { return index_entry<double>(t, x, i, n, s); }
int ai_eval_sum_f32(const ai_table_t* t, const float* x, double* r, int64_t n, void* s) { return sum_entry<float>(t, x, r, n, s); }
int ai_eval_sum_f64(const ai_table_t* t, const double* x, double* r, int64_t n, void* s) { return sum_entry<double>(t, x, r, n, s); }
int ai_eval_multi_f32(const ai_table_t* const* t, int m, const float* x, float* const* y, int64_t n, void* s, int* fused) {
    return multi_entry<float>(t, m, x, y, n, s, fused);
}
int ai_eval_multi_f64(const ai_table_t* const* t, int m, const double* x, double* const* y, int64_t n, void* s, int* fused) {
    return multi_entry<double>(t, m, x, y, n, s, fused);
}
int ai_eval_host_f32(const ai_table_t* t, const float* x, float* y, int64_t n, double* ms) { return host_entry<float>(t, x, y, n, ms); }
int ai_eval_host_f64(const ai_table_t* t, const double* x, double* y, int64_t n, double* ms) { return host_entry<double>(t, x, y, n, ms); }

int ai_measure_host_roundtrip(const ai_table_t* t, const void* x, void* y, int64_t n) {
    if (!t) return fail(AI_ERR_INVALID, "table handle is NULL");
    if (t->dtype == AI_DTYPE_F32) return host_entry<float>(t, static_cast<const float*>(x), static_cast<float*>(y), n, nullptr, false);
    return host_entry<double>(t, static_cast<const double*>(x), static_cast<double*>(y), n, nullptr, false);
}

int ai_host_alloc(void** p, int64_t bytes) {
    if (!p || bytes < 0) return fail(AI_ERR_INVALID, "bad arguments");
    AI_CUDA(cudaHostAlloc(p, (size_t)bytes, cudaHostAllocDefault));
    return AI_OK;
}

int ai_host_copy(void* dst, const void* src, int64_t bytes, int threads) {
    return guarded([&]() -> int {
        if (bytes < 0) return fail(AI_ERR_INVALID, "bytes = %lld is negative", (long long)bytes);
        if (bytes == 0) return AI_OK;
        if (!dst || !src) return fail(AI_ERR_INVALID, "dst or src is NULL");
        if (threads > 256) return fail(AI_ERR_INVALID, "threads = %d is more than 256", threads);
        CopyPool pool(threads > 0 ? threads : host_threads_default());
        const CopyPool::Job job = {static_cast<char*>(dst), static_cast<const char*>(src), (size_t)bytes};
        pool.run(&job, 1);
        return AI_OK;
    });
}

int ai_host_free(void* p) {
    if (p) AI_CUDA(cudaFreeHost(p));
    return AI_OK;
}

int ai_measure_fma_peak(int device, int dtype, double* tflops) {
    if (!tflops) return fail(AI_ERR_INVALID, "tflops is NULL");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(AI_ERR_CUDA, "cannot select device %d", device);
    cudaDeviceProp prop;
    AI_CUDA(cudaGetDeviceProperties(&prop, device));
    const int grid = prop.multiProcessorCount * 8, iters = 4096;
    void* out = nullptr;
    AI_CUDA(cudaMalloc(&out, 64));
    cudaEvent_t e0, e1;
    AI_CUDA(cudaEventCreate(&e0));
    AI_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        AI_CUDA(cudaEventRecord(e0, 0));
        if (dtype == AI_DTYPE_F32) fma_chain_kernel<float><<<grid, 256>>>(static_cast<float*>(out), iters, 1.0000001f, 1e-9f);
        else fma_chain_kernel<double><<<grid, 256>>>(static_cast<double*>(out), iters, 1.0000000001, 1e-12);
        AI_CUDA(cudaEventRecord(e1, 0));
        AI_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        AI_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    AI_CUDA(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    const double flops = 2.0 * 8.0 * (double)iters * 256.0 * (double)grid;
    *tflops = flops / ((double)best * 1e-3) / 1e12;
    return AI_OK;
}

int ai_copy_f32(const float* x, float* y, int64_t n, int device, void* stream) {
    if (n < 0 || (n > 0 && (!x || !y))) return fail(AI_ERR_INVALID, "bad arguments");
    if ((reinterpret_cast<uintptr_t>(x) % 16) || (reinterpret_cast<uintptr_t>(y) % 16) || (n % 4))
        return fail(AI_ERR_INVALID, "ai_copy_f32 needs 16-byte aligned buffers and n %% 4 == 0");
    if (n == 0) return AI_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(AI_ERR_CUDA, "cannot select device %d", device);
    int sms = 0;
    AI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    copy_kernel<float4><<<sms * 8, ai::kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), n / 4);
    AI_CUDA(cudaGetLastError());
    return AI_OK;
}

}  // extern "C"
