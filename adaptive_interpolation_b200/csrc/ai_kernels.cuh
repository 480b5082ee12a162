// ai_kernels.cuh -- device side of the B200 evaluator (sm_100a).
//
// One kernel template eval_kernel<BASIS, ORDER, T> replaces every kernel text the reference
// generates per approximation (adaptive_interpolation/generate.py: gen_cheb :62-102, gen_leg
// :106-144, gen_mono :28-57, gen_ispc :191-400).  Per point it does what that text does --
// find the leaf interval, rescale to [-1,1], evaluate the order-n polynomial -- but
//   * the interval is found with an EXACT power-of-two grid (cell = floor(x * 2^p), no rounding
//     anywhere) plus, only for tables whose breakpoints do not sit on cell edges, a fixed-trip
//     branch-free binary search over the stored breakpoints.  The result equals the reference's
//     BST walk (approximator.py:285-290) for every input, which the reference's own O(1) map
//     `f[(int)(S*x - O)]` (generate.py:301) does not;
//   * the table (grid, breakpoints, per-leaf records) is staged once per CTA into shared memory;
//   * x / y stream through HBM as 16-byte vectors with streaming cache hints.
// The work is HBM-bound integer/FP32/FP64 streaming: no tensor cores, by design.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ai {

// Tuning knobs (overridable with -D for tools/tune.py; the defaults are the measured best,
// see profiles/).
#ifndef AI_UNROLL
#define AI_UNROLL 6                // 16-byte vectors per thread per tile (float: 24 points)
#endif
#ifndef AI_MIN_BLOCKS
#define AI_MIN_BLOCKS 1            // __launch_bounds__ second argument (register cap)
#endif
#ifndef AI_PREFETCH
#define AI_PREFETCH 1              // 1: load tile k+1 before evaluating tile k; 2: tiles k+1 and k+2 in flight
#endif
#ifndef AI_UNIFORM_PATH
#define AI_UNIFORM_PATH 2          // 0: off, 1: test every tile, 2: adaptive (re-probe every 16 tiles)
#endif

#ifndef AI_PACKED_F32
#define AI_PACKED_F32 1            // fp32: two points per FFMA2 / FADD2 / FMUL2 (1: both paths, 2: leaf-uniform path only, 0: off)
#endif
// Debug build (tools/tune.py "debug_bounds"): every table access is range-checked and traps.
#ifdef AI_DEBUG_BOUNDS
#define AI_CHECK(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define AI_CHECK(cond) do { } while (0)
#endif
#ifndef AI_BLOCK
#define AI_BLOCK 512               // threads per CTA
#endif
// kSelect (two-leaf tables): how many of the LAST record words come from shared memory (128-bit
// loads) instead of a register select (0: none).  Measured on the saved order-13 tables (random
// points, % of HBM): fp64 0 / 6 / 10 / 16 words -> 83.5 / 87.4 / 86.0 / 59.7 (a select of a double
// is two instructions, a 128-bit load of two doubles is one); fp32 0 / 8 / 12 / 16 -> 74.6 / 72.2 /
// 71.6 / 58.9 (the packed FFMA2 operands are register PAIRS over two points: a vector load of one
// point's words needs moves to interleave them, a select writes the pair's half directly).
#ifndef AI_SEL_LDS_F32
#define AI_SEL_LDS_F32 0
#endif
#ifndef AI_SEL_LDS_F64
#define AI_SEL_LDS_F64 6
#endif
#ifndef AI_UNROLL_F64
#define AI_UNROLL_F64 5            // unroll for the double kernels (2 points per vector: 10 points)
#endif

constexpr int kBlock = AI_BLOCK;   // threads per CTA
constexpr int kUnroll = AI_UNROLL; // 16-byte vectors in flight per thread per tile (copy kernel)
template <typename T> struct Unroll { static constexpr int value = AI_UNROLL; };
template <> struct Unroll<double> { static constexpr int value = AI_UNROLL_F64; };

enum : int { kCheb = 0, kLeg = 1, kMono = 2 };

// Where a kernel reads the table from.
//   kSmem      image staged once per CTA into shared memory, records packed (stride 1)
//   kSmemBank  as kSmem, but every record element is replicated once per shared-memory bank
//              (32 copies of a float, 16 of a double) and lane l only ever reads copy l: gathers
//              by random leaf are conflict-free for ANY number of leaves (tables with more
//              leaves than banks otherwise serialise 2-4x on random inputs)
//   kGlobal    nothing staged: grid table, breakpoints and records are read through L1/L2 from
//              the global image (tables too large for shared memory)
//   kSmemBankN as kSmemBank with a run-time power-of-two copy count < 32 (16): tables whose full
//              replication does not fit in shared memory (lanes l mod N share a copy)
//   kSelect    tables with at most TWO leaves (the saved order-13 tables): nothing staged and no
//              gather at all.  Both records travel in the kernel parameter space (constant
//              bank), the leaf is the predicate !(x < mid) and every record element is a
//              register/constant SELECT, which runs on the ALU pipe instead of the 128 B/clk
//              shared-memory return path that bounds random points at high order
//              (tools/micro/shfl_bench.cu: LDS and SHFL share that path, LDC does not but
//              serialises per distinct address).
// In both bank modes the breakpoint array read by the search is replicated the same way.
//   kSmemP / kSmemBankP  as kSmem / kSmemBank with a ONE-element record header: dyadic tables (every
//              table the reference's bisection builds on a domain with few significant bits) have
//              a_i with >= 9 (fp32) / 32 (fp64) trailing zero mantissa bits and 2/(b_i-a_i) =
//              S0 * 2^-j_i exactly, so [a, 2/(b-a)] packs into one word (a's bits | j) that two
//              integer instructions unpack.  One shared-memory wavefront less per 32 points on the
//              return path that bounds random points (fp64: two less); the unpacked values are
//              the stored ones bit for bit (verified per leaf on the host, else not packed).
//   kSmemU / kSmemBankU  (the bank form for fp64 only) as kSmem / kSmemBank with NO record header at all: tables whose
//              leaves all have one width (one leaf per coarse cell: the 8-leaf order-13 / order-7
//              and the 64- / 128-leaf 2.2e-14 tables) have a_c = fma(c, w, lb) and one constant
//              2/(b-a); the kernel computes both from the cell index (I2F + FMA) instead of
//              gathering them: two more wavefronts off the return path.  Verified per cell on the
//              host against the stored header, bit for bit.
//   kSmemV     tables whose full per-bank replication does not fit in shared memory (hundreds of fp64
//              leaves): records in the 16-byte padded global layout, EIGHT copies, fetched with
//              128-bit loads.  A 128-bit shared load is served a quarter-warp (8 lanes) at a time, so
//              8 copies -- lane l reads copy l mod 8, each copy in its own 16-byte bank group -- make
//              the gather conflict-free with a quarter (fp32) / half (fp64) of the memory the
//              per-bank layouts need.  Breakpoints (read by the search, if any) are replicated
//              per bank when they fit beside the records.
//   kGlobalC   dyadic tables too large for shared memory (the reference's order-3 / 2.2e-14 fits: 10^4
//              leaves): a random point costs one L1 tag lookup per cache line per load instruction,
//              so what counts is how many load instructions touch the table per point.  Records are
//              stored once per COARSE CELL with NO header, padded to 32 bytes, and fetched with
//              256-bit loads (LDG.E.256: an order-3 fp64 record is ONE instruction and one sector);
//              the header is computed: a 4-bit level k per cell, staged in shared memory, gives the
//              leaf's first cell (cell with its low k bits cleared), a = fma(first cell, w, lb) and
//              2/(b-a) = S1 * 2^-k -- all verified bit for bit on the host.  No grid table, no
//              breakpoints and no header words are read: one global access per point instead of 5-6.
//   kSmemVC    kSmemV for DYADIC tables: the headerless cell records of kGlobalC ([c0..cn] padded to 16
//              bytes, one per coarse cell) in eight 128-bit copies, the 4-bit levels beside them;
//              header computed as in kGlobalC.  An order-5 fp64 record is 3 loads instead of 4 and
//              the cell index addresses it directly (no rank lookup).  The global image IS the
//              kGlobalC image (the index and reduction kernels read it as such).
enum : int { kSmem = 0, kSmemBank = 1, kGlobal = 2, kSmemBankN = 3, kSelect = 4, kSmemP = 5, kSmemBankP = 6,
             kSmemU = 7, kSmemBankU = 8, kSmemV = 9, kGlobalC = 10, kSmemVC = 11 };
constexpr int kVecCopies = 8;
__host__ __device__ constexpr bool mode_packed_hdr(int mode) { return mode == kSmemP || mode == kSmemBankP; }
__host__ __device__ constexpr bool mode_uniform_hdr(int mode) { return mode == kSmemU || mode == kSmemBankU; }
__host__ __device__ constexpr bool mode_bank(int mode) {
    return mode == kSmemBank || mode == kSmemBankN || mode == kSmemBankP || mode == kSmemBankU || mode == kSmemV;
}
constexpr int kMaxRecord = 16 + 1 + 2;         // AI_MAX_ORDER + 1 coefficients + [a, 2/(b-a)]
// kSelect's table, passed by value as a kernel parameter; empty in every other mode
template <typename T, int MODE> struct SelectTable {};
template <typename T> struct SelectTable<T, kSelect> {
    T mid;                          // breaks[1]; +inf for a single-leaf table
    T rec[2][kMaxRecord];           // [a, 2/(b-a), c0..cn] ([c0..cn] for the monomial basis)
};
template <typename T, int MODE> struct Copies {
    static constexpr int value =
        (MODE == kSmemBank || MODE == kSmemBankP || MODE == kSmemBankU) ? (sizeof(T) == 4 ? 32 : 16) : 1;
};

// Everything a kernel needs to know about a table; passed by value (kernel parameter space).
struct TableView {
    const unsigned char* image;    // packed image in global memory (16-byte aligned)
    int image_bytes;               // multiple of 16
    int off_first;                 // packed first[G]       : leaf holding the left edge of cell g
    int off_breaks;                // T breaks[L+1+pad]     : sorted breakpoints, +inf padded
    int off_records;               // T rec[L][R]           : [a, 2/(b-a), c0..cn, pad]
    int n_leaves;                  // L
    int n_records_elems;           // L * R (elements in the packed record array)
    int n_breaks_elems;            // L + 1 + padding (elements in the breakpoint array)
    int copies;                    // bank modes: copies per record element (kSmemV: per 16-byte chunk)
    int bcopies;                   // bank modes: copies per breakpoint element (a power of two)
    int soff_breaks, soff_records; // bank modes: byte offsets of the replicated sections in shared memory
    int g0;                        // cell id of the first cell: g = floor(x*2^p) - g0
    int kstep;                     // 2^(K-1), or 0 when the grid alone is exact
    int first_shift;               // first[] packing: 2 -> 8-bit entries (4 per word), 1 -> 16-bit, 0 -> 32-bit;
                                   // 3 -> no first[]: the section holds the rank words of rank_mode bit 2
    int rank_mode;                 // bit 0: leaf = rank in cont_mask of the coarse cell (no table read);
                                   // bit 1: rec[] holds one record per COARSE CELL (n_records of them);
                                   // bit 2: leaf = rank from one 32-bit word per 16 coarse cells
                                   //        (prefix count << 16 | continuation bits), see Lookup::rank16_guess
                                   // bit 3: first[] is indexed by the COARSE cell (kGlobalC tables)
    int n_records;                 // records in rec[]: L, or the coarse cell count in rank_mode 2
    unsigned div_mul, div_shift;   // coarse cell = (fine cell * div_mul) >> div_shift  (exact / q)
    unsigned div_add;              // -g0 * div_mul (mod 2^32): folds the "- g0" into the multiply
    unsigned cont_mask;            // bit c set: coarse cell c continues the leaf of cell c-1 (<= 32 cells)
    double scale;                  // 2^p (exact in T)
    double lo, hi;                 // g0 and g0+G-1 as floating point (exact in T)
    double inv_order;              // T(1./n) widened: the Legendre-like recurrence's divisor
    unsigned s0_hi, s0_lo;         // packed header: bits of S0 = max_i 2/(b_i-a_i) (fp32: s0_hi only)
    double u_w, u_lb, u_s;         // computed header (kSmemU / kSmemBankU): a_c = fma(c, u_w, u_lb), 2/(b-a) = u_s
                                   // (kGlobalC: u_w = coarse cell width, s0_hi / s0_lo = bits of S1 = 2 / u_w)
    int off_aux, aux_bytes;        // kGlobalC: 4-bit level per coarse cell (two per byte), staged in shared memory
    unsigned wide_mul, wide_shift; // kGlobalC: coarse cell = ((fine cell - g0) * wide_mul) >> wide_shift, 64-bit product
    // rank_mode bit 4, the PATH lookup of deep skewed trees (Lookup::path_leaf): first[] holds one
    // target word per row followed by the rows of leaf ordinals
    unsigned long long path_mul;   // floor(u / q) = mulhi(u, path_mul) >> path_shift; 0: q = 1
    unsigned path_shift;
    unsigned path_cshift;          // row = V >> path_cshift
    int path_rowlen;               // entries per row
    int path_zbias;                // entry = row * path_rowlen + clz(V ^ target[row]) - path_zbias
    int path_toff;                 // byte offset of the rows inside first[]
    int path_t16;                  // rows hold 16-bit leaf ordinals (more than 256 leaves), else bytes
    // per LAUNCH (set by launch_eval_t on its copy of the view): x + head is not 16-byte aligned although
    // y + head is (x and y can never be aligned together, e.g. x is a view that starts one element into
    // its buffer): the tile's x comes from element loads instead of 128-bit ones
    int x_elem_loads;
};

// ---------------------------------------------------------------------------------- helpers
template <typename T> struct Vec;
template <> struct Vec<float>  { using type = float4;  static constexpr int N = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };

template <typename T> __host__ __device__ constexpr int rec_header(int basis) { return basis == kMono ? 0 : 2; }
// record stride in elements.  Shared-memory layouts: odd, so that lanes on different leaves hit
// different banks.  Global (L2-resident) layout: padded to 16 bytes so that a record is fetched
// with 128-bit loads (a random-leaf gather costs one L1 tag lookup per line per instruction).
__host__ __device__ constexpr int rec_stride(int basis, int order) {
    return ((order + 1 + (basis == kMono ? 0 : 2)) | 1);
}
__host__ __device__ constexpr int rec_stride_global(int basis, int order, int elem_bytes) {
    const int per = 16 / elem_bytes;
    return (order + 1 + (basis == kMono ? 0 : 2) + per - 1) / per * per;
}
// kGlobalC: [c0..cn] padded to 32 bytes
__host__ __device__ constexpr int rec_stride_compact(int order, int elem_bytes) {
    const int per = 32 / elem_bytes;
    return (order + 1 + per - 1) / per * per;
}
// kSmemVC: [c0..cn] padded to 16 bytes (shared-memory copy of the kGlobalC records)
__host__ __device__ constexpr int rec_stride_vc(int order, int elem_bytes) {
    const int per = 16 / elem_bytes;
    return (order + 1 + per - 1) / per * per;
}
// packed header (kSmemP / kSmemBankP): [a|j, c0..cn]
__host__ __device__ constexpr int rec_stride_packed(int order) { return ((order + 2) | 1); }
template <int BASIS, int ORDER, typename T, int MODE>
__host__ __device__ constexpr int rec_stride_for() {
    return MODE == kSmemVC ? rec_stride_vc(ORDER, (int)sizeof(T))
           : MODE == kGlobalC ? rec_stride_compact(ORDER, (int)sizeof(T))
           : (MODE == kGlobal || MODE == kSmemV) ? rec_stride_global(BASIS, ORDER, (int)sizeof(T))
                           : (mode_packed_hdr(MODE) ? rec_stride_packed(ORDER)
                                                    : (mode_uniform_hdr(MODE) ? rec_stride(kMono, ORDER) : rec_stride(BASIS, ORDER)));
}

__device__ __forceinline__ float  fma_(float a, float b, float c)    { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }

// cell index relative to the first cell, clamped to the grid.  x*scale is exact (power of two),
// floor is exact, so cell(x) is the true floor(x*2^p) for every finite x.  NaN lands on a valid
// cell (F2I(NaN) = 0 before the integer clamp / fminf(NaN, hi) = hi).
__device__ __forceinline__ int cell_of(float x, float scale, float lo, float hi, int g0) {
    float u = x * scale;
    u = fmaxf(fminf(u, hi), lo);
    return __float2int_rd(u) - g0;
}
__device__ __forceinline__ int cell_of(double x, double scale, int glo, int ghi) {
    // cvt.rmi.s32.f64 saturates; clamp in the integer domain BEFORE subtracting g0 (no overflow)
    const int gi = __double2int_rd(x * scale);
    return min(max(gi, glo), ghi) - glo;
}

template <typename T> struct Lookup {
    T scale, lo, hi;
    int g0, ghi, kstep, lmax;
    int nbreaks;                   // elements in breaks[] incl. +inf padding (bounds checks)
    int fshift, rank_mode;
    unsigned div_mul, div_shift, div_add, cont;
    unsigned wide_mul, wide_shift;
    unsigned long long path_mul;
    unsigned path_shift, path_cshift;
    int path_rowlen, path_zbias, path_t16;
    const unsigned char* path_rows;
    const unsigned char* first;    // packed 8-bit (L <= 256) or 16-bit entries
    const unsigned char* bbase;    // breaks[] base (uniform); element j of this lane lives at
    int bunit_log2;                //   bbase + (j << bunit_log2) + blane
    int blane;                     // bank modes: byte offset of this lane's copy (< 1 << bunit_log2)

    __device__ __forceinline__ void init(const TableView& tv, const unsigned char* base) {
        scale = (T)tv.scale; lo = (T)tv.lo; hi = (T)tv.hi;
        g0 = tv.g0; ghi = (int)tv.hi; kstep = tv.kstep; lmax = tv.n_leaves - 1;
        fshift = tv.first_shift; rank_mode = tv.rank_mode; nbreaks = tv.n_breaks_elems;
        div_mul = tv.div_mul; div_shift = tv.div_shift; div_add = tv.div_add; cont = tv.cont_mask;
        wide_mul = tv.wide_mul; wide_shift = tv.wide_shift;
        path_mul = tv.path_mul; path_shift = tv.path_shift; path_cshift = tv.path_cshift;
        path_rowlen = tv.path_rowlen; path_zbias = tv.path_zbias; path_t16 = tv.path_t16;
        path_rows = base + tv.off_first + tv.path_toff;
        first = base + tv.off_first;
        bbase = base + tv.off_breaks;
        bunit_log2 = (sizeof(T) == 4) ? 2 : 3;
        blane = 0;
    }
    // bank modes: breaks[] lives at soff_breaks, `copies` copies per element, lane l reads copy l
    __device__ __forceinline__ void use_replicated_breaks(const TableView& tv, const unsigned char* base) {
        bbase = base + tv.soff_breaks;
        bunit_log2 = ((sizeof(T) == 4) ? 2 : 3) + (31 - __clz(tv.bcopies));
        blane = (int)(threadIdx.x & (tv.bcopies - 1)) * (int)sizeof(T);
    }
    // fine cell of x relative to the first cell, clamped to [0, G)
    __device__ __forceinline__ int cell(T x) const {
        if constexpr (sizeof(T) == 4) return cell_of(x, scale, lo, hi, g0);
        else return cell_of(x, scale, g0, ghi);
    }
    // leaf holding the left edge of fine cell g, from the packed table.  Entries are bytes when
    // L <= 256 (80 cells = 20 words: conflict-free), else 16-bit.
    template <typename E>
    __device__ __forceinline__ int table_guess(int g) const {
        AI_CHECK(g >= 0 && g <= ghi - g0);
        const int i = (int)reinterpret_cast<const E*>(first)[g];
        AI_CHECK(i >= 0 && i <= lmax);
        return i;
    }
    // absolute fine cell of x, clamped to [g0, g0+G).  cvt.rmi.s32 saturates (NaN -> 0), so the
    // clamp can be done after the conversion in the integer domain (one 3-input min/max).
    __device__ __forceinline__ unsigned cell_abs(T x) const {
        int gi;
        if constexpr (sizeof(T) == 4) gi = __float2int_rd(x * scale);
        else gi = __double2int_rd(x * scale);
        return (unsigned)min(max(gi, g0), ghi);
    }
    // dyadic tables with <= 32 coarse cells: no memory access at all.  The coarse cell is the
    // fine cell divided by the odd factor q of the leaf width (exact multiply-shift, verified on
    // the host for every fine cell); the leaf is the number of leaf starts at or before it.
    __device__ __forceinline__ unsigned coarse_cell(T x) const {
        return (cell_abs(x) * div_mul + div_add) >> div_shift;
    }
    // kGlobalC tables (grids of up to 2^20 fine cells): the same exact division with a 64-bit product
    __device__ __forceinline__ unsigned coarse_cell_wide(T x) const {
        return (unsigned)(((unsigned long long)(cell_abs(x) - (unsigned)g0) * wide_mul) >> wide_shift);
    }
    __device__ __forceinline__ int rank_guess(T x) const {
        const unsigned c = coarse_cell(x);
        AI_CHECK(c < 32u);
        const unsigned above = 0xfffffffeu << c;                 // bits strictly above c
        const int i = (int)c - __popc(cont & ~above);
        AI_CHECK(i >= 0 && i <= lmax);
        return i;
    }
    // Dyadic tables with more than 32 coarse cells: the same rank from memory, ONE 32-bit word per
    // 16 coarse cells = (continuation bits set in all earlier words) << 16 | this word's 16
    // continuation bits.  Up to 512 cells the words sit in 32 different banks (lanes on the same
    // word are a broadcast): one conflict-free wavefront, where random reads of a 16-bit first[]
    // table of 512 entries cost ~3.5.  Verified on the host for every fine cell.
    __device__ __forceinline__ int rank16_guess(T x) const {
        const unsigned c = coarse_cell(x);
        const unsigned w = reinterpret_cast<const unsigned*>(first)[c >> 4];
        const unsigned upto = (2u << (c & 15u)) - 1u;            // bits 0..(c & 15) of this word
        const int i = (int)c - (int)(w >> 16) - __popc(w & upto & 0xffffu);
        AI_CHECK(i >= 0 && i <= lmax);
        return i;
    }
    // Deep skewed trees (BASELINE config 4: bisection towards a discontinuity, depth 43): the exact
    // grid would need 2^45 cells and a coarser grid leaves a K-step search (K dependent loads of a
    // breakpoint, each with a compare and a select).  Such a tree is a PATH below some level Lc:
    // every split has one child that is a leaf.  With V = the point's position in units of the
    // narrowest leaf (floor(x * 2^p) is exact; the odd factor q of the domain width divides out by
    // an exact multiply-high), the leaf is decided by the first bit in which V leaves the path,
    // i.e. by clz(V ^ target) -- one load of the row's target word and one of the leaf ordinal,
    // whatever the depth.  The host checks the formula on both sides of every breakpoint.
    template <bool DIV>
    __device__ __forceinline__ int path_leaf(T x) const {
        int e;
        if constexpr (sizeof(T) == 4) {
            unsigned v = cell_abs(x) - (unsigned)g0;
            if constexpr (DIV) v = __umulhi(v, (unsigned)path_mul) >> path_shift;
            const unsigned row = v >> path_cshift;
            const unsigned tgt = reinterpret_cast<const unsigned*>(first)[row];
            e = (int)row * path_rowlen + __clz((int)(v ^ tgt)) - path_zbias;
        } else {
            // clamp as a double (lo, hi and every cell id are exact below 2^52; NaN -> lo)
            const double u = fmin(fmax(floor(x * scale), lo), hi) - lo;
            unsigned long long v = (unsigned long long)__double2ll_rz(u);
            if constexpr (DIV) v = __umul64hi(v, path_mul) >> path_shift;
            const unsigned row = (unsigned)(v >> path_cshift);
            const unsigned long long tgt = reinterpret_cast<const unsigned long long*>(first)[row];
            e = (int)row * path_rowlen + __clzll((long long)(v ^ tgt)) - path_zbias;
        }
        AI_CHECK(e >= 0);
        const int i = path_t16 ? (int)reinterpret_cast<const unsigned short*>(path_rows)[e]
                               : (int)path_rows[e];
        AI_CHECK(i >= 0 && i <= lmax);
        return i;
    }
    // RECORD index of N points held in registers: the leaf ordinal, except for tables whose
    // records are stored once per coarse cell (rank_mode 2: at most 32 / 16 cells), where the
    // coarse cell itself addresses the record and the rank (shift, mask, popc, subtract) is not
    // computed at all.  Mode branches are warp-uniform and sit OUTSIDE the per-point loops, so the
    // N dependent chains interleave (ILP).
    template <int N>
    __device__ __forceinline__ void leaves(const T (&x)[N], int (&i)[N]) const {
        if (rank_mode & 2) {
#pragma unroll
            for (int e = 0; e < N; ++e) i[e] = (int)coarse_cell(x[e]);
            return;
        }
        true_leaves<N>(x, i);
    }
    // the LEAF ordinal (what the index kernel reports), whatever the record layout
    template <int N>
    __device__ __forceinline__ void true_leaves(const T (&x)[N], int (&i)[N]) const {
        if (rank_mode & 16) {
            if (path_mul) {
#pragma unroll
                for (int e = 0; e < N; ++e) i[e] = path_leaf<true>(x[e]);
            } else {
#pragma unroll
                for (int e = 0; e < N; ++e) i[e] = path_leaf<false>(x[e]);
            }
            return;
        }
        if (rank_mode & 1) {
#pragma unroll
            for (int e = 0; e < N; ++e) i[e] = rank_guess(x[e]);
            return;
        }
        if (rank_mode & 4) {
#pragma unroll
            for (int e = 0; e < N; ++e) i[e] = rank16_guess(x[e]);
            return;
        }
        if (rank_mode & 8) {                                     // first[] by coarse cell
            if (fshift == 1) {
#pragma unroll
                for (int e = 0; e < N; ++e) i[e] = (int)reinterpret_cast<const unsigned short*>(first)[coarse_cell_wide(x[e])];
            } else {
#pragma unroll
                for (int e = 0; e < N; ++e) i[e] = (int)reinterpret_cast<const unsigned int*>(first)[coarse_cell_wide(x[e])];
            }
            return;
        }
        if (fshift == 2) {
#pragma unroll
            for (int e = 0; e < N; ++e) i[e] = table_guess<unsigned char>(cell(x[e]));
        } else if (fshift == 1) {
#pragma unroll
            for (int e = 0; e < N; ++e) i[e] = table_guess<unsigned short>(cell(x[e]));
        } else {
#pragma unroll
            for (int e = 0; e < N; ++e) i[e] = table_guess<unsigned int>(cell(x[e]));
        }
        if (kstep) {
            // Tables whose breakpoints are off the grid: branch-free forward binary search over
            // the stored breakpoints (+inf padded past L).  The candidate is carried as a BYTE
            // offset into breaks[] so that a step is add / load / compare / select.
            const int unit = 1 << bunit_log2;
#pragma unroll
            for (int e = 0; e < N; ++e) i[e] = (i[e] << bunit_log2) + blane;
            for (int sb = kstep << bunit_log2; sb >= unit; sb >>= 1) {
#pragma unroll
                for (int e = 0; e < N; ++e) {
                    const int q = i[e] + sb;
                    AI_CHECK(q >= 0 && (q >> bunit_log2) < nbreaks);
                    const T b = *reinterpret_cast<const T*>(bbase + q);
                    i[e] = !(x[e] < b) ? q : i[e];
                }
            }
#pragma unroll
            for (int e = 0; e < N; ++e) i[e] = min(i[e] >> bunit_log2, lmax);   // drops blane (< unit)
        }
    }
    __device__ __forceinline__ int leaf(T x) const {
        T xs[1] = {x};
        int is[1];
        leaves<1>(xs, is);
        return is[0];
    }
    // as leaf(), but also reproduces the reference for NaN: every `mid > x` test is false, the
    // walk always goes right and ends on the last leaf (approximator.py:287-290).
    __device__ __forceinline__ int leaf_nan_exact(T x) const {
        T xs[1] = {x};
        int is[1];
        true_leaves<1>(xs, is);                                  // the LEAF, also when records are per cell
        return (x != x) ? lmax : is[0];
    }
};

// ------------------------------------------------------------------------- basis evaluation
// rec points at the leaf record in shared memory.  Arithmetic follows the reference's operation
// order (adapt.py:97-139 as called from approximator.py:291-295) with each multiply-add pair
// fused; see DESIGN.md "numerics" for the measured deviation from the reference evaluator.
template <int BASIS, int ORDER, typename T> struct Record {
    static constexpr int HDR = (BASIS == kMono) ? 0 : 2;
    static constexpr int N = ORDER + 1 + HDR;
    T v[N];                                    // [a, 2/(b-a), c0..cn]  (monomial: [c0..cn])
    template <int STRIDE = 1>
    __device__ __forceinline__ void load(const T* __restrict__ rec) {
#pragma unroll
        for (int k = 0; k < N; ++k) v[k] = rec[k * STRIDE];
    }
    // packed header: rec = [w, c0..cn] with w = (bits of a) | j and 2/(b-a) = S0 * 2^-j.
    // fp32: a = w & ~0x1ff, bits(2/(b-a)) = bits(S0) - (j << 23) (one IMAD: the multiply by
    // -2^23 keeps only w's low 9 bits).  fp64: high word = a's high word (its low word is 0),
    // low word = j, bits(2/(b-a)) = bits(S0) - (j << 52).
    template <int STRIDE = 1>
    __device__ __forceinline__ void load_packed(const T* __restrict__ rec, unsigned s0_hi, unsigned s0_lo) {
        static_assert(HDR == 2, "the monomial basis has no header");
        const T w = rec[0];
#pragma unroll
        for (int k = 0; k <= ORDER; ++k) v[2 + k] = rec[(1 + k) * STRIDE];
        unpack_header(w, s0_hi, s0_lo);
    }
    // computed header: rec = [c0..cn]; a and 2/(b-a) from the cell index
    template <int STRIDE = 1>
    __device__ __forceinline__ void load_uniform(const T* __restrict__ rec, int cell, T w, T lb, T s) {
        static_assert(HDR == 2, "the monomial basis has no header");
#pragma unroll
        for (int k = 0; k <= ORDER; ++k) v[2 + k] = rec[k * STRIDE];
        v[0] = fma_((T)cell, w, lb);
        v[1] = s;
    }
    // run-time element stride (kSmemBankN)
    __device__ __forceinline__ void load_packed_strided(const T* __restrict__ rec, int stride, unsigned s0_hi,
                                                        unsigned s0_lo) {
        const T w = rec[0];
#pragma unroll
        for (int k = 0; k <= ORDER; ++k) v[2 + k] = rec[(1 + k) * stride];
        unpack_header(w, s0_hi, s0_lo);
    }
    __device__ __forceinline__ void unpack_header(T w, unsigned s0_hi, unsigned s0_lo) {
        if constexpr (sizeof(T) == 4) {
            const unsigned u = __float_as_uint(w);
            v[0] = __uint_as_float(u & 0xFFFFFE00u);
            v[1] = __uint_as_float(u * 0xFF800000u + s0_hi);
        } else {
            const int hi = __double2hiint(w), lo = __double2loint(w);
            v[0] = __hiloint2double(hi, 0);
            v[1] = __hiloint2double((int)(s0_hi - ((unsigned)lo << 20)), (int)s0_lo);
        }
    }
    __device__ __forceinline__ void load_strided(const T* __restrict__ rec, int stride) {
#pragma unroll
        for (int k = 0; k < N; ++k) v[k] = rec[k * stride];
    }
    // kSmemV: 16-byte chunk q of this lane's copy sits kVecCopies chunks after chunk q - 1
    __device__ __forceinline__ void load_vec_smem(const T* __restrict__ rec) {
        constexpr int PER = 16 / (int)sizeof(T);
        constexpr int NV = (N + PER - 1) / PER;
        const int4* p = reinterpret_cast<const int4*>(rec);
        T buf[NV * PER];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int4 w = p[q * kVecCopies];
            if constexpr (sizeof(T) == 4) {
                buf[4 * q] = __int_as_float(w.x); buf[4 * q + 1] = __int_as_float(w.y);
                buf[4 * q + 2] = __int_as_float(w.z); buf[4 * q + 3] = __int_as_float(w.w);
            } else {
                buf[2 * q] = __hiloint2double(w.y, w.x); buf[2 * q + 1] = __hiloint2double(w.w, w.z);
            }
        }
#pragma unroll
        for (int k = 0; k < N; ++k) v[k] = buf[k];
    }
    // kGlobalC: [c0..cn] in global memory, 32-byte aligned: 256-bit read-only loads (LDG.E.256).
    // Header from the coarse cell: level k -> first cell of the leaf -> a; 2/(b-a) = S1 * 2^-k.
    __device__ __forceinline__ void load_compact(const T* __restrict__ rec, unsigned cell, unsigned k, T w, T lb,
                                                 unsigned s1_hi, unsigned s1_lo) {
        constexpr int C = ORDER + 1;
        constexpr int PER = 32 / (int)sizeof(T);
        constexpr int NV = (C + PER - 1) / PER;
        T buf[NV * PER];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            if constexpr (sizeof(T) == 8) {
                asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                             : "=d"(buf[4 * q]), "=d"(buf[4 * q + 1]), "=d"(buf[4 * q + 2]), "=d"(buf[4 * q + 3])
                             : "l"(rec + 4 * q));
            } else {
                asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=f"(buf[8 * q]), "=f"(buf[8 * q + 1]), "=f"(buf[8 * q + 2]), "=f"(buf[8 * q + 3]),
                               "=f"(buf[8 * q + 4]), "=f"(buf[8 * q + 5]), "=f"(buf[8 * q + 6]), "=f"(buf[8 * q + 7])
                             : "l"(rec + 8 * q));
            }
        }
#pragma unroll
        for (int j = 0; j < C; ++j) v[HDR + j] = buf[j];
        if constexpr (HDR == 2) {
            const unsigned first_cell = cell & (0xffffffffu << k);
            v[0] = fma_((T)(int)first_cell, w, lb);
            if constexpr (sizeof(T) == 4) v[1] = __uint_as_float(s1_hi - (k << 23));
            else v[1] = __hiloint2double((int)(s1_hi - (k << 20)), (int)s1_lo);
        }
    }
    // kSmemVC: the same headerless record from this lane's shared-memory copy (chunk q sits
    // kVecCopies chunks after chunk q - 1), header computed as in load_compact
    __device__ __forceinline__ void load_compact_smem(const T* __restrict__ rec, unsigned cell, unsigned k, T w, T lb,
                                                      unsigned s1_hi, unsigned s1_lo) {
        constexpr int C = ORDER + 1;
        constexpr int PER = 16 / (int)sizeof(T);
        constexpr int NV = (C + PER - 1) / PER;
        const int4* p = reinterpret_cast<const int4*>(rec);
        T buf[NV * PER];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int4 v4 = p[q * kVecCopies];
            if constexpr (sizeof(T) == 4) {
                buf[4 * q] = __int_as_float(v4.x); buf[4 * q + 1] = __int_as_float(v4.y);
                buf[4 * q + 2] = __int_as_float(v4.z); buf[4 * q + 3] = __int_as_float(v4.w);
            } else {
                buf[2 * q] = __hiloint2double(v4.y, v4.x); buf[2 * q + 1] = __hiloint2double(v4.w, v4.z);
            }
        }
#pragma unroll
        for (int j = 0; j < C; ++j) v[HDR + j] = buf[j];
        if constexpr (HDR == 2) {
            const unsigned first_cell = cell & (0xffffffffu << k);
            v[0] = fma_((T)(int)first_cell, w, lb);
            if constexpr (sizeof(T) == 4) v[1] = __uint_as_float(s1_hi - (k << 23));
            else v[1] = __hiloint2double((int)(s1_hi - (k << 20)), (int)s1_lo);
        }
    }
    // 16-byte aligned record in global memory: 128-bit read-only loads
    __device__ __forceinline__ void load_vec(const T* __restrict__ rec) {
        constexpr int PER = 16 / (int)sizeof(T);
        constexpr int NV = (N + PER - 1) / PER;
        const int4* p = reinterpret_cast<const int4*>(rec);
        T buf[NV * PER];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int4 w = __ldg(p + q);
            if constexpr (sizeof(T) == 4) {
                buf[4 * q] = __int_as_float(w.x); buf[4 * q + 1] = __int_as_float(w.y);
                buf[4 * q + 2] = __int_as_float(w.z); buf[4 * q + 3] = __int_as_float(w.w);
            } else {
                buf[2 * q] = __hiloint2double(w.y, w.x); buf[2 * q + 1] = __hiloint2double(w.w, w.z);
            }
        }
#pragma unroll
        for (int k = 0; k < N; ++k) v[k] = buf[k];
    }
};

// Arithmetic on one value (float / double) or on a PAIR of floats held in a 64-bit register pair.
// sm_100 issues packed FP32 operations (FFMA2 / FADD2 / FMUL2: two IEEE operations per issued
// instruction, operand negation folded in), which halves the FP issue slots of the fp32 kernels;
// every lane result is the same IEEE operation as the scalar form, so the bits are identical.
template <typename V> struct Ops;
template <> struct Ops<float> {
    using S = float;
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float neg(float a) { return -a; }
    static __device__ __forceinline__ float splat(float s) { return s; }
};
template <> struct Ops<double> {
    using S = double;
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double neg(double a) { return -a; }
    static __device__ __forceinline__ double splat(double s) { return s; }
};
template <> struct Ops<float2> {
    using S = float;
    static __device__ __forceinline__ float2 fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
    static __device__ __forceinline__ float2 add(float2 a, float2 b) { return __fadd2_rn(a, b); }
    static __device__ __forceinline__ float2 mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
    static __device__ __forceinline__ float2 neg(float2 a) { return make_float2(-a.x, -a.y); }
    static __device__ __forceinline__ float2 splat(float s) { return make_float2(s, s); }
};

// v holds one leaf record per lane of V: [a, 2/(b-a), c0..cn] ([c0..cn] for the monomial basis).
// Arithmetic follows the reference's definitions (adapt.py:97-139 as called from
// approximator.py:291-295) with each multiply-add pair fused; see DESIGN.md "numerics" for the
// measured deviation from the reference evaluator.
template <int BASIS, int ORDER, typename V>
__device__ __forceinline__ V eval_values(const V* __restrict__ v, V x) {
    using O = Ops<V>;
    using S = typename O::S;
    if constexpr (BASIS == kMono) {
        // adapt.py:139 sum c_i x^i on the RAW x; Horner as generate.py:48-52
        V acc = v[ORDER];
#pragma unroll
        for (int j = ORDER - 1; j >= 0; --j) acc = O::fma(acc, x, v[j]);
        return acc;
    } else {
        const V* c = v + 2;
        // adapt.py:130  xscaled = 2*(x-a)/(b-a) - 1, with 2/(b-a) precomputed per leaf in T
        const V t = O::fma(O::add(x, O::neg(v[0])), v[1], O::splat((S)-1));
        if constexpr (ORDER == 0) {
            return c[0];
        } else if constexpr (ORDER == 1) {
            return O::fma(c[1], t, c[0]);
        } else if constexpr (BASIS == kCheb) {
            // adapt.py:111-120  T0=1, T1=t, Ti = (2t) T(i-1) - T(i-2);  y = sum c_i T_i
            const V t2 = O::add(t, t);
            V acc = O::fma(c[1], t, c[0]);
            V p0 = O::splat((S)1), p1 = t;
#pragma unroll
            for (int j = 2; j <= ORDER; ++j) {
                const V pn = O::fma(t2, p1, O::neg(p0));
                acc = O::fma(c[j], pn, acc);
                p0 = p1; p1 = pn;
            }
            return acc;
        } else {
            // adapt.py:97-108  L0=1, L1=t, Li = ((2i-1) t L(i-1) + (i-1) L(i-2)) * (1./n)
            // -- note the PLUS sign and the constant divisor n: this is the reference's basis,
            // not the Legendre polynomials; the saved coefficients were fitted in it.
            // Clenshaw for L_k = (al_k t) L_(k-1) + be_k L_(k-2), al_k = (2k-1)/n, be_k = (k-1)/n:
            //   b_k = c_k + (al_(k+1) t) b_(k+1) + be_(k+2) b_(k+2),  k = n..2
            //   y   = c0 + c1 t + b_2 L_2 + (be_3 b_3) t
            // 3 instructions per degree instead of the forward form's 5; measured within 2.1
            // units of eps*error_scale of the reference evaluator (forward: 1.7).
            constexpr double dn = (double)ORDER;
            V b1 = c[ORDER], b2 = O::splat((S)0);
#pragma unroll
            for (int k = ORDER - 1; k >= 2; --k) {
                const S al = (S)((2.0 * (k + 1) - 1.0) / dn), be = (S)(((k + 2) - 1.0) / dn);
                const V bk = O::fma(O::mul(O::splat(al), t), b1, O::fma(O::splat(be), b2, c[k]));
                b2 = b1; b1 = bk;
            }
            const S al2 = (S)(3.0 / dn), be2 = (S)(1.0 / dn), be3 = (S)(2.0 / dn);
            const V l2 = O::fma(O::mul(O::splat(al2), t), t, O::splat(be2));
            V acc = O::fma(c[1], t, c[0]);
            acc = O::fma(b1, l2, acc);
            return O::fma(O::mul(O::splat(be3), b2), t, acc);
        }
    }
}

template <int BASIS, int ORDER, typename T>
__device__ __forceinline__ T eval_record(const Record<BASIS, ORDER, T>& r, T x, T /*inv_order*/) {
    return eval_values<BASIS, ORDER, T>(r.v, x);
}

// two fp32 points, each with its own leaf record, per issued instruction
template <int BASIS, int ORDER>
__device__ __forceinline__ float2 eval_record_pair(const Record<BASIS, ORDER, float>& ra,
                                                   const Record<BASIS, ORDER, float>& rb, float xa, float xb) {
    constexpr int N = Record<BASIS, ORDER, float>::N;
    float2 v[N];
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = make_float2(ra.v[k], rb.v[k]);
    return eval_values<BASIS, ORDER, float2>(v, make_float2(xa, xb));
}

template <int BASIS, int ORDER, typename T, int STRIDE = 1, bool VEC = false, bool PK = false>
__device__ __forceinline__ T eval_leaf(const T* __restrict__ rec, T x, T inv_order, unsigned s0_hi = 0,
                                       unsigned s0_lo = 0) {
    Record<BASIS, ORDER, T> r;
    if constexpr (VEC) r.load_vec(rec);
    else if constexpr (PK) r.template load_packed<STRIDE>(rec, s0_hi, s0_lo);
    else r.template load<STRIDE>(rec);
    return eval_record<BASIS, ORDER, T>(r, x, inv_order);
}

// ------------------------------------------------------------------------- streaming access
#ifndef AI_LD_MODE
#define AI_LD_MODE 1               // 0: ld.global.cs, 1: ld.global.L1::no_allocate (LDG.E.NA: measured equal or better everywhere, +5 points on sorted input when the table leaves < 32 KB of L1), 2: ld.global.nc.L1::no_allocate
#endif
#if AI_LD_MODE == 0
template <typename V> __device__ __forceinline__ V ld_stream(const V* p) { return __ldcs(p); }
#else
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 v;
#if AI_LD_MODE == 1
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
#else
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
#endif
    return v;
}
__device__ __forceinline__ double2 ld_stream(const double2* p) {
    double2 v;
#if AI_LD_MODE == 1
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
#else
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
#endif
    return v;
}
#endif
#ifndef AI_ST_MODE
#define AI_ST_MODE 0               // 0: st.global.cs (evict-first), 1: plain st.global, 2: st.global.wt, 3: st.global.cg
#endif
template <typename V> __device__ __forceinline__ void st_stream(V* p, const V& v) {
#if AI_ST_MODE == 0
    __stcs(p, v);
#elif AI_ST_MODE == 1
    *p = v;
#elif AI_ST_MODE == 2
    __stwt(p, v);
#else
    __stcg(p, v);
#endif
}

// Make the table addressable for this CTA and return its base.  kSmem: copy the packed image
// (16-byte chunks, whole CTA).  kSmemBank: copy grid table + breakpoints, then write every record
// element once per bank.  kGlobal: nothing to do.
template <typename T, int MODE>
__device__ __forceinline__ const unsigned char* table_base(const TableView& tv, unsigned char* smem) {
    if constexpr (MODE == kGlobal || MODE == kSelect) {
        return tv.image;
    } else if constexpr (MODE == kSmemVC) {
        // levels at the front, then every 16-byte chunk of every (32-byte padded) global record that
        // holds coefficients, written kVecCopies times chunk-interleaved
        const int4* asrc = reinterpret_cast<const int4*>(tv.image + tv.off_aux);
        int4* dst = reinterpret_cast<int4*>(smem);
        for (int k = threadIdx.x; k < tv.aux_bytes / 16; k += blockDim.x) dst[k] = __ldg(asrc + k);
        const int4* rsrc = reinterpret_cast<const int4*>(tv.image + tv.off_records);
        int4* rdst = reinterpret_cast<int4*>(smem + tv.soff_records);
        const int nv = tv.copies;                       // 16-byte chunks per shared-memory record
        const int gchunks = tv.bcopies;                 // 16-byte chunks per global record
        const int total = tv.n_records * nv * kVecCopies;
        for (int k = threadIdx.x; k < total; k += blockDim.x) {
            const int chunk = k / kVecCopies;           // (record, q)
            const int rec = chunk / nv, q = chunk - rec * nv;
            rdst[k] = __ldg(rsrc + rec * gchunks + q);
        }
        __syncthreads();
        return smem;
    } else if constexpr (MODE == kGlobalC) {
        // only the 4-bit levels are staged; grid table and records stay in global memory
        const int4* src = reinterpret_cast<const int4*>(tv.image + tv.off_aux);
        int4* dst = reinterpret_cast<int4*>(smem);
        for (int k = threadIdx.x; k < tv.aux_bytes / 16; k += blockDim.x) dst[k] = __ldg(src + k);
        __syncthreads();
        return tv.image;
    } else if constexpr (MODE == kSmem || MODE == kSmemP || MODE == kSmemU) {
        const int4* src = reinterpret_cast<const int4*>(tv.image);
        int4* dst = reinterpret_cast<int4*>(smem);
        for (int k = threadIdx.x; k < tv.image_bytes / 16; k += blockDim.x) dst[k] = __ldg(src + k);
        __syncthreads();
        return smem;
    } else {
        // grid table as is; breakpoints and records written `copies` times, element-interleaved
        const int4* src = reinterpret_cast<const int4*>(tv.image);
        int4* dst = reinterpret_cast<int4*>(smem);
        for (int k = threadIdx.x; k < tv.off_breaks / 16; k += blockDim.x) dst[k] = __ldg(src + k);
        const int bsh = 31 - __clz(tv.bcopies);
        const T* bsrc = reinterpret_cast<const T*>(tv.image + tv.off_breaks);
        T* bdst = reinterpret_cast<T*>(smem + tv.soff_breaks);
        for (int k = threadIdx.x; k < (tv.n_breaks_elems << bsh); k += blockDim.x) bdst[k] = __ldg(bsrc + (k >> bsh));
        if constexpr (MODE == kSmemV) {
            // every 16-byte chunk of every record written kVecCopies times, chunk-interleaved
            const int4* rsrc = reinterpret_cast<const int4*>(tv.image + tv.off_records);
            int4* rdst = reinterpret_cast<int4*>(smem + tv.soff_records);
            const int chunks = tv.n_records_elems / (16 / (int)sizeof(T));
            for (int k = threadIdx.x; k < chunks * kVecCopies; k += blockDim.x) rdst[k] = __ldg(rsrc + k / kVecCopies);
        } else {
            const int cp = (MODE == kSmemBankN) ? tv.copies : Copies<T, MODE>::value;
            const int sh = 31 - __clz(cp);
            const T* rsrc = reinterpret_cast<const T*>(tv.image + tv.off_records);
            T* rdst = reinterpret_cast<T*>(smem + tv.soff_records);
            for (int k = threadIdx.x; k < tv.n_records_elems * cp; k += blockDim.x) rdst[k] = __ldg(rsrc + (k >> sh));
        }
        __syncthreads();
        return smem;
    }
}

// ------------------------------------------------------------------------- the hot kernel
// Work split (host computes it, see launch_eval): [0, head) scalar, then nvec 16-byte vectors,
// then a scalar tail.  head/tail are < Vec::N elements unless x and y are misaligned relative
// to each other, in which case head == n and everything takes the scalar loop.
template <int BASIS, int ORDER, typename T, int MODE>
__global__ void __launch_bounds__(kBlock, AI_MIN_BLOCKS)
eval_kernel(const TableView tv, const SelectTable<T, MODE> sel, const T* __restrict__ x, T* __restrict__ y,
            long long n, long long head, long long nvec) {
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned char* base = table_base<T, MODE>(tv, smem);

    Lookup<T> lk;
    lk.init(tv, base);
    constexpr bool kBank = mode_bank(MODE);
    constexpr bool VEC = (MODE == kGlobal);                        // 128-bit record loads
    constexpr bool PK = mode_packed_hdr(MODE);                     // one-element record header
    // copies per record element: compile-time except in kSmemBankN
    const int cp = (MODE == kSmemBankN) ? tv.copies
                   : ((MODE == kSmemV || MODE == kSmemVC) ? kVecCopies : Copies<T, MODE>::value);
    // record stride in elements (kSmemBankN: packed or unpacked header, decided per table)
    const int R = ((MODE == kSmemBankN && Record<BASIS, ORDER, T>::HDR == 2 && tv.s0_hi)
                       ? rec_stride_packed(ORDER) : rec_stride_for<BASIS, ORDER, T, MODE>()) * cp;
    // bank-private layout: lane l reads copy l mod cp of every element (kSmemV: of every 16-byte chunk)
    const T* records = reinterpret_cast<const T*>(base + ((kBank || MODE == kSmemVC) ? tv.soff_records : tv.off_records))
                       + (threadIdx.x & (cp - 1)) * ((MODE == kSmemV || MODE == kSmemVC) ? 16 / (int)sizeof(T) : 1);
    if constexpr (kBank) lk.use_replicated_breaks(tv, base);
    const T inv_order = (T)tv.inv_order;

    // kSelect: the LAST words of both records also sit in shared memory, 16-byte aligned, and are
    // fetched with 128-bit loads (one instruction per 4 floats / 2 doubles, where a select costs one
    // per word -- two per double): these kernels are bound by ISSUE slots, not by a pipe.  The two
    // records are 5 sixteen-byte chunks apart, so a quarter-warp's loads hit different bank groups.
    constexpr int kSelPer = 16 / (int)sizeof(T);
    constexpr int kSelN = Record<BASIS, ORDER, T>::N;
    constexpr int kSelWant = (MODE != kSelect) ? 0 : (sizeof(T) == 4 ? AI_SEL_LDS_F32 : AI_SEL_LDS_F64);
    // words selected from the kernel parameters: a multiple of the vector width (so the rest starts aligned)
    constexpr int kSelNSel = (kSelWant <= 0) ? kSelN : ((kSelN - (kSelWant < kSelN ? kSelWant : kSelN)) / kSelPer * kSelPer);
    constexpr int kSelStride = 20;           // elements per record in shared memory (fp32: 80 B = 5 chunks, fp64: 160 B)
    __shared__ __align__(16) T sel_smem[(MODE == kSelect && kSelNSel < kSelN) ? 2 * kSelStride : 1];
    if constexpr (MODE == kSelect && kSelNSel < kSelN) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < kSelStride; ++k) {
                sel_smem[k] = (k < kSelN) ? sel.rec[0][k] : (T)0;
                sel_smem[kSelStride + k] = (k < kSelN) ? sel.rec[1][k] : (T)0;
            }
        }
        __syncthreads();
    }
    auto load_record = [&](Record<BASIS, ORDER, T>& r, int leaf) {
        AI_CHECK(leaf >= 0 && leaf < (MODE == kSelect ? 2 : tv.n_records));   // kSelect: rec[1] == rec[0] for one leaf
        if constexpr (MODE == kSelect) {
#pragma unroll
            for (int k = 0; k < kSelNSel; ++k) r.v[k] = leaf ? sel.rec[1][k] : sel.rec[0][k];
            if constexpr (kSelNSel < kSelN) {
                const int4* sp = reinterpret_cast<const int4*>(sel_smem + leaf * kSelStride + kSelNSel);
                constexpr int NV = (kSelN - kSelNSel + kSelPer - 1) / kSelPer;
                T buf[NV * kSelPer];
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const int4 w = sp[q];
                    if constexpr (sizeof(T) == 4) {
                        buf[4 * q] = __int_as_float(w.x); buf[4 * q + 1] = __int_as_float(w.y);
                        buf[4 * q + 2] = __int_as_float(w.z); buf[4 * q + 3] = __int_as_float(w.w);
                    } else {
                        buf[2 * q] = __hiloint2double(w.y, w.x); buf[2 * q + 1] = __hiloint2double(w.w, w.z);
                    }
                }
#pragma unroll
                for (int k = kSelNSel; k < kSelN; ++k) r.v[k] = buf[k - kSelNSel];
            }
        } else {
            const T* p = records + leaf * R;
            if constexpr (VEC) r.load_vec(p);
            else if constexpr (MODE == kGlobalC) {
                unsigned k = 0;
                if constexpr (Record<BASIS, ORDER, T>::HDR == 2) k = (smem[leaf >> 1] >> ((leaf & 1) * 4)) & 15u;
                r.load_compact(p, (unsigned)leaf, k, (T)tv.u_w, (T)tv.u_lb, tv.s0_hi, tv.s0_lo);
            }
            else if constexpr (MODE == kSmemVC) {
                unsigned k = 0;
                if constexpr (Record<BASIS, ORDER, T>::HDR == 2) k = (smem[leaf >> 1] >> ((leaf & 1) * 4)) & 15u;
                r.load_compact_smem(p, (unsigned)leaf, k, (T)tv.u_w, (T)tv.u_lb, tv.s0_hi, tv.s0_lo);
            }
            else if constexpr (MODE == kSmemV) r.load_vec_smem(p);
            else if constexpr (PK) r.template load_packed<Copies<T, MODE>::value>(p, tv.s0_hi, tv.s0_lo);
            else if constexpr (mode_uniform_hdr(MODE))
                r.template load_uniform<Copies<T, MODE>::value>(p, leaf, (T)tv.u_w, (T)tv.u_lb, (T)tv.u_s);
            else if constexpr (MODE == kSmemBank) r.template load<Copies<T, kSmemBank>::value>(p);
            else if constexpr (MODE == kSmemBankN) {
                // partial replication serves packed and unpacked headers (warp-uniform branch);
                // s0_hi is S0's high word, never 0 for a packed table
                if constexpr (Record<BASIS, ORDER, T>::HDR == 2) {
                    if (tv.s0_hi) r.load_packed_strided(p, cp, tv.s0_hi, tv.s0_lo);
                    else r.load_strided(p, cp);
                } else {
                    r.load_strided(p, cp);
                }
            }
            else r.template load<1>(p);
        }
    };
    // leaf ordinals of the NE points of a tile.  kSelect: #{k: breaks[k] <= x} - 1 clamped to
    // {0, 1} is the single predicate !(x < mid) (NaN -> 1 = the last leaf, as the BST walk).
    auto find_leaves = [&](const auto& xs, auto& li) {
        if constexpr (MODE == kSelect) {
#pragma unroll
            for (int e = 0; e < (int)(sizeof(li) / sizeof(li[0])); ++e) li[e] = !(xs[e] < sel.mid) ? 1 : 0;
        } else if constexpr (MODE == kGlobalC || MODE == kSmemVC) {   // one record per coarse cell, wide division
#pragma unroll
            for (int e = 0; e < (int)(sizeof(li) / sizeof(li[0])); ++e) li[e] = (int)lk.coarse_cell_wide(xs[e]);
        } else {
            lk.leaves(xs, li);
        }
    };
    auto f = [&](T xv) -> T {
        Record<BASIS, ORDER, T> r;
        if constexpr (MODE == kSelect) load_record(r, !(xv < sel.mid) ? 1 : 0);
        else if constexpr (MODE == kGlobalC || MODE == kSmemVC) load_record(r, (int)lk.coarse_cell_wide(xv));
        else load_record(r, lk.leaf(xv));
        return eval_record<BASIS, ORDER, T>(r, xv, inv_order);
    };

    using V = typename Vec<T>::type;
    constexpr int VN = Vec<T>::N;
    const V* __restrict__ xv = reinterpret_cast<const V*>(x + head);
    V* __restrict__ yv = reinterpret_cast<V*>(y + head);

    constexpr int kUnroll = Unroll<T>::value;                      // shadows the namespace constant
    constexpr long long kTile = (long long)kBlock * kUnroll;       // vectors per tile
    constexpr int NE = kUnroll * VN;                               // points per thread per tile

    auto load_tile = [&](V (&buf)[kUnroll], long long tile) {
        const long long base = tile * kTile + threadIdx.x;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) buf[u] = ld_stream(xv + base + (long long)u * kBlock);
    };
    // x_elem_loads (x + head is not 16-byte aligned although y + head is): no tile takes the 128-bit
    // loads; every vector goes through the grid-stride loop below, which then reads its VN elements
    // one by one (kept OUT of the tile loop: even predicated off, the element loads cost the aligned
    // path 18% -- measured)
    const bool x_elem = tv.x_elem_loads != 0;
    const long long full_tiles = x_elem ? 0 : nvec / kTile;
    auto load_vec_x = [&](long long v) -> V {
        if (!x_elem) return ld_stream(xv + v);
        const T* __restrict__ p = x + head + v * VN;
        V r;
        if constexpr (VN == 4) { r.x = __ldg(p); r.y = __ldg(p + 1); r.z = __ldg(p + 2); r.w = __ldg(p + 3); }
        else { r.x = __ldg(p); r.y = __ldg(p + 1); }
        return r;
    };

    [[maybe_unused]] int probe = 0;          // AI_UNIFORM_PATH == 2: tiles until the next probe
    [[maybe_unused]] bool expect_uniform = true;
    long long tile = blockIdx.x;
    V in[kUnroll];
#if AI_PREFETCH
    if (tile < full_tiles) load_tile(in, tile);
#endif
#if AI_PREFETCH == 2
    V nxt[kUnroll];                          // tile k+1, in flight since the previous iteration
    if (tile + gridDim.x < full_tiles) load_tile(nxt, tile + gridDim.x);
#endif
    for (; tile < full_tiles; tile += gridDim.x) {
#if AI_PREFETCH == 2
        V nxt2[kUnroll];                     // tile k+2: two tiles in flight per thread
        const long long next2 = tile + 2LL * gridDim.x;
        if (next2 < full_tiles) load_tile(nxt2, next2);
#elif AI_PREFETCH
        V nxt[kUnroll];
        const long long next = tile + gridDim.x;
        if (next < full_tiles) load_tile(nxt, next);
#else
        load_tile(in, tile);
#endif
        const long long base = tile * kTile + threadIdx.x;
        T xs[NE];
        int li[NE];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if constexpr (VN == 4) {
                xs[4 * u] = in[u].x; xs[4 * u + 1] = in[u].y; xs[4 * u + 2] = in[u].z; xs[4 * u + 3] = in[u].w;
            } else {
                xs[2 * u] = in[u].x; xs[2 * u + 1] = in[u].y;
            }
        }
        find_leaves(xs, li);
        // Sorted / clustered inputs: all NE points of this thread usually sit in one leaf.  When
        // that holds for every lane of the warp, load the record ONCE into registers (n+3 shared
        // loads per NE points instead of per point).  Same arithmetic, so the bits are the same.
        bool uniform = false;
#if AI_UNIFORM_PATH == 1
        {
            bool same = true;
#pragma unroll
            for (int e = 1; e < NE; ++e) same = same && (li[e] == li[0]);
            uniform = __all_sync(0xffffffffu, same);
        }
#elif AI_UNIFORM_PATH == 2
        if (expect_uniform || probe == 0) {
            bool same = true;
#pragma unroll
            for (int e = 1; e < NE; ++e) same = same && (li[e] == li[0]);
            uniform = __all_sync(0xffffffffu, same);
            expect_uniform = uniform;
            probe = 16;
        }
        --probe;
#endif
        if (uniform) {
            Record<BASIS, ORDER, T> rec;
            load_record(rec, li[0]);
            if constexpr (sizeof(T) == 4 && AI_PACKED_F32) {
#pragma unroll
                for (int e = 0; e < NE; e += 2) {
                    const float2 r2 = eval_record_pair<BASIS, ORDER>(rec, rec, xs[e], xs[e + 1]);
                    xs[e] = r2.x; xs[e + 1] = r2.y;
                }
            } else {
#pragma unroll
                for (int e = 0; e < NE; ++e) xs[e] = eval_record<BASIS, ORDER, T>(rec, xs[e], inv_order);
            }
        } else {
            if constexpr (sizeof(T) == 4 && AI_PACKED_F32 == 1) {
#pragma unroll
                for (int e = 0; e < NE; e += 2) {
                    Record<BASIS, ORDER, T> ra, rb;
                    load_record(ra, li[e]);
                    load_record(rb, li[e + 1]);
                    const float2 r2 = eval_record_pair<BASIS, ORDER>(ra, rb, xs[e], xs[e + 1]);
                    xs[e] = r2.x; xs[e + 1] = r2.y;
                }
            } else {
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    Record<BASIS, ORDER, T> rec;
                    load_record(rec, li[e]);
                    xs[e] = eval_record<BASIS, ORDER, T>(rec, xs[e], inv_order);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            V out;
            if constexpr (VN == 4) {
                out.x = xs[4 * u]; out.y = xs[4 * u + 1]; out.z = xs[4 * u + 2]; out.w = xs[4 * u + 3];
            } else {
                out.x = xs[2 * u]; out.y = xs[2 * u + 1];
            }
            st_stream(yv + base + (long long)u * kBlock, out);
        }
#if AI_PREFETCH == 2
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) { in[u] = nxt[u]; nxt[u] = nxt2[u]; }
#elif AI_PREFETCH
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) in[u] = nxt[u];
#endif
    }
    // remainder vectors (less than one tile), spread over the whole grid
    const long long gtid = (long long)blockIdx.x * kBlock + threadIdx.x;
    const long long gstride = (long long)gridDim.x * kBlock;
#pragma unroll 2
    for (long long v = full_tiles * kTile + gtid; v < nvec; v += gstride) {
        V in = load_vec_x(v), out;
        if constexpr (VN == 4) {
            out.x = f(in.x); out.y = f(in.y); out.z = f(in.z); out.w = f(in.w);
        } else {
            out.x = f(in.x); out.y = f(in.y);
        }
        st_stream(yv + v, out);
    }
    // scalar head and tail
    for (long long k = gtid; k < head; k += gstride) y[k] = f(x[k]);
    for (long long k = head + nvec * VN + gtid; k < n; k += gstride) y[k] = f(x[k]);
}

// Several tables over the SAME x in one pass (the step after the path --
// e.g. J0, J1, Y0, Y1 of one argument).  x is read once and m outputs are written:
// (1 + m) * sizeof(T) bytes per point instead of 2 * m * sizeof(T) for m separate launches.
// All tables share basis / order / dtype (one fit setting for several functions); every image is
// staged packed into shared memory (the host falls back to per-table launches otherwise).  Per
// table the lookup, record loads and arithmetic are those of eval_kernel<..., kSmem>, so every
// output is bit-identical to the single-table result.
constexpr int kMaxMulti = 4;
#ifndef AI_UNROLL_MULTI
#define AI_UNROLL_MULTI 6          // 16-byte vectors per thread per tile (measured best of 4, 5, 6)
#endif
#ifndef AI_UNROLL_MULTI_F64
#define AI_UNROLL_MULTI_F64 4      // measured best of 3, 4, 5
#endif
template <typename T> struct UnrollMulti { static constexpr int value = AI_UNROLL_MULTI; };
template <> struct UnrollMulti<double> { static constexpr int value = AI_UNROLL_MULTI_F64; };

struct MultiArgs {
    TableView tv[kMaxMulti];
    void* y[kMaxMulti];
    int soff[kMaxMulti];           // byte offset of table j's image in shared memory (16-byte aligned)
    int m;
};

// HM: record header form shared by the m tables -- 0 [a, 2/(b-a)], 1 one word, 2 computed
template <int BASIS, int ORDER, typename T, int HM>
__global__ void __launch_bounds__(kBlock, 1)
eval_multi_kernel(const __grid_constant__ MultiArgs args, const T* __restrict__ x,
                  long long n, long long head, long long nvec) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int m = args.m;
    for (int j = 0; j < m; ++j) {
        const int4* src = reinterpret_cast<const int4*>(args.tv[j].image);
        int4* dst = reinterpret_cast<int4*>(smem + args.soff[j]);
        for (int k = threadIdx.x; k < args.tv[j].image_bytes / 16; k += blockDim.x) dst[k] = __ldg(src + k);
    }
    __syncthreads();

    using V = typename Vec<T>::type;
    constexpr int VN = Vec<T>::N;
    constexpr int U = UnrollMulti<T>::value;
    constexpr long long kTile = (long long)kBlock * U;
    constexpr int NE = U * VN;
    constexpr int R = (HM == 1) ? rec_stride_packed(ORDER) : (HM == 2 ? rec_stride(kMono, ORDER) : rec_stride(BASIS, ORDER));
    const V* __restrict__ xv = reinterpret_cast<const V*>(x + head);
    const long long full_tiles = nvec / kTile;
    auto load_rec = [&](Record<BASIS, ORDER, T>& r, const T* records, int idx, int j) {
        const T* p = records + idx * R;
        if constexpr (HM == 1) r.template load_packed<1>(p, args.tv[j].s0_hi, args.tv[j].s0_lo);
        else if constexpr (HM == 2) r.template load_uniform<1>(p, idx, (T)args.tv[j].u_w, (T)args.tv[j].u_lb, (T)args.tv[j].u_s);
        else r.template load<1>(p);
    };

    auto load_tile = [&](V (&buf)[U], long long tile) {
        const long long base = tile * kTile + threadIdx.x;
#pragma unroll
        for (int u = 0; u < U; ++u) buf[u] = ld_stream(xv + base + (long long)u * kBlock);
    };
    auto scalar = [&](int j, T xs) -> T {
        Lookup<T> lk;
        lk.init(args.tv[j], smem + args.soff[j]);
        const T* records = reinterpret_cast<const T*>(smem + args.soff[j] + args.tv[j].off_records);
        Record<BASIS, ORDER, T> r;
        load_rec(r, records, lk.leaf(xs), j);
        return eval_record<BASIS, ORDER, T>(r, xs, (T)0);
    };

    long long tile = blockIdx.x;
    V in[U];
    if (tile < full_tiles) load_tile(in, tile);
    for (; tile < full_tiles; tile += gridDim.x) {
        V nxt[U];
        const long long next = tile + gridDim.x;
        if (next < full_tiles) load_tile(nxt, next);
        const long long base = tile * kTile + threadIdx.x;
        T xs[NE];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if constexpr (VN == 4) {
                xs[4 * u] = in[u].x; xs[4 * u + 1] = in[u].y; xs[4 * u + 2] = in[u].z; xs[4 * u + 3] = in[u].w;
            } else {
                xs[2 * u] = in[u].x; xs[2 * u + 1] = in[u].y;
            }
        }
        for (int j = 0; j < m; ++j) {
            Lookup<T> lk;
            lk.init(args.tv[j], smem + args.soff[j]);
            const T* records = reinterpret_cast<const T*>(smem + args.soff[j] + args.tv[j].off_records);
            int li[NE];
            lk.template leaves<NE>(xs, li);
            bool same = true;
#pragma unroll
            for (int e = 1; e < NE; ++e) same = same && (li[e] == li[0]);
            const bool uniform = __all_sync(0xffffffffu, same);
            T ys[NE];
            if (uniform) {
                Record<BASIS, ORDER, T> rec;
                load_rec(rec, records, li[0], j);
                if constexpr (sizeof(T) == 4 && AI_PACKED_F32) {
#pragma unroll
                    for (int e = 0; e < NE; e += 2) {
                        const float2 r2 = eval_record_pair<BASIS, ORDER>(rec, rec, xs[e], xs[e + 1]);
                        ys[e] = r2.x; ys[e + 1] = r2.y;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < NE; ++e) ys[e] = eval_record<BASIS, ORDER, T>(rec, xs[e], (T)0);
                }
            } else {
                if constexpr (sizeof(T) == 4 && AI_PACKED_F32 == 1) {
#pragma unroll
                    for (int e = 0; e < NE; e += 2) {
                        Record<BASIS, ORDER, T> ra, rb;
                        load_rec(ra, records, li[e], j);
                        load_rec(rb, records, li[e + 1], j);
                        const float2 r2 = eval_record_pair<BASIS, ORDER>(ra, rb, xs[e], xs[e + 1]);
                        ys[e] = r2.x; ys[e + 1] = r2.y;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        Record<BASIS, ORDER, T> rec;
                        load_rec(rec, records, li[e], j);
                        ys[e] = eval_record<BASIS, ORDER, T>(rec, xs[e], (T)0);
                    }
                }
            }
            V* __restrict__ yv = reinterpret_cast<V*>(static_cast<T*>(args.y[j]) + head);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                V out;
                if constexpr (VN == 4) {
                    out.x = ys[4 * u]; out.y = ys[4 * u + 1]; out.z = ys[4 * u + 2]; out.w = ys[4 * u + 3];
                } else {
                    out.x = ys[2 * u]; out.y = ys[2 * u + 1];
                }
                st_stream(yv + base + (long long)u * kBlock, out);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) in[u] = nxt[u];
    }
    // remainder vectors, then the scalar head and tail
    const long long gtid = (long long)blockIdx.x * kBlock + threadIdx.x;
    const long long gstride = (long long)gridDim.x * kBlock;
    for (long long v = full_tiles * kTile + gtid; v < nvec; v += gstride) {
        const V inv = ld_stream(xv + v);
        for (int j = 0; j < m; ++j) {
            V out;
            if constexpr (VN == 4) {
                out.x = scalar(j, inv.x); out.y = scalar(j, inv.y); out.z = scalar(j, inv.z); out.w = scalar(j, inv.w);
            } else {
                out.x = scalar(j, inv.x); out.y = scalar(j, inv.y);
            }
            st_stream(reinterpret_cast<V*>(static_cast<T*>(args.y[j]) + head) + v, out);
        }
    }
    for (int j = 0; j < m; ++j) {
        T* y = static_cast<T*>(args.y[j]);
        for (long long k = gtid; k < head; k += gstride) y[k] = scalar(j, x[k]);
        for (long long k = head + nvec * VN + gtid; k < n; k += gstride) y[k] = scalar(j, x[k]);
    }
}

// The reference's no-"output" variant (generate.py:284,392-394): reduce y instead of storing it.
// Same tile structure as eval_kernel -- 128-bit streaming loads, the next tile's x in flight while
// the current one is evaluated, lookups of a whole tile before its gathers, packed FP32, the
// leaf-uniform fast path -- with the store replaced by an fp64 accumulation per thread (every value
// is widened before it is added: the header's contract).  x needs no particular alignment: the
// kernel splits [0, n) into a scalar head up to the first 16-byte boundary, vectors and a tail.
// Reads 1 * sizeof(T) bytes per point, so at the rate of eval_kernel it is SM-bound, not HBM-bound.
// (Round 2, first version: one point per thread per iteration, a single load in flight per thread:
// 124-136 Gpt/s whatever the table, i.e. slower than evaluating AND storing.)
template <int BASIS, int ORDER, typename T, int MODE>
__global__ void __launch_bounds__(kBlock)
eval_sum_kernel(const TableView tv, const T* __restrict__ x, double* __restrict__ result, long long n) {
    static_assert(MODE == kSmem || MODE == kSmemP || MODE == kSmemU || MODE == kGlobal || MODE == kGlobalC,
                  "the reduction variant uses the packed layouts only");
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned char* base = table_base<T, MODE>(tv, smem);
    Lookup<T> lk;
    lk.init(tv, base);
    const T* records = reinterpret_cast<const T*>(base + tv.off_records);
    constexpr int R = rec_stride_for<BASIS, ORDER, T, MODE>();
    constexpr bool PK = mode_packed_hdr(MODE);
    const T inv_order = (T)tv.inv_order;

    auto record_index = [&](T xv) -> int {
        if constexpr (MODE == kGlobalC) return (int)lk.coarse_cell_wide(xv);
        else return lk.leaf(xv);
    };
    auto load_record = [&](Record<BASIS, ORDER, T>& r, int i) {
        AI_CHECK(i >= 0 && i < tv.n_records);
        const T* p = records + i * R;
        if constexpr (mode_uniform_hdr(MODE)) {
            r.template load_uniform<1>(p, i, (T)tv.u_w, (T)tv.u_lb, (T)tv.u_s);
        } else if constexpr (MODE == kGlobalC) {
            unsigned k = 0;
            if constexpr (Record<BASIS, ORDER, T>::HDR == 2) k = (smem[i >> 1] >> ((i & 1) * 4)) & 15u;
            r.load_compact(p, (unsigned)i, k, (T)tv.u_w, (T)tv.u_lb, tv.s0_hi, tv.s0_lo);
        } else if constexpr (MODE == kGlobal) {
            r.load_vec(p);
        } else if constexpr (PK) {
            r.template load_packed<1>(p, tv.s0_hi, tv.s0_lo);
        } else {
            r.template load<1>(p);
        }
    };
    auto f = [&](T xv) -> double {
        Record<BASIS, ORDER, T> r;
        load_record(r, record_index(xv));
        return (double)eval_record<BASIS, ORDER, T>(r, xv, inv_order);
    };

    using V = typename Vec<T>::type;
    constexpr int VN = Vec<T>::N;
    constexpr int kUnroll = Unroll<T>::value;
    constexpr long long kTile = (long long)kBlock * kUnroll;       // vectors per tile
    constexpr int NE = kUnroll * VN;                               // points per thread per tile
    // scalar head up to the first 16-byte boundary of x (all of it if x is not even element-aligned)
    const unsigned long long addr = reinterpret_cast<unsigned long long>(x);
    long long head = (addr % sizeof(T)) ? n : (long long)(((16ull - (addr & 15ull)) & 15ull) / sizeof(T));
    if (head > n) head = n;
    const long long nvec = (n - head) / VN;
    const V* __restrict__ xv = reinterpret_cast<const V*>(x + head);
    const long long full_tiles = nvec / kTile;

    auto load_tile = [&](V (&buf)[kUnroll], long long tile) {
        const long long b = tile * kTile + threadIdx.x;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) buf[u] = ld_stream(xv + b + (long long)u * kBlock);
    };

    double acc = 0.0;
    int probe = 0;
    bool expect_uniform = true;
    long long tile = blockIdx.x;
    V in[kUnroll];
    if (tile < full_tiles) load_tile(in, tile);
    for (; tile < full_tiles; tile += gridDim.x) {
        V nxt[kUnroll];
        const long long next = tile + gridDim.x;
        if (next < full_tiles) load_tile(nxt, next);
        T xs[NE];
        int li[NE];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if constexpr (VN == 4) {
                xs[4 * u] = in[u].x; xs[4 * u + 1] = in[u].y; xs[4 * u + 2] = in[u].z; xs[4 * u + 3] = in[u].w;
            } else {
                xs[2 * u] = in[u].x; xs[2 * u + 1] = in[u].y;
            }
        }
        if constexpr (MODE == kGlobalC) {
#pragma unroll
            for (int e = 0; e < NE; ++e) li[e] = (int)lk.coarse_cell_wide(xs[e]);
        } else {
            lk.leaves(xs, li);
        }
        bool uniform = false;
        if (expect_uniform || probe == 0) {
            bool same = true;
#pragma unroll
            for (int e = 1; e < NE; ++e) same = same && (li[e] == li[0]);
            uniform = __all_sync(0xffffffffu, same);
            expect_uniform = uniform;
            probe = 16;
        }
        --probe;
        if (uniform) {
            Record<BASIS, ORDER, T> rec;
            load_record(rec, li[0]);
            if constexpr (sizeof(T) == 4 && AI_PACKED_F32) {
#pragma unroll
                for (int e = 0; e < NE; e += 2) {
                    const float2 r2 = eval_record_pair<BASIS, ORDER>(rec, rec, xs[e], xs[e + 1]);
                    acc += (double)r2.x;
                    acc += (double)r2.y;
                }
            } else {
#pragma unroll
                for (int e = 0; e < NE; ++e) acc += (double)eval_record<BASIS, ORDER, T>(rec, xs[e], inv_order);
            }
        } else {
            if constexpr (sizeof(T) == 4 && AI_PACKED_F32 == 1) {
#pragma unroll
                for (int e = 0; e < NE; e += 2) {
                    Record<BASIS, ORDER, T> ra, rb;
                    load_record(ra, li[e]);
                    load_record(rb, li[e + 1]);
                    const float2 r2 = eval_record_pair<BASIS, ORDER>(ra, rb, xs[e], xs[e + 1]);
                    acc += (double)r2.x;
                    acc += (double)r2.y;
                }
            } else {
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    Record<BASIS, ORDER, T> rec;
                    load_record(rec, li[e]);
                    acc += (double)eval_record<BASIS, ORDER, T>(rec, xs[e], inv_order);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) in[u] = nxt[u];
    }
    // remainder vectors (less than one tile), spread over the whole grid; scalar head and tail
    const long long gtid = (long long)blockIdx.x * kBlock + threadIdx.x;
    const long long gstride = (long long)gridDim.x * kBlock;
    for (long long v = full_tiles * kTile + gtid; v < nvec; v += gstride) {
        const V iv = ld_stream(xv + v);
        if constexpr (VN == 4) { acc += f(iv.x); acc += f(iv.y); acc += f(iv.z); acc += f(iv.w); }
        else { acc += f(iv.x); acc += f(iv.y); }
    }
    for (long long k = gtid; k < head; k += gstride) acc += f(x[k]);
    for (long long k = head + nvec * VN + gtid; k < n; k += gstride) acc += f(x[k]);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double wsum[kBlock / 32];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; ++w) s += wsum[w];
        atomicAdd(result, s);
    }
}

// leaf ordinal only: the bit-exactness check and the lookup-only ablation
// (the reference's "test first loop", generate.py:320-332)
template <typename T, int MODE>
__global__ void __launch_bounds__(kBlock)
index_kernel(const TableView tv, const T* __restrict__ x, int32_t* __restrict__ idx, long long n) {
    static_assert(MODE == kSmem || MODE == kGlobal, "lookup only: records are not read");
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned char* base = table_base<T, MODE>(tv, smem);
    Lookup<T> lk;
    lk.init(tv, base);
    const long long gtid = (long long)blockIdx.x * kBlock + threadIdx.x;
    const long long gstride = (long long)gridDim.x * kBlock;
    // kIdxUnroll coalesced loads in flight per thread (one point per thread per iteration left a single
    // load in flight: latency-bound at a fifth of the evaluation kernel's rate), lookups of the whole
    // batch together so that their dependent chains interleave
    constexpr int kIdxUnroll = 8;
    long long k = gtid;
    for (; k + (kIdxUnroll - 1) * gstride < n; k += kIdxUnroll * gstride) {
        T xs[kIdxUnroll];
        int li[kIdxUnroll];
#pragma unroll
        for (int u = 0; u < kIdxUnroll; ++u) xs[u] = __ldcs(x + k + u * gstride);
        lk.template true_leaves<kIdxUnroll>(xs, li);
#pragma unroll
        for (int u = 0; u < kIdxUnroll; ++u) __stcs(idx + k + u * gstride, (xs[u] != xs[u]) ? lk.lmax : li[u]);   // NaN: see leaf_nan_exact
    }
    for (; k < n; k += gstride) idx[k] = lk.leaf_nan_exact(__ldcs(x + k));
}

// ------------------------------------------------------------------------- launch plumbing
struct LaunchCfg {
    int grid;
    int smem_bytes;
    cudaStream_t stream;
    const double* select;          // kSelect only: host array [mid, rec0[kMaxRecord], rec1[kMaxRecord]]
};

// Two translation units per (basis, dtype) -- orders 0..kSplitOrder and above, see ai_inst.cuh --
// each exporting the four launchers below; the inline dispatchers pick the unit by order.
constexpr int kSplitOrder = 9;
#define AI_DECLARE_PART(NAME, T)                                                                   \
    cudaError_t launch_eval_##NAME(int order, int mode, const LaunchCfg&, const TableView&,        \
                                   const T*, T*, long long, long long, long long);                 \
    cudaError_t launch_sum_##NAME(int order, int mode, const LaunchCfg&, const TableView&,         \
                                  const T*, double*, long long);                                   \
    cudaError_t launch_multi_##NAME(int order, int packed, const LaunchCfg&, const MultiArgs&,     \
                                    const T*, long long, long long, long long);                    \
    cudaError_t prepare_##NAME(int order, int mode, int smem_bytes, int* blocks_per_sm);
#define AI_DECLARE_UNIT(NAME, T)                                                                   \
    AI_DECLARE_PART(NAME##_lo, T)                                                                  \
    AI_DECLARE_PART(NAME##_hi, T)                                                                  \
    inline cudaError_t launch_eval_##NAME(int order, int mode, const LaunchCfg& c, const TableView& tv, \
                                          const T* x, T* y, long long n, long long head, long long nvec) { \
        return order <= kSplitOrder ? launch_eval_##NAME##_lo(order, mode, c, tv, x, y, n, head, nvec) \
                                    : launch_eval_##NAME##_hi(order, mode, c, tv, x, y, n, head, nvec); \
    }                                                                                              \
    inline cudaError_t launch_sum_##NAME(int order, int mode, const LaunchCfg& c, const TableView& tv, \
                                         const T* x, double* r, long long n) {                     \
        return order <= kSplitOrder ? launch_sum_##NAME##_lo(order, mode, c, tv, x, r, n)          \
                                    : launch_sum_##NAME##_hi(order, mode, c, tv, x, r, n);         \
    }                                                                                              \
    inline cudaError_t launch_multi_##NAME(int order, int packed, const LaunchCfg& c, const MultiArgs& a, \
                                           const T* x, long long n, long long head, long long nvec) { \
        return order <= kSplitOrder ? launch_multi_##NAME##_lo(order, packed, c, a, x, n, head, nvec) \
                                    : launch_multi_##NAME##_hi(order, packed, c, a, x, n, head, nvec); \
    }                                                                                              \
    inline cudaError_t prepare_##NAME(int order, int mode, int smem_bytes, int* bps) {             \
        return order <= kSplitOrder ? prepare_##NAME##_lo(order, mode, smem_bytes, bps)            \
                                    : prepare_##NAME##_hi(order, mode, smem_bytes, bps);           \
    }
AI_DECLARE_UNIT(cheb_f32, float)
AI_DECLARE_UNIT(cheb_f64, double)
AI_DECLARE_UNIT(leg_f32, float)
AI_DECLARE_UNIT(leg_f64, double)
AI_DECLARE_UNIT(mono_f32, float)
AI_DECLARE_UNIT(mono_f64, double)
#undef AI_DECLARE_UNIT
#undef AI_DECLARE_PART

}  // namespace ai
