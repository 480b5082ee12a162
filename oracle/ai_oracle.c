/*
 * ai_oracle.c -- CPU restatement of the reference evaluator.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product (adaptive_interpolation_b200/) never does; it has no CPU path.
 *
 * What is restated (reference file:line, all under /root/reference/adaptive_interpolation/):
 *   approximator.py:273-296  Approximator.evaluate: BST walk over tree_1d, then basis + dot
 *   adapt.py:129-139         Interpolant.basis_function: rescale and basis dispatch
 *   adapt.py:111-120         chebyshev(n, x)
 *   adapt.py:97-108          legendre(n, x)   (the reference's Legendre-LIKE recurrence)
 *   adapt.py:139             monomial: [x**i]
 * Arithmetic is done in the table dtype, one rounding per operation, in the reference's operation
 * order; compile with -ffp-contract=off so the compiler fuses nothing.  The final np.dot of the
 * reference goes through BLAS (summation order unspecified); here it is a left-to-right sum, so
 * values agree with the reference to a few ulp of sum|c_i B_i|, not bit for bit.  Leaf selection
 * IS bit-exact.
 *
 * Parity pinned: tests/test_oracle.py checks these functions against the frozen outputs of the
 * real reference evaluator for all 44 tables under tests/golden/tables/.
 */
#include <math.h>
#include <stdint.h>

#define AI_CHEB 0
#define AI_LEG 1
#define AI_MONO 2
#define AI_MAXN 64

#define DEFINE_ORACLE(T, SUFFIX, POW)                                                            \
    /* approximator.py:285-290: the walk; returns the element offset of the selected leaf */     \
    static int64_t walk_##SUFFIX(const T* tree, int order, int num_levels, T xn) {               \
        int64_t index = 0;                                                                       \
        const int child = order + 4;                                                             \
        for (int i = 1; i < num_levels; ++i) {                                                   \
            if (tree[index] > xn)                                                                \
                index = (int64_t)tree[index + child];                                            \
            else                                                                                 \
                index = (int64_t)tree[index + child + 1];                                        \
        }                                                                                        \
        return index;                                                                            \
    }                                                                                            \
    /* adapt.py:129-139 + :97-120: basis values B[0..n] in T arithmetic */                       \
    static void basis_##SUFFIX(int basis, int n, T x, T a, T b, T* B) {                          \
        if (basis == AI_MONO) {                                                                  \
            for (int i = 0; i <= n; ++i) B[i] = (T)POW(x, (T)i);          /* adapt.py:139 */     \
            return;                                                                              \
        }                                                                                        \
        T t = (T)2 * (x - a);                                             /* adapt.py:130 */     \
        t = t / (b - a);                                                                         \
        t = t - (T)1;                                                                            \
        B[0] = (T)1;                                                                             \
        if (n == 0) return;                                                                      \
        B[1] = t;                                                                                \
        if (basis == AI_CHEB) {                                                                  \
            for (int i = 2; i <= n; ++i) {                                /* adapt.py:118-119 */ \
                T p = (T)2 * t;                                                                  \
                p = p * B[i - 1];                                                                \
                B[i] = p - B[i - 2];                                                             \
            }                                                                                    \
        } else {                                                                                 \
            const T invn = (T)(1.0 / (double)n);                          /* adapt.py:107 */     \
            for (int i = 2; i <= n; ++i) {                                /* adapt.py:104-107 */ \
                T first = (T)(2 * i - 1) * t;                                                    \
                first = first * B[i - 1];                                                        \
                T second = (T)(i - 1) * B[i - 2];                                                \
                T s = first + second;                                                            \
                B[i] = s * invn;                                                                 \
            }                                                                                    \
        }                                                                                        \
    }                                                                                            \
    /* approximator.py:273-296 for n points.  leaf_off / cond may be NULL.                    */ \
    /* cond[k] = sum_i |c_i * B_i(x_k)| in double: the scale rounding differences live on.    */ \
    void ai_oracle_eval_##SUFFIX(const T* tree, int order, int num_levels, int basis,           \
                                 const T* x, T* y, int64_t* leaf_off, double* cond, int64_t n,  \
                                 int threads) {                                                  \
        (void)threads;                                                                           \
        _Pragma("omp parallel for num_threads(threads) schedule(static)")                        \
        for (int64_t k = 0; k < n; ++k) {                                                        \
            T B[AI_MAXN];                                                                        \
            const T xn = x[k];                                                                   \
            const int64_t index = walk_##SUFFIX(tree, order, num_levels, xn);                    \
            const T a = tree[index + order + 2], b = tree[index + order + 3];                    \
            basis_##SUFFIX(basis, order, xn, a, b, B);                                           \
            T s = (T)0;                                                                          \
            double c = 0.0;                                                                      \
            for (int i = 0; i <= order; ++i) {                            /* np.dot */           \
                const T term = tree[index + 1 + i] * B[i];                                       \
                s = s + term;                                                                    \
                c += fabs((double)tree[index + 1 + i] * (double)B[i]);                           \
            }                                                                                    \
            y[k] = s;                                                                            \
            if (leaf_off) leaf_off[k] = index;                                                   \
            if (cond) cond[k] = c;                                                               \
        }                                                                                        \
    }

DEFINE_ORACLE(float, f32, powf)
DEFINE_ORACLE(double, f64, pow)

int ai_oracle_abi_version(void) { return 1; }
