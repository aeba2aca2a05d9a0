/* C restatement of the reference's CPU algorithm STRUCTURE for the batched
 * trafo-chain path -- the timed CPU baseline of bench.py.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The reference is pure Julia and Julia is not installed in
 * this image, so this is a "port", not the reference itself.
 *
 * It is deliberately UNFUSED, because the reference is: every trafo makes one
 * broadcast pass that allocates its output, a second broadcast pass that
 * allocates a D x N ladj temporary, and a column reduction; every Householder
 * reflection is a dot-product pass plus an update pass.  Formulas are the
 * reference's literal ones (file:line cited at each function).  Single-threaded
 * is the faithful setting (the reference has no threading apart from BLAS inside
 * `v' * x`, src/householder_trafo.jl:4); OMP_NUM_THREADS > 1 parallelises every
 * pass over columns and is reported as the "generous" figure.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define NAME(x) CAT(x, SUF)

/* ---- scalar kernels, generic over REAL ------------------------------------------- */
#define DEFINE_SCALARS                                                                                   \
    /* src/center_stretch.jl:4-8 */                                                                       \
    static inline REAL NAME(center_stretch)(REAL x, REAL a, REAL b, REAL c) {                             \
        const REAL E = EXP(FABS(b * x));                                                                  \
        const REAL om = (REAL)1 - E;                                                                      \
        const REAL sg = (x > 0) - (x < 0);                                                                \
        return sg * LOG((SQRT(om * om * EXP((REAL)2 * b * a) + (REAL)4 * E) - om * EXP(b * a)) / (REAL)2) / b + c; \
    }                                                                                                     \
    /* src/center_stretch.jl:11-15 */                                                                     \
    static inline REAL NAME(center_contract)(REAL x, REAL a, REAL b, REAL c) {                            \
        const REAL u = x - c;                                                                             \
        return (LOG((REAL)1 + EXP(b * (u - a))) - LOG((REAL)1 + EXP(-b * (u + a)))) / b;                  \
    }                                                                                                     \
    /* src/center_stretch.jl:17-22 */                                                                     \
    static inline REAL NAME(center_contract_ladj)(REAL x, REAL a, REAL b, REAL c) {                       \
        const REAL u = x - c;                                                                             \
        const REAL d = (REAL)1 / ((REAL)1 + EXP(-b * (u - a))) + (REAL)1 / ((REAL)1 + EXP(b * (u + a)));  \
        return LOG(FABS(d));                                                                              \
    }                                                                                                     \
    /* src/johnson_trafo.jl:29-32 */                                                                      \
    static inline REAL NAME(johnsontrafo)(REAL x, REAL g, REAL d, REAL xi, REAL l) {                      \
        return g + d * ASINH((x - xi) / l);                                                               \
    }                                                                                                     \
    /* src/johnson_trafo.jl:34-37 */                                                                      \
    static inline REAL NAME(johnsontrafo_inv)(REAL x, REAL g, REAL d, REAL xi, REAL l) {                  \
        return l * SINH((x - g) / d) + xi;                                                                \
    }                                                                                                     \
    /* src/johnson_trafo.jl:39-42,49-52 */                                                                \
    static inline REAL NAME(johnsontrafo_ladj)(REAL x, REAL g, REAL d, REAL xi, REAL l) {                 \
        const REAL z = (x - xi) / l;                                                                      \
        return LOG(FABS((d / l) * ((REAL)1 / SQRT((REAL)1 + z * z))));                                    \
    }

#define REAL float
#define SUF f32
#define EXP expf
#define LOG logf
#define SQRT sqrtf
#define FABS fabsf
#define ASINH asinhf
#define SINH sinhf
#define FMA fmaf
DEFINE_SCALARS
#include "enf_ref_cpu_impl.h"
#undef REAL
#undef SUF
#undef EXP
#undef LOG
#undef SQRT
#undef FABS
#undef ASINH
#undef SINH
#undef FMA

#define REAL double
#define SUF f64
#define EXP exp
#define LOG log
#define SQRT sqrt
#define FABS fabs
#define ASINH asinh
#define SINH sinh
#define FMA fma
DEFINE_SCALARS
#include "enf_ref_cpu_impl.h"

int ref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void ref_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void)n;
#endif
}
