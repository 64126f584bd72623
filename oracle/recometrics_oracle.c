/* TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.
 *
 * CPU restatement (plain C) of the per-user evaluation path of david-cortes/recometrics,
 * `calc_metrics<real_t>` (/root/reference/src/recometrics.hpp:359-965), written from the
 * behavioural spec in SURVEY.md Appendix A.  It is the checker for the CUDA path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (recometrics_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks this restatement bit-for-bit
 * against oracle/_ref/librecometrics_ref.so (the unmodified reference compiled by
 * oracle/Makefile from /root/reference/src where it lies) and against the committed
 * fixtures in tests/golden/ that were generated from that same library
 * (tests/golden/make_golden.py).  The known-answer values of the reference's own R tests
 * (tests/testthat/test-auc.R, test-ndcg.R) are restated in tests/test_reference_kats.py.
 *
 * Deliberate, documented differences from the reference (none affects finite, tie-free input):
 *   - ties between equal scores are broken by ascending item id (the reference leaves them to
 *     libstdc++'s heap-select / introsort);
 *   - a NaN candidate score gives a NaN row (the reference's documented rule, hpp:195-197,
 *     which its noise-off branch does not enforce);
 *   - fix_quirks=1 computes Hit@K / RR@K when requested alone (reference: uninitialised, Q2);
 *   - PR-AUC without ROC-AUC is computed on the full order (reference: partial order, Q3).
 *
 * Besides the reference's ten outputs it can return what the reference never exposes and the
 * parity bar asks for: top-K item ids and scores, and the rank of every held-out item.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int rmo_has_openmp(void)
{
#ifdef _OPENMP
    return 1;
#else
    return 0;
#endif
}

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL float
#define FMA fmaf
#define FN(x) CAT(x, _f32)
#include "recometrics_oracle_impl.h"
#undef REAL
#undef FMA
#undef FN

#define REAL double
#define FMA fma
#define FN(x) CAT(x, _f64)
#include "recometrics_oracle_impl.h"
#undef REAL
#undef FMA
#undef FN
