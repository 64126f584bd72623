/* recometrics_b200_shim.hpp -- header-only C++ adapter between the reference's bindings and the
 * C-ABI of librecometrics_b200.so.
 *
 * It defines, with the reference's own names and parameter lists,
 *
 *   calc_metrics_float / calc_metrics_double   (reference: src/recometrics_signatures.hpp:48-98,
 *                                               defined in src/recometrics_instantiated.cpp:43-143,
 *                                               called by recometrics/wrapper.pyx:282-304, :381-403)
 *   template <class real_t> calc_metrics(...)  (reference: src/recometrics.hpp:359-385,
 *                                               called by src/Rwrapper.cpp:250-274)
 *   get_has_openmp()                           (reference: src/recometrics_signatures.hpp:46)
 *
 * so that wrapper.pyx builds unmodified when this header replaces recometrics_signatures.hpp and
 * recometrics_instantiated.cpp is dropped from the sources, and Rwrapper.cpp swaps one #include.
 * See INTEGRATION.md for the two build recipes.
 *
 * Error behaviour follows the reference: a failing call throws (std::runtime_error for an
 * interrupt as in src/recometrics.hpp:171, std::bad_alloc for memory), which Cython's `except +`
 * (wrapper.pyx:62) and Rcpp's BEGIN_RCPP/END_RCPP turn into the host language's exception.
 * Include it in exactly ONE translation unit per binary (it defines non-template functions), as
 * the reference's own header requires (src/recometrics.hpp:115-125 defines globals).
 */
#ifndef RECOMETRICS_B200_SHIM_HPP
#define RECOMETRICS_B200_SHIM_HPP

#include <cstddef>
#include <cstdint>
#include <new>
#include <stdexcept>
#include <string>

#include "recometrics_b200.h"

#ifndef restrict
#   if defined(__GNUG__) || defined(__GNUC__) || defined(_MSC_VER) || defined(__clang__) || defined(__INTEL_COMPILER)
#       define restrict __restrict
#       define RMB200_SHIM_DEFINED_RESTRICT
#   else
#       define restrict
#       define RMB200_SHIM_DEFINED_RESTRICT
#   endif
#endif

namespace rmb200_shim {

inline void raise_for_status(const int rc)
{
    if (rc == RMB200_OK) return;
    const std::string msg = std::string("recometrics_b200: ") + rmb200_last_error();
    switch (rc) {
        case RMB200_ERR_OOM: throw std::bad_alloc();
        case RMB200_ERR_BAD_ARG: throw std::invalid_argument(msg);
        case RMB200_ERR_INTERRUPTED: throw std::runtime_error("Error: procedure was interrupted.\n");
        default: throw std::runtime_error(msg);
    }
}

inline int call(const float *A, size_t lda, const float *B, size_t ldb, int32_t m, int32_t n, int32_t k,
                const int32_t *trp, const int32_t *tri, const int32_t *tep, const int32_t *tei, const float *tev,
                int32_t k_metrics, bool cumulative, bool noise,
                float *p, float *tp, float *r, float *ap, float *tap, float *ndcg, float *hit, float *rr,
                float *roc, float *pr, bool ccs, int32_t mip, int32_t mpt, int32_t nthreads, uint64_t seed)
{
    return rmb200_calc_metrics_f32(A, lda, B, ldb, m, n, k, trp, tri, tep, tei, tev, k_metrics, cumulative, noise,
                                   p, tp, r, ap, tap, ndcg, hit, rr, roc, pr, ccs, mip, mpt, nthreads, seed);
}

inline int call(const double *A, size_t lda, const double *B, size_t ldb, int32_t m, int32_t n, int32_t k,
                const int32_t *trp, const int32_t *tri, const int32_t *tep, const int32_t *tei, const double *tev,
                int32_t k_metrics, bool cumulative, bool noise,
                double *p, double *tp, double *r, double *ap, double *tap, double *ndcg, double *hit, double *rr,
                double *roc, double *pr, bool ccs, int32_t mip, int32_t mpt, int32_t nthreads, uint64_t seed)
{
#ifdef RMB200_SHIM_NAN_BITS
    /* R builds: undefined metrics are NA_REAL, not a plain NaN (src/recometrics.hpp:75-80 under _FOR_R).  Compile the
     * wrapper with -DRMB200_SHIM_NAN_BITS=0x7FF00000000007A2ull (the bit pattern of NA_REAL). */
    rmb200_extra_t ex = rmb200_extra_t();
    ex.struct_size = (int32_t)sizeof(ex);
    ex.device = -1;
    ex.has_nan_bits = 1;
    ex.nan_bits = (uint64_t)(RMB200_SHIM_NAN_BITS);
    return rmb200_calc_metrics_ex_f64(A, lda, B, ldb, m, n, k, trp, tri, tep, tei, tev, k_metrics, cumulative, noise,
                                      p, tp, r, ap, tap, ndcg, hit, rr, roc, pr, ccs, mip, mpt, nthreads, seed, NULL, &ex);
#else
    return rmb200_calc_metrics_f64(A, lda, B, ldb, m, n, k, trp, tri, tep, tei, tev, k_metrics, cumulative, noise,
                                   p, tp, r, ap, tap, ndcg, hit, rr, roc, pr, ccs, mip, mpt, nthreads, seed);
#endif
}

}  // namespace rmb200_shim

/* src/recometrics_signatures.hpp:46 -- "can this build run in parallel": here, "is there a GPU". */
inline bool get_has_openmp() { return rmb200_device_count() > 0; }

/* src/recometrics.hpp:359-385 (the template Rwrapper.cpp:250-274 instantiates for float / double). */
template <class real_t>
void calc_metrics
(
    const real_t *restrict A, const size_t lda, const real_t *restrict B, const size_t ldb,
    const int32_t m, const int32_t n, const int32_t k,
    const int32_t *restrict Xtrain_csr_p, const int32_t *restrict Xtrain_csr_i,
    const int32_t *restrict Xtest_csr_p, int32_t *restrict Xtest_csr_i, const real_t *restrict Xtest_csr,
    const int32_t k_metrics,
    const bool cumulative,
    const bool break_ties_with_noise,
    real_t *restrict p_at_k,
    real_t *restrict tp_at_k,
    real_t *restrict r_at_k,
    real_t *restrict ap_at_k,
    real_t *restrict tap_at_k,
    real_t *restrict ndcg_at_k,
    real_t *restrict hit_at_k,
    real_t *restrict rr_at_k,
    real_t *restrict roc_auc,
    real_t *restrict pr_auc,
    const bool consider_cold_start,
    int32_t min_items_pool,
    int32_t min_pos_test,
    int32_t nthreads,
    uint64_t seed
)
{
    rmb200_shim::raise_for_status(rmb200_shim::call(
        A, lda, B, ldb, m, n, k, Xtrain_csr_p, Xtrain_csr_i, Xtest_csr_p, Xtest_csr_i, Xtest_csr,
        k_metrics, cumulative, break_ties_with_noise,
        p_at_k, tp_at_k, r_at_k, ap_at_k, tap_at_k, ndcg_at_k, hit_at_k, rr_at_k, roc_auc, pr_auc,
        consider_cold_start, min_items_pool, min_pos_test, nthreads, seed));
}

/* src/recometrics_signatures.hpp:48-72 / src/recometrics_instantiated.cpp:43-92 */
inline void calc_metrics_double
(
    const double *restrict A, const size_t lda, const double *restrict B, const size_t ldb,
    const int32_t m, const int32_t n, const int32_t k,
    const int32_t *restrict Xtrain_csr_p, const int32_t *restrict Xtrain_csr_i,
    const int32_t *restrict Xtest_csr_p, int32_t *restrict Xtest_csr_i, const double *restrict Xtest_csr,
    const int32_t k_metrics,
    const bool cumulative,
    const bool break_ties_with_noise,
    double *restrict p_at_k,
    double *restrict tp_at_k,
    double *restrict r_at_k,
    double *restrict ap_at_k,
    double *restrict tap_at_k,
    double *restrict ndcg_at_k,
    double *restrict hit_at_k,
    double *restrict rr_at_k,
    double *restrict roc_auc,
    double *restrict pr_auc,
    const bool consider_cold_start,
    int32_t min_items_pool,
    int32_t min_pos_test,
    int32_t nthreads,
    uint64_t seed
)
{
    calc_metrics<double>(A, lda, B, ldb, m, n, k, Xtrain_csr_p, Xtrain_csr_i, Xtest_csr_p, Xtest_csr_i, Xtest_csr,
                         k_metrics, cumulative, break_ties_with_noise,
                         p_at_k, tp_at_k, r_at_k, ap_at_k, tap_at_k, ndcg_at_k, hit_at_k, rr_at_k, roc_auc, pr_auc,
                         consider_cold_start, min_items_pool, min_pos_test, nthreads, seed);
}

/* src/recometrics_signatures.hpp:73-98 / src/recometrics_instantiated.cpp:94-143 */
inline void calc_metrics_float
(
    const float *restrict A, const size_t lda, const float *restrict B, const size_t ldb,
    const int32_t m, const int32_t n, const int32_t k,
    const int32_t *restrict Xtrain_csr_p, const int32_t *restrict Xtrain_csr_i,
    const int32_t *restrict Xtest_csr_p, int32_t *restrict Xtest_csr_i, const float *restrict Xtest_csr,
    const int32_t k_metrics,
    const bool cumulative,
    const bool break_ties_with_noise,
    float *restrict p_at_k,
    float *restrict tp_at_k,
    float *restrict r_at_k,
    float *restrict ap_at_k,
    float *restrict tap_at_k,
    float *restrict ndcg_at_k,
    float *restrict hit_at_k,
    float *restrict rr_at_k,
    float *restrict roc_auc,
    float *restrict pr_auc,
    const bool consider_cold_start,
    int32_t min_items_pool,
    int32_t min_pos_test,
    int32_t nthreads,
    uint64_t seed
)
{
    calc_metrics<float>(A, lda, B, ldb, m, n, k, Xtrain_csr_p, Xtrain_csr_i, Xtest_csr_p, Xtest_csr_i, Xtest_csr,
                        k_metrics, cumulative, break_ties_with_noise,
                        p_at_k, tp_at_k, r_at_k, ap_at_k, tap_at_k, ndcg_at_k, hit_at_k, rr_at_k, roc_auc, pr_auc,
                        consider_cold_start, min_items_pool, min_pos_test, nthreads, seed);
}

#ifdef RMB200_SHIM_DEFINED_RESTRICT
#   undef restrict
#   undef RMB200_SHIM_DEFINED_RESTRICT
#endif

#endif /* RECOMETRICS_B200_SHIM_HPP */
