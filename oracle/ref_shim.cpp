/* TEST INFRASTRUCTURE -- not part of the product.
 *
 * extern "C" doorway onto the UNMODIFIED reference, so that tests and the
 * bench's cpu_baseline / --impl reference legs can call it through ctypes.
 *
 * This file contains no reference code: it only includes the reference's own
 * declaration header (src/recometrics_signatures.hpp:46-98) and forwards to
 * calc_metrics_float / calc_metrics_double, which oracle/Makefile compiles
 * from /root/reference/src/recometrics_instantiated.cpp where it lies.
 * The resulting library lands in oracle/_ref/ (git-ignored, shipped by gpurun).
 */
#include "recometrics_signatures.hpp"

extern "C" {

int rmref_has_openmp(void) { return get_has_openmp() ? 1 : 0; }

/* Returns 0 on success, 1 if the reference threw (interrupt / bad_alloc). */
int rmref_calc_metrics_f32(
    const float *A, size_t lda, const float *B, size_t ldb,
    int32_t m, int32_t n, int32_t k,
    const int32_t *Xtrain_p, const int32_t *Xtrain_i,
    const int32_t *Xtest_p, int32_t *Xtest_i, const float *Xtest_v,
    int32_t k_metrics, int cumulative, int break_ties_with_noise,
    float *p_at_k, float *tp_at_k, float *r_at_k, float *ap_at_k, float *tap_at_k,
    float *ndcg_at_k, float *hit_at_k, float *rr_at_k, float *roc_auc, float *pr_auc,
    int consider_cold_start, int32_t min_items_pool, int32_t min_pos_test,
    int32_t nthreads, uint64_t seed)
{
    try {
        calc_metrics_float(A, lda, B, ldb, m, n, k, Xtrain_p, Xtrain_i, Xtest_p, Xtest_i, Xtest_v,
                           k_metrics, cumulative != 0, break_ties_with_noise != 0,
                           p_at_k, tp_at_k, r_at_k, ap_at_k, tap_at_k, ndcg_at_k, hit_at_k, rr_at_k,
                           roc_auc, pr_auc, consider_cold_start != 0, min_items_pool, min_pos_test,
                           nthreads, seed);
    } catch (...) { return 1; }
    return 0;
}

int rmref_calc_metrics_f64(
    const double *A, size_t lda, const double *B, size_t ldb,
    int32_t m, int32_t n, int32_t k,
    const int32_t *Xtrain_p, const int32_t *Xtrain_i,
    const int32_t *Xtest_p, int32_t *Xtest_i, const double *Xtest_v,
    int32_t k_metrics, int cumulative, int break_ties_with_noise,
    double *p_at_k, double *tp_at_k, double *r_at_k, double *ap_at_k, double *tap_at_k,
    double *ndcg_at_k, double *hit_at_k, double *rr_at_k, double *roc_auc, double *pr_auc,
    int consider_cold_start, int32_t min_items_pool, int32_t min_pos_test,
    int32_t nthreads, uint64_t seed)
{
    try {
        calc_metrics_double(A, lda, B, ldb, m, n, k, Xtrain_p, Xtrain_i, Xtest_p, Xtest_i, Xtest_v,
                            k_metrics, cumulative != 0, break_ties_with_noise != 0,
                            p_at_k, tp_at_k, r_at_k, ap_at_k, tap_at_k, ndcg_at_k, hit_at_k, rr_at_k,
                            roc_auc, pr_auc, consider_cold_start != 0, min_items_pool, min_pos_test,
                            nthreads, seed);
    } catch (...) { return 1; }
    return 0;
}

} /* extern "C" */
