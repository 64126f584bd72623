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

/* ---- doors onto the reference's splitters (declared in src/recometrics_signatures.hpp, defined in
 * src/recometrics_instantiated.cpp:145-387).  The std::vector outputs are copied into caller-sized arrays:
 * pointer arrays m+1, index / value arrays nnz, users_test m.  sizes[] receives the vector lengths:
 * [0] users_test, [1] rem_p, [2] rem_i, [3] train_p, [4] train_i, [5] test_p, [6] test_i.
 * Returns 0, or 1 when the reference threw (what() is copied into err[256]). ---- */
#include <cstring>
#include <exception>
#include <vector>

namespace {
template <class V, class T> void put(const std::vector<V> &v, T *dst) { if (!v.empty() && dst) std::memcpy(dst, v.data(), v.size() * sizeof(V)); }
void say(char *err, const char *what) { if (err) { std::strncpy(err, what, 255); err[255] = 0; } }
}

extern "C" {

#define RMREF_SPLIT_ALL(SUFFIX, REAL, FN)                                                                              \
int rmref_split_selected_users_##SUFFIX(const int32_t *Xp, const int32_t *Xi, const REAL *Xv, int32_t m, int32_t n,    \
        double test_fraction, uint64_t seed, int32_t *trp, int32_t *tri, REAL *trv, int32_t *tep, int32_t *tei,        \
        REAL *tev, int64_t *sizes, char *err)                                                                          \
{                                                                                                                      \
    try {                                                                                                              \
        std::vector<int32_t> a, b, d, e; std::vector<REAL> c, f;                                                       \
        FN(Xp, Xi, Xv, m, n, a, b, c, d, e, f, test_fraction, seed);                                                   \
        put(a, trp); put(b, tri); put(c, trv); put(d, tep); put(e, tei); put(f, tev);                                  \
        sizes[0] = 0; sizes[1] = 0; sizes[2] = 0; sizes[3] = (int64_t)a.size(); sizes[4] = (int64_t)b.size();          \
        sizes[5] = (int64_t)d.size(); sizes[6] = (int64_t)e.size();                                                    \
    } catch (const std::exception &ex) { say(err, ex.what()); return 1; }                                              \
    return 0;                                                                                                          \
}
RMREF_SPLIT_ALL(f32, float, split_data_selected_users_float)
RMREF_SPLIT_ALL(f64, double, split_data_selected_users_double)

#define RMREF_SPLIT_USERS(SUFFIX, REAL, FN_SEP, FN_JOIN)                                                               \
int rmref_split_users_##SUFFIX(const int32_t *Xp, const int32_t *Xi, const REAL *Xv, int32_t m, int32_t n,             \
        int32_t n_users_test, double test_fraction, int consider_cold_start, int32_t min_items_pool,                   \
        int32_t min_pos_test, uint64_t seed, int joined, int32_t *users_test, int32_t *rep, int32_t *rei, REAL *rev,   \
        int32_t *trp, int32_t *tri, REAL *trv, int32_t *tep, int32_t *tei, REAL *tev, int64_t *sizes, char *err)       \
{                                                                                                                      \
    try {                                                                                                              \
        std::vector<int32_t> ut, rp, ri, a, b, d, e; std::vector<REAL> rv, c, f;                                       \
        if (joined) FN_JOIN(Xp, Xi, Xv, m, n, ut, a, b, c, d, e, f, n_users_test, test_fraction,                       \
                            consider_cold_start != 0, min_items_pool, min_pos_test, seed);                             \
        else FN_SEP(Xp, Xi, Xv, m, n, ut, rp, ri, rv, a, b, c, d, e, f, n_users_test, test_fraction,                   \
                    consider_cold_start != 0, min_items_pool, min_pos_test, seed);                                     \
        put(ut, users_test); put(rp, rep); put(ri, rei); put(rv, rev);                                                 \
        put(a, trp); put(b, tri); put(c, trv); put(d, tep); put(e, tei); put(f, tev);                                  \
        sizes[0] = (int64_t)ut.size(); sizes[1] = (int64_t)rp.size(); sizes[2] = (int64_t)ri.size();                   \
        sizes[3] = (int64_t)a.size(); sizes[4] = (int64_t)b.size(); sizes[5] = (int64_t)d.size();                      \
        sizes[6] = (int64_t)e.size();                                                                                  \
    } catch (const std::exception &ex) { say(err, ex.what()); return 1; }                                              \
    return 0;                                                                                                          \
}
RMREF_SPLIT_USERS(f32, float, split_data_separate_users_float, split_data_joined_users_float)
RMREF_SPLIT_USERS(f64, double, split_data_separate_users_double, split_data_joined_users_double)

} /* extern "C" */
