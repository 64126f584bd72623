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
 *   split_data_{selected,separate,joined}_users_float / _double
 *                                              (reference: src/recometrics_signatures.hpp:100-221,
 *                                               defined in src/recometrics_instantiated.cpp:145-387,
 *                                               called by recometrics/wrapper.pyx:541-762)
 *   template <class real_t> split_data_{selected,separate,joined}_users(...)
 *                                              (reference: src/recometrics.hpp:1015-1106, :1201-1322, :1439-1505,
 *                                               called by src/Rwrapper.cpp:452, :529, :570)
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
#include <vector>

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
        case RMB200_ERR_RUNTIME: throw std::runtime_error(rmb200_last_error());   /* the reference's own message */
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


#ifndef RMB200_SHIM_NO_SPLITTERS   /* define it to keep the reference's CPU splitters (src/recometrics_instantiated.cpp) */
/* ---------------------------------------------------------------------------------------------------------------------
 * Train/test splitters.  The reference fills std::vectors; the C-ABI hands out one rmb200_split_t of library-owned arrays
 * that are copied into the caller's vectors and released (also when the copy throws).
 * ------------------------------------------------------------------------------------------------------------------- */
namespace rmb200_shim {

struct split_guard {
    rmb200_split_t s;
    split_guard() : s() {}
    ~split_guard() { rmb200_split_free(&s); }
};

template <class real_t>
inline void take(const rmb200_csr_t &c, std::vector<int32_t> &p, std::vector<int32_t> &i, std::vector<real_t> &v)
{
    if (!c.indptr) return;                      /* (split_data_selected_users on m = 0 leaves its outputs untouched) */
    p.assign(c.indptr, c.indptr + (size_t)c.rows + 1);
    i.assign(c.indices, c.indices + (size_t)c.nnz);
    v.assign((const real_t *)c.values, (const real_t *)c.values + (size_t)c.nnz);
}

inline int split_selected(const int32_t *p, const int32_t *i, const float *v, int32_t m, int32_t n, double f, uint64_t seed, rmb200_split_t *o)
{ return rmb200_split_selected_users_f32(p, i, v, m, n, f, seed, -1, o); }
inline int split_selected(const int32_t *p, const int32_t *i, const double *v, int32_t m, int32_t n, double f, uint64_t seed, rmb200_split_t *o)
{ return rmb200_split_selected_users_f64(p, i, v, m, n, f, seed, -1, o); }
inline int split_users(bool joined, const int32_t *p, const int32_t *i, const float *v, int32_t m, int32_t n, int32_t nu, double f,
                       bool cold, int32_t pool, int32_t pos, uint64_t seed, rmb200_split_t *o)
{
    return joined ? rmb200_split_joined_users_f32(p, i, v, m, n, nu, f, cold, pool, pos, seed, -1, o)
                  : rmb200_split_separate_users_f32(p, i, v, m, n, nu, f, cold, pool, pos, seed, -1, o);
}
inline int split_users(bool joined, const int32_t *p, const int32_t *i, const double *v, int32_t m, int32_t n, int32_t nu, double f,
                       bool cold, int32_t pool, int32_t pos, uint64_t seed, rmb200_split_t *o)
{
    return joined ? rmb200_split_joined_users_f64(p, i, v, m, n, nu, f, cold, pool, pos, seed, -1, o)
                  : rmb200_split_separate_users_f64(p, i, v, m, n, nu, f, cold, pool, pos, seed, -1, o);
}

}  /* namespace rmb200_shim */

/* src/recometrics.hpp:1015-1030 */
template <class real_t>
void split_data_selected_users
(
    const int32_t *restrict X_csr_p,
    const int32_t *restrict X_csr_i,
    const real_t *restrict X_csr,
    const int32_t m, const int32_t n,
    std::vector<int32_t> &Xtrain_csr_p,
    std::vector<int32_t> &Xtrain_csr_i,
    std::vector<real_t> &Xtrain_csr,
    std::vector<int32_t> &Xtest_csr_p,
    std::vector<int32_t> &Xtest_csr_i,
    std::vector<real_t> &Xtest_csr,
    const double test_fraction,
    uint64_t seed
)
{
    rmb200_shim::split_guard g;
    rmb200_shim::raise_for_status(rmb200_shim::split_selected(X_csr_p, X_csr_i, X_csr, m, n, test_fraction, seed, &g.s));
    rmb200_shim::take<real_t>(g.s.train, Xtrain_csr_p, Xtrain_csr_i, Xtrain_csr);
    rmb200_shim::take<real_t>(g.s.test, Xtest_csr_p, Xtest_csr_i, Xtest_csr);
}

/* src/recometrics.hpp:1201-1224 */
template <class real_t>
void split_data_separate_users
(
    const int32_t *restrict X_csr_p,
    const int32_t *restrict X_csr_i,
    const real_t *restrict X_csr,
    int32_t m, int32_t n,
    std::vector<int32_t> &users_test,
    std::vector<int32_t> &Xrem_csr_p,
    std::vector<int32_t> &Xrem_csr_i,
    std::vector<real_t> &Xrem_csr,
    std::vector<int32_t> &Xtrain_csr_p,
    std::vector<int32_t> &Xtrain_csr_i,
    std::vector<real_t> &Xtrain_csr,
    std::vector<int32_t> &Xtest_csr_p,
    std::vector<int32_t> &Xtest_csr_i,
    std::vector<real_t> &Xtest_csr,
    const int32_t n_users_test,
    const double test_fraction,
    const bool consider_cold_start,
    const int32_t min_items_pool,
    const int32_t min_pos_test,
    uint64_t seed
)
{
    rmb200_shim::split_guard g;
    rmb200_shim::raise_for_status(rmb200_shim::split_users(false, X_csr_p, X_csr_i, X_csr, m, n, n_users_test, test_fraction,
                                                           consider_cold_start, min_items_pool, min_pos_test, seed, &g.s));
    users_test.assign(g.s.users_test, g.s.users_test + g.s.n_users_test);
    rmb200_shim::take<real_t>(g.s.rem, Xrem_csr_p, Xrem_csr_i, Xrem_csr);
    rmb200_shim::take<real_t>(g.s.train, Xtrain_csr_p, Xtrain_csr_i, Xtrain_csr);
    rmb200_shim::take<real_t>(g.s.test, Xtest_csr_p, Xtest_csr_i, Xtest_csr);
}

/* src/recometrics.hpp:1439-1459 */
template <class real_t>
void split_data_joined_users
(
    const int32_t *restrict X_csr_p,
    const int32_t *restrict X_csr_i,
    const real_t *restrict X_csr,
    int32_t m, int32_t n,
    std::vector<int32_t> &users_test,
    std::vector<int32_t> &Xtrain_csr_p,
    std::vector<int32_t> &Xtrain_csr_i,
    std::vector<real_t> &Xtrain_csr,
    std::vector<int32_t> &Xtest_csr_p,
    std::vector<int32_t> &Xtest_csr_i,
    std::vector<real_t> &Xtest_csr,
    const int32_t n_users_test,
    const double test_fraction,
    const bool consider_cold_start,
    const int32_t min_items_pool,
    const int32_t min_pos_test,
    uint64_t seed
)
{
    rmb200_shim::split_guard g;
    rmb200_shim::raise_for_status(rmb200_shim::split_users(true, X_csr_p, X_csr_i, X_csr, m, n, n_users_test, test_fraction,
                                                           consider_cold_start, min_items_pool, min_pos_test, seed, &g.s));
    users_test.assign(g.s.users_test, g.s.users_test + g.s.n_users_test);
    rmb200_shim::take<real_t>(g.s.train, Xtrain_csr_p, Xtrain_csr_i, Xtrain_csr);
    rmb200_shim::take<real_t>(g.s.test, Xtest_csr_p, Xtest_csr_i, Xtest_csr);
}

/* src/recometrics_signatures.hpp:100-221 / src/recometrics_instantiated.cpp:145-387.  The float entry points declare the
 * fraction `const float`: it is rounded to float on the way in, exactly as there. */
#define RMB200_SHIM_SPLIT_LINKAGE(SUFFIX, REAL, FRAC)                                                                          \
inline void split_data_selected_users_##SUFFIX(const int32_t *restrict X_csr_p, const int32_t *restrict X_csr_i,              \
    const REAL *restrict X_csr, const int32_t m, const int32_t n, std::vector<int32_t> &Xtrain_csr_p,                          \
    std::vector<int32_t> &Xtrain_csr_i, std::vector<REAL> &Xtrain_csr, std::vector<int32_t> &Xtest_csr_p,                      \
    std::vector<int32_t> &Xtest_csr_i, std::vector<REAL> &Xtest_csr, const FRAC test_fraction, uint64_t seed)                  \
{                                                                                                                              \
    split_data_selected_users<REAL>(X_csr_p, X_csr_i, X_csr, m, n, Xtrain_csr_p, Xtrain_csr_i, Xtrain_csr, Xtest_csr_p,        \
                                    Xtest_csr_i, Xtest_csr, test_fraction, seed);                                              \
}                                                                                                                              \
inline void split_data_separate_users_##SUFFIX(const int32_t *restrict X_csr_p, const int32_t *restrict X_csr_i,              \
    const REAL *restrict X_csr, int32_t m, int32_t n, std::vector<int32_t> &users_test, std::vector<int32_t> &Xrem_csr_p,      \
    std::vector<int32_t> &Xrem_csr_i, std::vector<REAL> &Xrem_csr, std::vector<int32_t> &Xtrain_csr_p,                         \
    std::vector<int32_t> &Xtrain_csr_i, std::vector<REAL> &Xtrain_csr, std::vector<int32_t> &Xtest_csr_p,                      \
    std::vector<int32_t> &Xtest_csr_i, std::vector<REAL> &Xtest_csr, const int32_t n_users_test, const FRAC test_fraction,     \
    const bool consider_cold_start, const int32_t min_items_pool, const int32_t min_pos_test, uint64_t seed)                   \
{                                                                                                                              \
    split_data_separate_users<REAL>(X_csr_p, X_csr_i, X_csr, m, n, users_test, Xrem_csr_p, Xrem_csr_i, Xrem_csr, Xtrain_csr_p, \
                                    Xtrain_csr_i, Xtrain_csr, Xtest_csr_p, Xtest_csr_i, Xtest_csr, n_users_test,               \
                                    test_fraction, consider_cold_start, min_items_pool, min_pos_test, seed);                   \
}                                                                                                                              \
inline void split_data_joined_users_##SUFFIX(const int32_t *restrict X_csr_p, const int32_t *restrict X_csr_i,                \
    const REAL *restrict X_csr, int32_t m, int32_t n, std::vector<int32_t> &users_test, std::vector<int32_t> &Xtrain_csr_p,    \
    std::vector<int32_t> &Xtrain_csr_i, std::vector<REAL> &Xtrain_csr, std::vector<int32_t> &Xtest_csr_p,                      \
    std::vector<int32_t> &Xtest_csr_i, std::vector<REAL> &Xtest_csr, const int32_t n_users_test, const FRAC test_fraction,     \
    const bool consider_cold_start, const int32_t min_items_pool, const int32_t min_pos_test, uint64_t seed)                   \
{                                                                                                                              \
    split_data_joined_users<REAL>(X_csr_p, X_csr_i, X_csr, m, n, users_test, Xtrain_csr_p, Xtrain_csr_i, Xtrain_csr,           \
                                  Xtest_csr_p, Xtest_csr_i, Xtest_csr, n_users_test, test_fraction, consider_cold_start,       \
                                  min_items_pool, min_pos_test, seed);                                                         \
}
RMB200_SHIM_SPLIT_LINKAGE(double, double, double)
RMB200_SHIM_SPLIT_LINKAGE(float, float, float)
#undef RMB200_SHIM_SPLIT_LINKAGE
#endif /* RMB200_SHIM_NO_SPLITTERS */

#ifdef RMB200_SHIM_DEFINED_RESTRICT
#   undef restrict
#   undef RMB200_SHIM_DEFINED_RESTRICT
#endif

#endif /* RECOMETRICS_B200_SHIM_HPP */
