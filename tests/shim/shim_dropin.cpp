// Drop-in check of include/recometrics_b200_shim.hpp (tests/test_capi_host.py builds and runs this):
//  * with -DHAVE_REFERENCE_HEADER the reference's own declarations (src/recometrics_signatures.hpp:46-98) are included
//    FIRST: the shim's definitions must then be definitions of exactly those functions -- taking their addresses with
//    the reference's parameter lists is ambiguous or fails to link otherwise;
//  * the calls go through the C-ABI of librecometrics_b200.so; without a CUDA device they must throw (no CPU path).
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <vector>
#ifdef HAVE_REFERENCE_HEADER
#include "recometrics_signatures.hpp"
#endif
#include "recometrics_b200_shim.hpp"

typedef void (*metrics_f32_t)(const float*, const size_t, const float*, const size_t, const int32_t, const int32_t, const int32_t,
                              const int32_t*, const int32_t*, const int32_t*, int32_t*, const float*, const int32_t, const bool, const bool,
                              float*, float*, float*, float*, float*, float*, float*, float*, float*, float*,
                              const bool, int32_t, int32_t, int32_t, uint64_t);
typedef void (*metrics_f64_t)(const double*, const size_t, const double*, const size_t, const int32_t, const int32_t, const int32_t,
                              const int32_t*, const int32_t*, const int32_t*, int32_t*, const double*, const int32_t, const bool, const bool,
                              double*, double*, double*, double*, double*, double*, double*, double*, double*, double*,
                              const bool, int32_t, int32_t, int32_t, uint64_t);

// the splitters' linkage functions (src/recometrics_signatures.hpp:100-221) and the templates src/Rwrapper.cpp:452-570 instantiates
typedef std::vector<int32_t> ivec;
typedef void (*split_all_f64_t)(const int32_t*, const int32_t*, const double*, const int32_t, const int32_t, ivec&, ivec&, std::vector<double>&,
                                ivec&, ivec&, std::vector<double>&, const double, uint64_t);
typedef void (*split_all_f32_t)(const int32_t*, const int32_t*, const float*, const int32_t, const int32_t, ivec&, ivec&, std::vector<float>&,
                                ivec&, ivec&, std::vector<float>&, const float, uint64_t);
typedef void (*split_sep_f64_t)(const int32_t*, const int32_t*, const double*, int32_t, int32_t, ivec&, ivec&, ivec&, std::vector<double>&,
                                ivec&, ivec&, std::vector<double>&, ivec&, ivec&, std::vector<double>&, const int32_t, const double,
                                const bool, const int32_t, const int32_t, uint64_t);
typedef void (*split_sep_f32_t)(const int32_t*, const int32_t*, const float*, int32_t, int32_t, ivec&, ivec&, ivec&, std::vector<float>&,
                                ivec&, ivec&, std::vector<float>&, ivec&, ivec&, std::vector<float>&, const int32_t, const float,
                                const bool, const int32_t, const int32_t, uint64_t);
typedef void (*split_join_f64_t)(const int32_t*, const int32_t*, const double*, int32_t, int32_t, ivec&, ivec&, ivec&, std::vector<double>&,
                                 ivec&, ivec&, std::vector<double>&, const int32_t, const double, const bool, const int32_t, const int32_t, uint64_t);
typedef void (*split_join_f32_t)(const int32_t*, const int32_t*, const float*, int32_t, int32_t, ivec&, ivec&, ivec&, std::vector<float>&,
                                 ivec&, ivec&, std::vector<float>&, const int32_t, const float, const bool, const int32_t, const int32_t, uint64_t);

static int check_splitters()
{
    split_all_f64_t a64 = &split_data_selected_users_double;
    split_all_f32_t a32 = &split_data_selected_users_float;
    split_sep_f64_t s64 = &split_data_separate_users_double;
    split_sep_f32_t s32 = &split_data_separate_users_float;
    split_join_f64_t j64 = &split_data_joined_users_double;
    split_join_f32_t j32 = &split_data_joined_users_float;
    (void)a32; (void)s32; (void)j32;
    // 4 users x 6 items
    const int32_t Xp[5] = {0, 3, 3, 7, 9}, Xi[9] = {0, 2, 5, 1, 2, 3, 4, 0, 5};
    const double Xv[9] = {1., 2., 3., 4., 5., 6., 7., 8., 9.};
    int threw = 0;
    try {
        ivec trp, tri, tep, tei, ut, rp, ri;
        std::vector<double> trv, tev, rv;
        a64(Xp, Xi, Xv, 4, 6, trp, tri, trv, tep, tei, tev, 0.5, 1);
        if (trp.size() != 5 || tep.size() != 5 || tri.size() + tei.size() != 9 || tep[4] != 2 + 0 + 2 + 1) return 5;
        s64(Xp, Xi, Xv, 4, 6, ut, rp, ri, rv, trp, tri, trv, tep, tei, tev, 2, 0.5, false, 2, 1, 1);
        if (ut.size() != 2 || rp.size() != 3 || trp.size() != 3 || tep.size() != 3) return 6;
        j64(Xp, Xi, Xv, 4, 6, ut, trp, tri, trv, tep, tei, tev, 2, 0.5, false, 2, 1, 1);
        if (ut.size() != 2 || trp.size() != 5 || tep.size() != 3 || tri.size() + tei.size() != 9) return 7;
        // the template form of the Rcpp wrapper, and the reference's refusal with its own message
        split_data_selected_users<double>(Xp, Xi, Xv, 4, 6, trp, tri, trv, tep, tei, tev, 0.5, 1);
        try {
            split_data_separate_users<double>(Xp, Xi, Xv, 4, 6, ut, rp, ri, rv, trp, tri, trv, tep, tei, tev, 5, 0.5, false, 2, 1, 1);
            return 8;
        } catch (const std::runtime_error& e) {
            if (std::strcmp(e.what(), "Target number of test users is larger than available users.\n") != 0) return 9;
        }
        std::printf("split ok users=%d,%d\n", ut[0], ut[1]);
    } catch (const std::runtime_error& e) {
        threw = 1;
        std::printf("split threw runtime_error: %s\n", e.what());
    }
    if (!get_has_openmp() && !threw) return 10;    // no device and no exception: a CPU path would be a bug
    return 0;
}

int main()
{
    if (const int rc = check_splitters()) return rc;
    metrics_f32_t f32 = &calc_metrics_float;
    metrics_f64_t f64 = &calc_metrics_double;
    // 3 users x 4 items, 2 factors; every user holds out one item
    const float A[6] = {1.f, 0.f, 0.f, 1.f, 1.f, 1.f}, B[8] = {1.f, 0.f, 0.f, 1.f, .5f, .5f, -1.f, 2.f};
    const double Ad[6] = {1., 0., 0., 1., 1., 1.}, Bd[8] = {1., 0., 0., 1., .5, .5, -1., 2.};
    const int32_t trp[4] = {0, 0, 0, 0}, tep[4] = {0, 1, 2, 3};
    int32_t tei[3] = {0, 1, 3};
    const float tev[3] = {1.f, 1.f, 1.f};
    const double tevd[3] = {1., 1., 1.};
    float p[3], ndcg[3];
    double pd[3], ndcgd[3];
    std::printf("has_gpu=%d\n", (int)get_has_openmp());
    int threw = 0;
    try {
        f32(A, 2, B, 2, 3, 4, 2, trp, nullptr, tep, tei, tev, 2, false, false, p, nullptr, nullptr, nullptr, nullptr, ndcg, nullptr, nullptr,
            nullptr, nullptr, true, 2, 1, 1, 1);
        f64(Ad, 2, Bd, 2, 3, 4, 2, trp, nullptr, tep, tei, tevd, 2, false, false, pd, nullptr, nullptr, nullptr, nullptr, ndcgd, nullptr, nullptr,
            nullptr, nullptr, true, 2, 1, 1, 1);
        // template form used by the Rcpp wrapper (src/Rwrapper.cpp:250-274)
        calc_metrics<double>(Ad, 2, Bd, 2, 3, 4, 2, trp, nullptr, tep, tei, tevd, 2, false, false, pd, nullptr, nullptr, nullptr, nullptr, ndcgd,
                             nullptr, nullptr, nullptr, nullptr, true, 2, 1, 1, 1);
        // users 0 and 1 rank their held-out item first, user 2 ranks item 3 (score 1) behind items 0..2? scores: 1, 1, 1, 1 -> tie row
        std::printf("ok p=%g,%g ndcg=%g,%g pd=%g\n", p[0], p[1], ndcg[0], ndcg[1], pd[0]);
        if (!(std::fabs(p[0] - 0.5f) < 1e-6f && std::fabs(p[1] - 0.5f) < 1e-6f && std::fabs(pd[0] - 0.5) < 1e-12)) return 3;
#ifdef RMB200_SHIM_NAN_BITS
        {
            // R build of the shim: a user without held-out items gets NA_REAL rows (src/recometrics.hpp:75-80), bit for bit
            const int32_t tep2[4] = {0, 1, 1, 2};              // user 1 holds nothing out
            int32_t tei2[2] = {0, 3};
            const double tev2[2] = {1., 1.};
            double q[3] = {0., 0., 0.};
            calc_metrics<double>(Ad, 2, Bd, 2, 3, 4, 2, trp, nullptr, tep2, tei2, tev2, 2, false, false, q, nullptr, nullptr, nullptr, nullptr,
                                 nullptr, nullptr, nullptr, nullptr, nullptr, true, 2, 1, 1, 1);
            uint64_t bits;
            std::memcpy(&bits, &q[1], 8);
            std::printf("na_bits=%s (%llx)\n", bits == (uint64_t)(RMB200_SHIM_NAN_BITS) ? "ok" : "WRONG", (unsigned long long)bits);
            if (bits != (uint64_t)(RMB200_SHIM_NAN_BITS) || !(std::fabs(q[0] - 0.5) < 1e-12)) return 4;
        }
#endif
    } catch (const std::runtime_error& e) {
        threw = 1;
        std::printf("threw runtime_error: %s\n", e.what());
    }
    if (!get_has_openmp() && !threw) return 2;     // no device and no exception: a CPU path would be a bug
    return 0;
}
