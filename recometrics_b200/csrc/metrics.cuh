// metrics.cuh -- per-user metrics from the ranked top-K, the held-out row and the rank buckets.
//
// One thread per user; everything is accumulated sequentially in rank order, in double, with the
// same operations as the reference so that -- given the same ranking -- the results are
// bit-identical:
//   validity of the ranking        /root/reference/src/recometrics.hpp:537-563
//   top-K walk                      :605-708   (hits, AP, DCG, RR; cumulative writes :625-635)
//   post-hoc NaN rules              :750-788
//   ROC-AUC / PR-AUC                :795-865   (from rank counts instead of a sorted walk)
//   NDCG normalisation              :868-961   (incl. the frozen cumulative tail :951-956)
// log2(ix+2) comes from a host table (glibc values) so the discount factors match the host's.
#pragma once
#include "score_select.cuh"

namespace rmb {

template <typename T>
struct MetricsParams {
    int n, K, C, user0, mb, cumulative;
    int want_roc, want_pr, count_ranks;   // count_ranks: the rank-counting pass ran
    int noise;                            // break_ties_with_noise: validity rules of the noise branch (hpp:516-528)
    const int* trp; const int* tep; const int* tei; const T* tev;
    const int* ustatus; const int* uflags;
    const T* cand_score; const int* cand_item; const int* cand_count;
    const unsigned int* auc_cnt; const unsigned long long* umin; const int* pos_perm;
    const double* log2tab;                // [K]: log2(ix + 2)
    T nan_value;
    // outputs, already offset to the batch's first user (row stride: K if cumulative else 1)
    T *p, *tp, *r, *ap, *tap, *ndcg, *hit, *rr, *roc, *pr;
    int* status_out;                      // optional, offset to the batch
    int* topk_items; T* topk_scores;      // optional [mb][K], offset to the batch
    long long* pos_rank;                  // optional [nnz_test] (absolute)
};

template <typename T>
__device__ __forceinline__ void fill_row(T* out, const size_t ul, const int K, const int cumulative, const T v)
{
    if (!out) return;
    if (!cumulative) out[ul] = v;
    else for (int c = 0; c < K; c++) out[ul * (size_t)K + c] = v;
}

template <typename T>
__device__ void all_nan(const MetricsParams<T>& P, const size_t ul)
{
    fill_row(P.p, ul, P.K, P.cumulative, P.nan_value);
    fill_row(P.tp, ul, P.K, P.cumulative, P.nan_value);
    fill_row(P.r, ul, P.K, P.cumulative, P.nan_value);
    fill_row(P.ap, ul, P.K, P.cumulative, P.nan_value);
    fill_row(P.tap, ul, P.K, P.cumulative, P.nan_value);
    fill_row(P.ndcg, ul, P.K, P.cumulative, P.nan_value);
    fill_row(P.hit, ul, P.K, P.cumulative, P.nan_value);
    fill_row(P.rr, ul, P.K, P.cumulative, P.nan_value);
    if (P.roc) P.roc[ul] = P.nan_value;
    if (P.pr) P.pr[ul] = P.nan_value;
}

template <typename T>
__global__ void user_metrics_kernel(const __grid_constant__ MetricsParams<T> P)
{
    const int uli = blockIdx.x * blockDim.x + threadIdx.x;
    if (uli >= P.mb) return;
    const size_t ul = (size_t)uli;
    const int u = P.user0 + uli;
    const int K = P.K, n = P.n;
    const T NaN = P.nan_value;

    if (P.topk_items) for (int c = 0; c < K; c++) P.topk_items[ul * K + c] = -1;
    if (P.topk_scores) for (int c = 0; c < K; c++) P.topk_scores[ul * K + c] = NaN;

    int st = P.ustatus[u];
    if (st != 0) {
        all_nan(P, ul);
        if (P.status_out) P.status_out[uli] = st;
        return;
    }
    const int ntrain = P.trp[u + 1] - P.trp[u];
    const int tp0 = P.tep[u];
    const int npos = P.tep[u + 1] - tp0;
    const int cand = n - ntrain;
    const bool only_ndcg = (ntrain + npos) >= n;             // hpp:479-482
    const bool k_leq_n = cand <= K;                          // hpp:483
    const int walk = K < cand ? K : cand;
    const T* cs = P.cand_score + ul * (size_t)P.C;
    const int* ci = P.cand_item + ul * (size_t)P.C;

    // ---- validity of the ranking (hpp:537-563; NaN anywhere => NaN row, hpp:195-197) ----
    {
        bool bad = (P.uflags[u] & 1) != 0 || P.cand_count[uli] < walk;
        if (!bad) {
            const bool partial_path = ((!P.want_roc || only_ndcg) && K < cand);
            const T pred_max = cs[0];
            T pred_min;
            if (partial_path) pred_min = cs[K - 1];
            else if (cand <= K) pred_min = cs[cand - 1];
            else pred_min = NumTraits<T>::from_orderable(P.umin[u]);   // full order: smallest candidate score
            // noise off: the sorted extremes decide (hpp:541-548 / :555-562).  Noise on: the rule is "all candidates equal"
            // (hpp:527), decided before the noise went in: flag 4 from the scoring stage, or -- full order -- the exact extremes
            const bool equal = !P.noise ? (pred_max == pred_min)
                                        : ((P.uflags[u] & 4) != 0 || (!partial_path && cand > K && pred_max == pred_min));
            bad = (pred_max != pred_max) || (pred_min != pred_min) || isinf(pred_max) || isinf(pred_min) || equal;
        }
        if (bad) {
            all_nan(P, ul);
            if (P.status_out) P.status_out[uli] = 3;
            return;
        }
    }
    if (P.status_out) P.status_out[uli] = 0;
    if (P.topk_items) for (int c = 0; c < walk; c++) P.topk_items[ul * K + c] = ci[c];
    if (P.topk_scores) for (int c = 0; c < walk; c++) P.topk_scores[ul * K + c] = cs[c];

    const int* ti = P.tei + tp0;
    const T* tv = P.tev ? (P.tev + tp0) : nullptr;
    const size_t rs = P.cumulative ? (size_t)K : 1;          // row stride of the top-K outputs
    T* p_u = P.p ? P.p + ul * rs : nullptr;
    T* tp_u = P.tp ? P.tp + ul * rs : nullptr;
    T* r_u = P.r ? P.r + ul * rs : nullptr;
    T* ap_u = P.ap ? P.ap + ul * rs : nullptr;
    T* tap_u = P.tap ? P.tap + ul * rs : nullptr;
    T* ndcg_u = P.ndcg ? P.ndcg + ul * rs : nullptr;
    T* hit_u = P.hit ? P.hit + ul * rs : nullptr;
    T* rr_u = P.rr ? P.rr + ul * rs : nullptr;

    // ---- top-K walk (hpp:605-708).  Unlike the reference (quirk Q2) Hit@K / RR@K requested on their
    //      own are computed too. ----
    int hits = 0;
    double avg_p = 0, dcg = 0;
    int min_rank = INT_MAX;
    bool did_walk = false;
    const bool calc_top = P.p || P.tp || P.r || P.ap || P.tap || P.ndcg || P.hit || P.rr;
    if (calc_top && (!k_leq_n || P.ap || P.tap || P.rr || P.ndcg)) {
        did_walk = true;
        for (int ix = 0; ix < walk; ix++) {
            const int item = ci[ix];
            int lo = 0, hi = npos;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (ti[mid] < item) lo = mid + 1; else hi = mid; }
            if (lo < npos && ti[lo] == item) {
                hits++;
                avg_p += hits / (double)(ix + 1);
                dcg += tv ? ((double)tv[lo] / P.log2tab[ix]) : 0.;
                if (ix < min_rank) min_rank = ix;
            }
            if (P.cumulative) {
                const int tn = (ix + 1) < npos ? (ix + 1) : npos;
                if (p_u) p_u[ix] = (T)(hits / (double)(ix + 1));
                if (tp_u) tp_u[ix] = (T)(hits / (double)tn);
                if (r_u) r_u[ix] = (T)(hits / (double)npos);
                if (ap_u) ap_u[ix] = (T)(avg_p / (double)npos);
                if (tap_u) tap_u[ix] = (T)(avg_p / (double)tn);
                if (ndcg_u) ndcg_u[ix] = (T)dcg;
                if (hit_u) hit_u[ix] = (T)(hits > 0);
                if (rr_u) rr_u[ix] = (T)(hits ? (1. / (double)(min_rank + 1)) : 0.);
            }
            if (!P.cumulative && hits >= cand) break;        // hpp:637-638
        }
        if (!P.cumulative) {
            const int tn = K < npos ? K : npos;
            if (p_u) *p_u = (T)((double)hits / (double)K);
            if (tp_u) *tp_u = (T)((double)hits / (double)tn);
            if (r_u) *r_u = (T)((double)hits / (double)npos);
            if (ap_u) *ap_u = (T)(avg_p / (double)npos);
            if (tap_u) *tap_u = (T)(avg_p / (double)tn);
            if (hit_u) *hit_u = (T)(hits > 0);
            if (rr_u) *rr_u = (T)(hits ? (1. / (double)(min_rank + 1)) : 0.);
        }
    }

    // ---- post-hoc NaN rules (hpp:750-788) ----
    if (k_leq_n) {
        if (!P.cumulative) {
            if (p_u) *p_u = NaN;
            if (tp_u) *tp_u = NaN;
            if (r_u) *r_u = NaN;
            if (hit_u) *hit_u = NaN;
        } else if (!did_walk) {
            for (int c = 0; c < K; c++) {
                if (p_u) p_u[c] = NaN;
                if (tp_u) tp_u[c] = NaN;
                if (r_u) r_u[c] = NaN;
                if (hit_u) hit_u[c] = NaN;
            }
        }
        if (!did_walk) {   // outputs the reference leaves untouched here; keep every element defined
            const int cnt = P.cumulative ? K : 1;
            for (int c = 0; c < cnt; c++) {
                if (ap_u) ap_u[c] = NaN;
                if (tap_u) tap_u[c] = NaN;
                if (rr_u) rr_u[c] = NaN;
                if (ndcg_u) ndcg_u[c] = NaN;
            }
        }
    } else if (only_ndcg) {
        const int cnt = P.cumulative ? K : 1;
        for (int c = 0; c < cnt; c++) {
            if (p_u) p_u[c] = NaN;
            if (tp_u) tp_u[c] = NaN;
            if (r_u) r_u[c] = NaN;
            if (ap_u) ap_u[c] = NaN;
            if (tap_u) tap_u[c] = NaN;
            if (hit_u) hit_u[c] = NaN;
            if (rr_u) rr_u[c] = NaN;
        }
    }

    // ---- ROC-AUC / PR-AUC (hpp:795-865) from the rank counts ----
    if (P.roc || P.pr || P.pos_rank) {
        if (only_ndcg || !P.count_ranks) {
            if (P.roc) P.roc[ul] = NaN;
            if (P.pr) P.pr[ul] = NaN;
        } else {
            // above[j] = number of candidates scoring strictly above the j-th smallest held-out score
            // (counted by score_select_kernel); walk the held-out items from the best down.
            const unsigned int* ab = P.auc_cnt + (size_t)tp0;
            unsigned long long prev_rank = 0, sum_ranks = 0;
            double ap_full = 0;
            for (int h = 1; h <= npos; h++) {
                const int i = npos - h + 1;
                unsigned long long rank = (unsigned long long)ab[i - 1] + 1;
                if (rank <= prev_rank) rank = prev_rank + 1;   // tied held-out scores take consecutive ranks
                prev_rank = rank;
                sum_ranks += rank;
                ap_full += (double)h / (double)rank;
                if (P.pos_rank) P.pos_rank[(size_t)tp0 + P.pos_perm[tp0 + i - 1]] = (long long)rank;
            }
            const unsigned long long np = (unsigned long long)npos;
            const unsigned long long nneg = (unsigned long long)cand - np;
            if (P.roc)   // hpp:821-822 (long double there; double here, difference < 1e-15)
                P.roc[ul] = (T)(1. - (double)(sum_ranks - (np * (np + 1)) / 2) / (double)(np * nneg));
            if (P.pr) P.pr[ul] = (T)(ap_full / (double)npos);
        }
    }

    // ---- NDCG normalisation (hpp:868-961) ----
    if (ndcg_u && did_walk) {
        const int L = K < npos ? K : npos;
        // what partial_sort + the vmax/vmin checks (hpp:870-887) decide, from one scan
        bool has_nan = false;
        int n_neg_inf = 0;
        T vmax = -NumTraits<T>::inf();
        for (int j = 0; j < npos; j++) {
            const T v = tv[j];
            has_nan |= (v != v);
            n_neg_inf += (v == -NumTraits<T>::inf());
            if (v > vmax) vmax = v;
        }
        const bool bad = has_nan || isinf(vmax) || vmax <= 0 || (npos - n_neg_inf) < L;
        if (bad) {
            const int cnt = P.cumulative ? K : 1;
            for (int c = 0; c < cnt; c++) ndcg_u[c] = NaN;
            return;
        }
        // ideal DCG: the user's values in descending order (ties by position), selected one by one
        double idcg = 0;
        T prev_v = 0;
        int prev_j = -1;
        int ix = 0;
        bool stopped = false;
        for (; ix < L; ix++) {
            T best_v = 0;
            int best_j = -1;
            for (int j = 0; j < npos; j++) {
                const T v = tv[j];
                const bool after_prev = (prev_j < 0) || (v < prev_v) || (v == prev_v && j > prev_j);
                if (after_prev && (best_j < 0 || v > best_v)) { best_v = v; best_j = j; }
            }
            prev_v = best_v;
            prev_j = best_j;
            const double val = (double)best_v;
            if (!P.cumulative) {
                if (val <= 0) { stopped = true; break; }                 // hpp:907-910 (== :901-902 when all >= 0)
                idcg += val / P.log2tab[ix];
            } else {
                if (val < 0) { stopped = true; break; }                  // hpp:938
                idcg += val / P.log2tab[ix];
                ndcg_u[ix] = (T)((double)ndcg_u[ix] / idcg);             // hpp:927, :940
            }
        }
        if (!P.cumulative) {
            *ndcg_u = (T)(dcg / idcg);                                   // hpp:903, :912
        } else {
            if (stopped) for (; ix < L; ix++) ndcg_u[ix] = (T)((double)ndcg_u[ix] / idcg);   // hpp:946-948
            if (npos < K) {                                              // hpp:951-956 frozen tail (quirk Q4)
                const int upto = K < cand ? K : cand;
                for (int c = npos; c < upto; c++) ndcg_u[c] = ndcg_u[npos - 1];
            }
        }
    }
}

// break_ties_with_noise on the FMA path without rank counting: the reference's noise branch turns a user whose candidates
// all score the same into a NaN row (hpp:524-527).  Necessary: the K best (clean) scores are equal -- only for those users
// (rare) one block re-scores every candidate with the tile kernel's fma chain and raises flag 4 when none differs.
template <typename T>
__global__ void all_equal_check_kernel(const T* __restrict__ cand_score, const int* __restrict__ cand_count, const int C,
                                       const int user0, const int K, const int n,
                                       const int* __restrict__ trp, const int* __restrict__ tri, const int* __restrict__ ustatus,
                                       const T* __restrict__ At, const T* __restrict__ Bt, const T* __restrict__ bias, const int p_pad,
                                       int* __restrict__ uflags)
{
    const int ul = blockIdx.x, u = user0 + ul;
    if (ustatus[u] != 0) return;
    const int t0 = trp[u], t1 = trp[u + 1];
    const int cand = n - (t1 - t0), walk = K < cand ? K : cand;
    if (walk <= 0 || cand_count[ul] < walk) return;
    const T ref = cand_score[(size_t)ul * C];
    if (!(ref == cand_score[(size_t)ul * C + walk - 1])) return;
    constexpr int BN = NumTraits<T>::BN;
    const T* a = At + (size_t)(ul / BM) * p_pad * BM + (ul % BM);
    int differs = 0;
    for (int item = threadIdx.x; item < n && !differs; item += blockDim.x) {
        int lo = t0, hi = t1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (tri[mid] < item) lo = mid + 1; else hi = mid; }
        if (lo < t1 && tri[lo] == item) continue;
        const T* b = Bt + (size_t)(item / BN) * p_pad * BN + (item % BN);
        T acc = (T)0;
        for (int k = 0; k < p_pad; k++) acc = NumTraits<T>::fma(a[(size_t)k * BM], b[(size_t)k * BN], acc);
        if (bias != nullptr) acc += bias[item];
        if (!(acc == ref)) differs = 1;
    }
    if (!__syncthreads_or(differs) && threadIdx.x == 0) atomicOr(&uflags[u], 4);
}

// ---------------------------------------------------------------- per-metric means over users (extension)
// numpy.nanmean of the per-user rows, on the device and in a fixed summation order:
//   metric_partial_kernel  one block per 256 users of a batch and metric (blockIdx.y): sum and count of the non-NaN
//                          entries of every column (W = K if cumulative else 1), users in ascending order
//   metric_final_kernel    one thread per (metric, column): adds the blocks' partials in ascending order, divides
constexpr int MEAN_CHUNK = 256;

template <typename T>
struct MeanParams {
    const T* out[10];        // metric rows of the batch (nullptr: not requested), row stride W (roc/pr: 1)
    int nb, W;
    double* part_sum;        // [blocks_total][10][W]
    int* part_cnt;
    int block0;              // first partial block of this batch
};

template <typename T>
__global__ void metric_partial_kernel(const __grid_constant__ MeanParams<T> P)
{
    const int q = blockIdx.y;
    const T* src = P.out[q];
    const int W = P.W, Wq = q < 8 ? W : 1;
    const int u0 = blockIdx.x * MEAN_CHUNK;
    const int u1 = min(u0 + MEAN_CHUNK, P.nb);
    double* ps = P.part_sum + ((size_t)(P.block0 + blockIdx.x) * 10 + q) * W;
    int* pc = P.part_cnt + ((size_t)(P.block0 + blockIdx.x) * 10 + q) * W;
    if (Wq > 1) {
        // a thread per column walks the chunk's users in order (consecutive threads read consecutive columns)
        for (int c = threadIdx.x; c < W; c += blockDim.x) {
            double s = 0.;
            int cnt = 0;
            if (src)
                for (int u = u0; u < u1; u++) {
                    const double v = (double)src[(size_t)u * Wq + c];
                    if (v == v) { s += v; cnt++; }
                }
            ps[c] = s;
            pc[c] = cnt;
        }
    } else {
        // one value per user: fixed-shape tree over the chunk
        __shared__ double sh_s[MEAN_CHUNK];
        __shared__ int sh_c[MEAN_CHUNK];
        const int u = u0 + threadIdx.x;
        double v = (src && u < u1) ? (double)src[u] : CUDART_NAN;
        sh_s[threadIdx.x] = (v == v) ? v : 0.;
        sh_c[threadIdx.x] = (v == v) ? 1 : 0;
        __syncthreads();
        for (int o = MEAN_CHUNK / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) { sh_s[threadIdx.x] += sh_s[threadIdx.x + o]; sh_c[threadIdx.x] += sh_c[threadIdx.x + o]; }
            __syncthreads();
        }
        if (threadIdx.x == 0) { ps[0] = sh_s[0]; pc[0] = sh_c[0]; }
        for (int c = 1 + threadIdx.x; c < W; c += blockDim.x) { ps[c] = 0.; pc[c] = 0; }
    }
}

__global__ void metric_final_kernel(const double* __restrict__ part_sum, const int* __restrict__ part_cnt, const int blocks,
                                    const int W, double* __restrict__ means, long long* __restrict__ counts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;     // (metric, column)
    if (i >= 10 * W) return;
    double s = 0.;
    long long c = 0;
    for (int b = 0; b < blocks; b++) { s += part_sum[(size_t)b * 10 * W + i]; c += part_cnt[(size_t)b * 10 * W + i]; }
    if (means) means[i] = c > 0 ? s / (double)c : CUDART_NAN;
    if (counts) counts[i] = c;
}

}  // namespace rmb
