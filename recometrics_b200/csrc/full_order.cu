// full_order.cu -- the full-order path: score EVERY candidate of a user, add the reference's tie-breaking noise when asked,
// sort the whole candidate list, read the ranked top-K and the held-out items' ranks off the sorted list.
//
// This is the reference's own procedure (/root/reference/src/recometrics.hpp:499-563: dot1 over the candidate list :499-512,
// noise + validity :514-535, std::sort :552-554) carried out for a chunk of users at a time in HBM.  The two selection
// paths (score_select.cuh, filter_select.cuh) avoid materialising scores and are what the benchmarks run; they keep a
// bounded candidate buffer (k_metrics <= 384) and count ranks on noise-free scores.  This path has neither limit:
//   * k_metrics up to n -- the reference's only bound (hpp:391, recometrics/__init__.py:531-532);
//   * break_ties_with_noise with ROC/PR-AUC or on the all-FMA shapes: candidate number ix (ascending item id) gets draw ix
//     of mt19937(seed + user) (tie_noise.cuh) BEFORE sorting, so top-K AND ranks are those of the noisy scores.
// Scores are the sequential fma chain over k = 0..p_pad-1 (+ bias) of the other kernels: bit-identical values.
// Ties (equal scores after the noise, if any) rank by ascending item id: the radix sort is stable and starts from item order.
//
// HBM-bound by construction (n scores + n keys + n ids per user, four radix passes over them): a fallback, not a benchmark
// path.  Chunks are sized so that scores, keys and values of a chunk fit a few GB.
#include "full_order.h"

#include <cub/device/device_segmented_radix_sort.cuh>

#include "score_select.cuh"
#include "tie_noise.cuh"

namespace rmb {

namespace {

template <typename T> struct SortKey;
template <> struct SortKey<float> { typedef unsigned type; };
template <> struct SortKey<double> { typedef unsigned long long type; };

constexpr int FO_ROWS = 16;              // users per block of score_rows_kernel
constexpr long long FO_MAX_ELEMS = 1ll << 27;   // scores per chunk (also keeps CUB's 32-bit item counts safe)

// ascending sort of these keys = descending order of the scores; NaN of either sign first (the row is a NaN row then)
template <typename T>
__device__ __forceinline__ typename SortKey<T>::type sort_key(const T s)
{
    typedef typename SortKey<T>::type key_t;
    if (s != s) return (key_t)0;
    return (key_t)~(key_t)NumTraits<T>::orderable(s);
}
template <typename T>
__device__ __forceinline__ T score_of_key(const typename SortKey<T>::type k)
{
    typedef typename SortKey<T>::type key_t;
    if (k == (key_t)0) return NumTraits<T>::nan();
    return NumTraits<T>::from_orderable((u64)(key_t)~k);
}

// scores[r][item] for rows r0.. of the chunk: block = BN threads (one item each) x FO_ROWS users
template <typename T>
__global__ void score_rows_kernel(const T* __restrict__ At, const T* __restrict__ Bt, const T* __restrict__ bias,
                                  const int p_pad, const int n, const int row0, const int rows, T* __restrict__ scores)
{
    constexpr int BN = NumTraits<T>::BN;
    extern __shared__ __align__(16) unsigned char fo_smem[];
    T* As = reinterpret_cast<T*>(fo_smem);                         // [p_pad][FO_ROWS]
    const int tile = blockIdx.x, r0 = blockIdx.y * FO_ROWS, tid = threadIdx.x;
    for (int i = tid; i < p_pad * FO_ROWS; i += BN) {
        const int k = i / FO_ROWS, r = i % FO_ROWS;
        const int ul = row0 + r0 + r;                               // row of At
        As[i] = (r0 + r < rows) ? At[(size_t)(ul / BM) * p_pad * BM + (size_t)k * BM + (ul % BM)] : (T)0;
    }
    __syncthreads();
    T acc[FO_ROWS];
#pragma unroll
    for (int r = 0; r < FO_ROWS; r++) acc[r] = (T)0;
    const T* b = Bt + (size_t)tile * p_pad * BN + tid;
    for (int k = 0; k < p_pad; k++) {
        const T bv = b[(size_t)k * BN];
#pragma unroll
        for (int r = 0; r < FO_ROWS; r++) acc[r] = NumTraits<T>::fma(As[k * FO_ROWS + r], bv, acc[r]);
    }
    const int item = tile * BN + tid;
    if (item >= n) return;
    const T bi = bias != nullptr ? bias[item] : (T)0;
#pragma unroll
    for (int r = 0; r < FO_ROWS; r++)
        if (r0 + r < rows) scores[(size_t)(r0 + r) * n + item] = bias != nullptr ? acc[r] + bi : acc[r];
}

// One warp per user of the chunk, noise branch only (hpp:514-535): validity scan over the candidates (any NaN, all equal,
// an infinite extreme => NaN row), then draw ix of mt19937(seed + user) is added to candidate ix (ascending item id).
template <typename T>
__global__ void noise_rows_kernel(T* __restrict__ scores, const int n, const int row0, const int rows, const int user0,
                                  const int* __restrict__ umap, const int* __restrict__ trp, const int* __restrict__ tri,
                                  const int* __restrict__ ustatus, int* __restrict__ uflags, const unsigned long long seed_user0)
{
    extern __shared__ __align__(16) unsigned char fo_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * (blockDim.x >> 5) + warp;
    if (r >= rows) return;
    const int ul = umap ? umap[row0 + r] : row0 + r;
    const int u = user0 + ul;
    if (ustatus[u] != 0) return;
    T* s = scores + (size_t)r * n;
    const int t0 = trp[u], t1 = trp[u + 1];
    // ---- validity (hpp:517-528) ----
    {
        bool has_nan = false;
        T mx = -NumTraits<T>::inf(), mn = NumTraits<T>::inf();
        int t = t0;
        for (int j0 = 0; j0 < n; j0 += 32) {
            unsigned mask = 0;
            while (t < t1) { const int it = tri[t]; if (it >= j0 + 32) break; mask |= 1u << (it - j0); t++; }
            const int j = j0 + lane;
            if (j < n && !((mask >> lane) & 1u)) {
                const T v = s[j];
                has_nan |= (v != v);
                mx = v > mx ? v : mx;
                mn = v < mn ? v : mn;
            }
        }
        has_nan = __any_sync(FULL, has_nan);
        for (int o = 16; o > 0; o >>= 1) {
            const T a = __shfl_xor_sync(FULL, mx, o), b = __shfl_xor_sync(FULL, mn, o);
            mx = a > mx ? a : mx;
            mn = b < mn ? b : mn;
        }
        int f = 0;
        if (has_nan || isinf(mx) || isinf(mn)) f |= 1;
        if (mx == mn) f |= 4;
        if (f) { if (lane == 0) atomicOr(&uflags[u], f); return; }
    }
    // ---- noise (hpp:531-534) ----
    unsigned* x = reinterpret_cast<unsigned*>(fo_smem) + warp * MT_N;
    mt_seed_warp(x, seed_user0 + (unsigned long long)ul, lane);
    mt_twist_warp(x, lane);
    long long blk_lo = 0;                       // first word of the block in x
    constexpr int W = TieNoise<T>::WORDS;
    int t = t0, cbase = 0;                       // candidates before item j0
    for (int j0 = 0; j0 < n; j0 += 32) {
        unsigned mask = 0;
        while (t < t1) { const int it = tri[t]; if (it >= j0 + 32) break; mask |= 1u << (it - j0); t++; }
        const int j = j0 + lane;
        if (j0 + 32 > n) mask |= ~0u << (n - j0);                   // past the catalogue
        const bool is_cand = !((mask >> lane) & 1u);
        const long long w = (long long)(cbase + __popc(~mask & ((1u << lane) - 1u))) * W;      // first word of this lane's draw
        bool pending = is_cand;
        while (__any_sync(FULL, pending)) {
            if (pending && w < blk_lo + MT_N) {
                const unsigned w0 = mt_temper(x[w - blk_lo]);
                const unsigned w1 = W == 2 ? mt_temper(x[w - blk_lo + 1]) : 0u;
                s[j] += TieNoise<T>::draw(w0, w1);
                pending = false;
            }
            if (__any_sync(FULL, pending)) { __syncwarp(); mt_twist_warp(x, lane); blk_lo += MT_N; }
        }
        cbase += __popc(~mask);
    }
}

template <typename T>
__global__ void keys_kernel(const T* __restrict__ scores, const long long total, const int n,
                            typename SortKey<T>::type* __restrict__ keys, int* __restrict__ vals)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    keys[i] = sort_key<T>(scores[i]);
    vals[i] = (int)(i % n);
}

// train items are no candidates: to the end of the order.  One warp per user of the chunk.
template <typename T>
__global__ void mask_train_kernel(typename SortKey<T>::type* __restrict__ keys, const int n, const int row0, const int rows,
                                  const int user0, const int* __restrict__ umap, const int* __restrict__ trp, const int* __restrict__ tri)
{
    typedef typename SortKey<T>::type key_t;
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int u = user0 + (umap ? umap[row0 + r] : row0 + r);
    for (int t = trp[u] + lane; t < trp[u + 1]; t += 32) {
        const int it = tri[t];
        if (it >= 0 && it < n) keys[(size_t)r * n + it] = ~(key_t)0;
    }
}

__global__ void offsets_kernel(int* __restrict__ off, const int rows, const int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= rows) off[i] = i * n;
}

// One warp per user: the first min(K, cand) entries of the sorted list are the ranked top-K; the smallest candidate score;
// and (rank outputs) the position of every held-out item in the sorted list.
template <typename T>
__global__ void gather_order_kernel(const typename SortKey<T>::type* __restrict__ keys, const int* __restrict__ vals, const int n,
                                    const int row0, const int rows, const FullOrderArgs<T> a)
{
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int ul = a.umap ? a.umap[row0 + r] : row0 + r;
    const int u = a.user0 + ul;
    if (a.ustatus[u] != 0) { if (lane == 0) a.cand_count[ul] = 0; return; }
    const int cand = n - (a.trp[u + 1] - a.trp[u]);
    const int walk = a.K < cand ? a.K : cand;
    const typename SortKey<T>::type* k = keys + (size_t)r * n;
    const int* v = vals + (size_t)r * n;
    T* cs = a.cand_score + (size_t)ul * a.C;
    int* ci = a.cand_item + (size_t)ul * a.C;
    for (int i = lane; i < walk; i += 32) { cs[i] = score_of_key<T>(k[i]); ci[i] = v[i]; }
    if (lane == 0) {
        a.cand_count[ul] = walk;
        if (a.umin != nullptr && cand > 0) a.umin[u] = NumTraits<T>::orderable(score_of_key<T>(k[cand - 1]));
        if (cand > 0 && k[0] == 0) atomicOr(&a.uflags[u], 1);             // a NaN among the candidates
    }
    if (a.auc_cnt == nullptr) return;
    const int tp0 = a.tep[u], npos = a.tep[u + 1] - tp0;
    const int* ti = a.tei + tp0;
    int found = 0;
    for (int base = 0; base < cand && found < npos; base += 32) {
        const int i = base + lane;
        int e = -1;
        if (i < cand) {
            const int item = v[i];
            int lo = 0, hi = npos;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (ti[mid] < item) lo = mid + 1; else hi = mid; }
            if (lo < npos && ti[lo] == item) e = lo;
        }
        const unsigned m = __ballot_sync(FULL, e >= 0);
        if (e >= 0) {
            const int h = found + __popc(m & (FULL >> (31 - lane)));         // this is the h-th best held-out item
            a.auc_cnt[(size_t)tp0 + npos - h] = (unsigned)i;                  // candidates ranked before it
            a.pos_perm[(size_t)tp0 + npos - h] = e;
        }
        found += __popc(m);
    }
}

}  // namespace

void full_order_plan(const int n, const int elem_bytes, const int max_users, int* chunk_users, size_t* scratch_bytes)
{
    long long uc = FO_MAX_ELEMS / (n > 0 ? n : 1);
    if (uc < 1) uc = 1;
    if (uc > max_users) uc = max_users;
    uc = (uc + FO_ROWS - 1) / FO_ROWS * FO_ROWS;
    if (uc * (long long)n > 0x7fffffffLL) uc = 0x7fffffffLL / (n > 0 ? n : 1);      // (offsets and CUB's item counts are 32-bit)
    if (uc < 1) uc = 1;
    const size_t elems = (size_t)uc * (size_t)n;
    const size_t key_bytes = elem_bytes == 4 ? 4 : 8;
    size_t temp = 0;
    if (elem_bytes == 4)
        cub::DeviceSegmentedRadixSort::SortPairs(nullptr, temp, (const unsigned*)nullptr, (unsigned*)nullptr, (const int*)nullptr, (int*)nullptr,
                                                 (int)elems, (int)uc, (const int*)nullptr, (const int*)nullptr);
    else
        cub::DeviceSegmentedRadixSort::SortPairs(nullptr, temp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                                 (const int*)nullptr, (int*)nullptr, (int)elems, (int)uc, (const int*)nullptr, (const int*)nullptr);
    auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
    *chunk_users = (int)uc;
    *scratch_bytes = up(elems * elem_bytes) + 2 * up(elems * key_bytes) + 2 * up(elems * 4) + up((size_t)(uc + 1) * 4) + up(temp) + 256;
}

template <typename T>
cudaError_t full_order_run(const FullOrderArgs<T>& a, cudaStream_t st, long long* launches)
{
    typedef typename SortKey<T>::type key_t;
    constexpr int BN = NumTraits<T>::BN;
    const int n = a.n, uc = a.chunk_users;
    const size_t elems = (size_t)uc * (size_t)n;
    auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
    unsigned char* base = static_cast<unsigned char*>(a.scratch);
    T* scores = reinterpret_cast<T*>(base); base += up(elems * sizeof(T));
    key_t* keys0 = reinterpret_cast<key_t*>(base); base += up(elems * sizeof(key_t));
    key_t* keys1 = reinterpret_cast<key_t*>(base); base += up(elems * sizeof(key_t));
    int* vals0 = reinterpret_cast<int*>(base); base += up(elems * 4);
    int* vals1 = reinterpret_cast<int*>(base); base += up(elems * 4);
    int* offs = reinterpret_cast<int*>(base); base += up((size_t)(uc + 1) * 4);
    void* temp = base;
    size_t temp_bytes = a.scratch_bytes - (size_t)(base - static_cast<unsigned char*>(a.scratch));

    const int NT = (n + BN - 1) / BN;
    const size_t smem_a = (size_t)a.p_pad * FO_ROWS * sizeof(T);
    cudaError_t e = cudaFuncSetAttribute(score_rows_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
    if (e != cudaSuccess) return e;
    offsets_kernel<<<(uc + 1 + 255) / 256, 256, 0, st>>>(offs, uc, n);
    (*launches)++;
    for (int row0 = 0; row0 < a.nb; row0 += uc) {
        const int rows = (a.nb - row0) < uc ? (a.nb - row0) : uc;
        const long long total = (long long)rows * n;
        score_rows_kernel<T><<<dim3(NT, (rows + FO_ROWS - 1) / FO_ROWS), BN, smem_a, st>>>(a.At, a.Bt, a.bias, a.p_pad, n, row0, rows, scores);
        (*launches)++;
        if (a.noise) {
            noise_rows_kernel<T><<<(rows + 3) / 4, 128, 4 * MT_N * sizeof(unsigned), st>>>(scores, n, row0, rows, a.user0, a.umap, a.trp, a.tri,
                                                                                            a.ustatus, a.uflags, a.seed_user0);
            (*launches)++;
        }
        keys_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(scores, total, n, keys0, vals0);
        mask_train_kernel<T><<<(rows + 7) / 8, 256, 0, st>>>(keys0, n, row0, rows, a.user0, a.umap, a.trp, a.tri);
        (*launches) += 2;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        e = cub::DeviceSegmentedRadixSort::SortPairs(temp, temp_bytes, (const key_t*)keys0, keys1, (const int*)vals0, vals1, (int)total, rows,
                                                     (const int*)offs, (const int*)offs + 1, 0, (int)sizeof(key_t) * 8, st);
        if (e != cudaSuccess) return e;
        (*launches) += 1;
        gather_order_kernel<T><<<(rows + 7) / 8, 256, 0, st>>>(keys1, vals1, n, row0, rows, a);
        (*launches)++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template cudaError_t full_order_run<float>(const FullOrderArgs<float>&, cudaStream_t, long long*);
template cudaError_t full_order_run<double>(const FullOrderArgs<double>&, cudaStream_t, long long*);

}  // namespace rmb
