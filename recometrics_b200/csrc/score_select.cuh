// score_select.cuh -- fused "score -> exclude train -> running top-K (+ rank counting)" kernel.
//
// Replaces, for a tile of 128 users at a time, the per-user loop body of the reference:
//   candidate list      /root/reference/src/recometrics.hpp:491-497
//   dot1 scoring        /root/reference/src/recometrics.hpp:84-112, :499-512
//   partial_sort / sort /root/reference/src/recometrics.hpp:537-563
//   (the AUC walks of :795-865 become a counting pass over the same score tiles)
//
// Design (sm_100a; FP32 on the packed FFMA2 pipe, FP64 on DFMA):
//   * Factors are re-tiled once per call into slabs  At[user tile][k][128]  /  Bt[item tile][k][128]
//     (k-major inside a 128-wide tile), so the operand chunk of one pipeline stage is ONE contiguous
//     block: a single TMA bulk copy (cp.async.bulk, SASS UBLKCP) lands it in shared memory in
//     exactly the layout the FMA micro-kernel reads.
//   * CTA = 8 compute warps + 1 producer warp.  The producer's elected lane runs the TMA ring
//     (full/empty mbarriers, STAGES deep) over the flattened (item tile, k chunk) space; compute
//     warps never meet at a CTA barrier.
//   * CTA tile 128 users x 128 items; a compute warp OWNS 16 users x 128 items (thread micro-tile
//     8 users x 8 items in registers), so all per-user selection state is warp-private.
//     fp32: the 8x8 micro-tile is 32 packed accumulators updated with fma.rn.f32x2, user factor as
//     the scalar-broadcast operand (SASS: FFMA2 Rd, Ra.F32, Rb.F32x2, Rc.F32x2) -- half the issue
//     slots of FFMA and no register-bank conflicts (profiles/r01_ubench_fma.txt).
//   * The score tile never leaves registers: per user row a thread takes the max of its 8 scores
//     and compares it with the user's running K-th best (tau).  Survivors (~K ln(n/K) per user over
//     the whole catalogue) are checked against the user's train row (only in item tiles the sorted
//     train row intersects -- a per-row cursor keeps the next train item id) and appended to a
//     per-user candidate buffer of C entries in global memory (L2 resident).  When a buffer passes
//     C - BN entries its warp cuts it back to the best K with a bitwise radix select on the
//     order-preserving integer image of the scores (warp REDUX counts; ties at the cut resolved by
//     item id), which refreshes tau.  rank_topk_kernel finally orders the K survivors.
//   * AUC mode (ROC/PR requested): the warp stages its 16 x BN score block in shared memory, masks
//     train items there, and every lane counts, for the held-out items it owns, the candidates
//     scoring strictly higher (compare + add, no atomics).  Scores of held-out items are
//     pre-computed with the same FMA order (prep.cuh) so comparisons are exact.
//   * Everything outside the FMA loop is written as rolled loops / out-of-line calls: the epilogue
//     runs once per item tile and must not evict the FMA loop from the instruction cache.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <limits.h>
#include <string.h>

// build-time tuning knobs of the FMA loop (tools/variants.sh times the alternatives)
#ifndef RMB_SWPIPE
#define RMB_SWPIPE 0        // 1: explicit register double buffering of the operand fragments
#endif
#ifndef RMB_UNROLL
#define RMB_UNROLL 8        // factors per unrolled block of the FMA loop (multiple of KPAD)
#endif
#ifndef RMB_EARLYTRY
#define RMB_EARLYTRY 0      // 1: probe the next stage's mbarrier one block before it is needed
#endif

namespace rmb {

constexpr int BM = 128;             // users per CTA tile (items per tile: NumTraits<T>::BN)
constexpr int NCWARPS = 8;          // compute warps (16 user rows each)
constexpr int NTHREADS = (NCWARPS + 1) * 32;   // + 1 producer warp
constexpr int KPAD = 8;             // factors are zero-padded to a multiple of this
constexpr unsigned FULL = 0xffffffffu;
typedef unsigned long long u64;

template <typename T> struct NumTraits;
template <> struct NumTraits<float> {
    static constexpr int BN = 128;      // items per tile (thread micro-tile 8 users x 8 items)
    static constexpr int BK = 32;       // factors per pipeline stage
    static constexpr int STAGES = 4;
    typedef unsigned key_t;             // order-preserving integer image of a score
    static constexpr int KEYBITS = 32;
    __device__ __forceinline__ static key_t key(float x) { return (key_t)orderable(x); }
    __device__ __forceinline__ static float inf() { return CUDART_INF_F; }
    __device__ __forceinline__ static float nan() { return CUDART_NAN_F; }
    __device__ __forceinline__ static float fma(float a, float b, float c) { return fmaf(a, b, c); }
    __device__ __forceinline__ static u64 orderable(float x) {
        unsigned u = __float_as_uint(x);
        u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;
        return (u64)u;
    }
    __host__ __device__ static float from_orderable(u64 o) {
        unsigned u = (unsigned)o;
        u ^= (u >> 31) ? 0x80000000u : 0xffffffffu;
#ifdef __CUDA_ARCH__
        return __uint_as_float(u);
#else
        float f; memcpy(&f, &u, 4); return f;
#endif
    }
};
template <> struct NumTraits<double> {
    static constexpr int BN = 64;       // items per tile (thread micro-tile 8 users x 4 items: the
                                        // register file of a 9-warp CTA holds 168 registers per thread)
    static constexpr int BK = 16;
    static constexpr int STAGES = 4;
    typedef u64 key_t;
    static constexpr int KEYBITS = 64;
    __device__ __forceinline__ static key_t key(double x) { return orderable(x); }
    __device__ __forceinline__ static double inf() { return CUDART_INF; }
    __device__ __forceinline__ static double nan() { return CUDART_NAN; }
    __device__ __forceinline__ static double fma(double a, double b, double c) { return ::fma(a, b, c); }
    __device__ __forceinline__ static u64 orderable(double x) {
        u64 u = (u64)__double_as_longlong(x);
        u ^= (u >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull;
        return u;
    }
    __host__ __device__ static double from_orderable(u64 u) {
        u ^= (u >> 63) ? 0x8000000000000000ull : 0xffffffffffffffffull;
#ifdef __CUDA_ARCH__
        return __longlong_as_double((long long)u);
#else
        double f; memcpy(&f, &u, 8); return f;
#endif
    }
};

// ------------------------------------------------------------------ ranking helpers
// total order used for ranking: score descending, ties by ascending item id
// (the reference's comparator is a strict '>' on the score, hpp:538-540 / :552-554; its tie order
//  is whatever libstdc++ does -- SURVEY quirk Q8 -- so a deterministic refinement is chosen here)
template <typename T>
__device__ __forceinline__ bool ranks_before(T sa, int ia, T sb, int ib)
{
    return (sa > sb) || (sa == sb && ia < ib);
}

// Bitonic sorting network over 32*E (score,item) pairs held E per lane (element i = e*32 + lane),
// best first.  Fully unrolled; strides >= 32 are register exchanges, strides < 32 are shuffles.
template <typename T, int E>
__device__ __forceinline__ void warp_sort_ranked(T (&s)[E], int (&it)[E], const int lane)
{
    constexpr int N = 32 * E;
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int je = j >> 5;
#pragma unroll
                for (int e = 0; e < E; e++) {
                    if ((e & je) == 0) {
                        const int e2 = e | je;
                        const bool desc = (((e * 32) & k) == 0);
                        const bool first_better = ranks_before<T>(s[e], it[e], s[e2], it[e2]);
                        const bool sw = desc ? !first_better : first_better;
                        if (sw) {
                            const T ts = s[e]; s[e] = s[e2]; s[e2] = ts;
                            const int ti = it[e]; it[e] = it[e2]; it[e2] = ti;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const int i = e * 32 + lane;
                    const bool desc = ((i & k) == 0);
                    const bool lower = ((lane & j) == 0);
                    const T so = __shfl_xor_sync(FULL, s[e], j);
                    const int io = __shfl_xor_sync(FULL, it[e], j);
                    const bool mine_better = ranks_before<T>(s[e], it[e], so, io);
                    const bool keep_better = (lower == desc);
                    if (mine_better != keep_better) { s[e] = so; it[e] = io; }
                }
            }
        }
    }
}

template <typename T>
struct ScoreSelectParams {
    const T* __restrict__ At;      // [user tiles][p_pad][128]  user factors of this batch, zero padded
    const T* __restrict__ Bt;      // [item tiles][p_pad][128]  item factors, zero padded
    const T* __restrict__ bias;    // [item tiles * 128] item biases (zero padded) or nullptr
    int p_pad;
    int n;                          // items
    int mb;                         // users in this batch
    int user0;                      // row (of the CSR / status arrays) of the batch's first user
    const int* __restrict__ trp;    // train CSR
    const int* __restrict__ tri;
    const int* __restrict__ tep;    // test CSR indptr
    const int* __restrict__ ustatus;// [m] 0 = user is ranked, !=0 = NaN row decided before scoring
    T* cand_score;                  // [mb_pad][C]
    int* cand_item;                 // [mb_pad][C]
    int* cand_count;                // [mb_pad] entries left at the head of the buffer (<= K, unordered)
    int* uflags;                    // [m] bit0: a candidate score was NaN
    int K;
    // rank counting (AUC mode)
    const T* __restrict__ pos_sorted;   // [nnz_test] held-out item scores, ascending per user
    unsigned int* auc_cnt;              // [nnz_test] number of candidates scoring strictly above
                                        // the j-th smallest held-out score of the row (zeroed)
    u64* umin;                          // [m] orderable(min candidate score), init ~0
    const int* __restrict__ umap;       // optional [mb]: row r of this launch is batch-local user umap[r] (CSR rows, status, flags and
                                        // candidate buffers are those of the mapped user; At holds the launch's rows in order).  Used
                                        // to re-run only the users the tensor-core filter handed back (api.cu)
};

// ------------------------------------------------------------------ PTX wrappers (TMA ring)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(const unsigned bar, const unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(const unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(const unsigned bar, const unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(const unsigned bar, const unsigned parity)
{
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// non-blocking probe (the result arrives ~90 cycles later: issue it early, test it late)
__device__ __forceinline__ unsigned mbar_try(const unsigned bar, const unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
// global -> shared bulk copy executed by the TMA unit, completion counted on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(const unsigned dst, const void* src, const unsigned bytes, const unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ u64 pack2(const float lo, const float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(const u64 v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// d = a * b + d on two packed fp32 lanes (each lane an IEEE fma, same rounding as fmaf)
__device__ __forceinline__ void fma2(u64& d, const u64 a, const u64 b)
{
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ float max_nan(const float a, const float b)
{
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ double max_nan(const double a, const double b)
{
    return (a != a || b != b) ? CUDART_NAN : (a > b ? a : b);
}

__device__ __forceinline__ void lds_vec(const float* p, float (&v)[4])
{
    const float4 f = *reinterpret_cast<const float4*>(p);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
}
__device__ __forceinline__ void lds_vec(const double* p, double (&v)[2])
{
    const double2 f = *reinterpret_cast<const double2*>(p);
    v[0] = f.x; v[1] = f.y;
}
__device__ __forceinline__ void sts4(float* p, const float* v)
{
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void sts4(double* p, const double* v)
{
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}

// ------------------------------------------------------------------ register micro-tiles
// Thread (ly = lane >> 4, lx = lane & 15) of a compute warp holds, of the warp's 16 x BN block,
//   rows  ly*4 + (i & 3) + (i >> 2) * 8      i = 0..7
//   cols  lx*4 + (c & 3) + (c >> 2) * 64     c = 0..NC-1   (NC = BN / 16: 8 for fp32, 4 for fp64)
template <typename T> struct MicroTile;

template <> struct MicroTile<float> {
    static constexpr int NC = 8;
    u64 acc[8][4];   // acc[i][c/2] = scores (i, c), (i, c+1)
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc[i][q] = 0ull;
    }
    struct Frag { float a[8]; u64 b[4]; };
    // sA: the warp's 16 user values of factor k; sB: the tile's 128 item values of factor k
    __device__ __forceinline__ static void load(Frag& f, const float* __restrict__ sA, const float* __restrict__ sB, const int ly, const int lx)
    {
        const float4 a0 = *reinterpret_cast<const float4*>(sA + ly * 4);
        const float4 a1 = *reinterpret_cast<const float4*>(sA + 8 + ly * 4);
        const ulonglong2 b0 = *reinterpret_cast<const ulonglong2*>(sB + lx * 4);
        const ulonglong2 b1 = *reinterpret_cast<const ulonglong2*>(sB + 64 + lx * 4);
        f.a[0] = a0.x; f.a[1] = a0.y; f.a[2] = a0.z; f.a[3] = a0.w;
        f.a[4] = a1.x; f.a[5] = a1.y; f.a[6] = a1.z; f.a[7] = a1.w;
        f.b[0] = b0.x; f.b[1] = b0.y; f.b[2] = b1.x; f.b[3] = b1.y;
    }
    __device__ __forceinline__ void compute(const Frag& f)
    {
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int i = 0; i < 8; i++) fma2(acc[i][q], pack2(f.a[i], f.a[i]), f.b[q]);
    }
    __device__ __forceinline__ void row(const int i, float (&s)[8]) const
    {
#pragma unroll
        for (int q = 0; q < 4; q++) unpack2(acc[i][q], s[2 * q], s[2 * q + 1]);
    }
};

template <> struct MicroTile<double> {
    static constexpr int NC = 4;
    double acc[8][4];
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[i][c] = 0.;
    }
    struct Frag { double a[8]; double b[4]; };
    __device__ __forceinline__ static void load(Frag& f, const double* __restrict__ sA, const double* __restrict__ sB, const int ly, const int lx)
    {
#pragma unroll
        for (int g = 0; g < 2; g++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const double2 va = *reinterpret_cast<const double2*>(sA + g * 8 + ly * 4 + h * 2);
                f.a[g * 4 + h * 2] = va.x; f.a[g * 4 + h * 2 + 1] = va.y;
            }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const double2 vb = *reinterpret_cast<const double2*>(sB + lx * 4 + h * 2);
            f.b[h * 2] = vb.x; f.b[h * 2 + 1] = vb.y;
        }
    }
    __device__ __forceinline__ void compute(const Frag& f)
    {
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i][c] = ::fma(f.a[i], f.b[c], acc[i][c]);
    }
    __device__ __forceinline__ void row(const int i, double (&s)[4]) const
    {
#pragma unroll
        for (int c = 0; c < 4; c++) s[c] = acc[i][c];
    }
};


// ------------------------------------------------------------------ per-CTA shared state
template <typename T>
struct RowState {
    T tau[BM];          // running K-th best score; +inf = row is not ranked (padding / NaN-row user)
    int cnt[BM];        // entries in the row's candidate buffer
    int nan[BM];        // a candidate score was NaN
    int nxt_train[BM];  // smallest train item id >= first item of the current tile (INT_MAX: none)
    int cur_train[BM];  // its position in the train CSR
    int end_train[BM];
    int tp0[BM];        // first entry of the held-out (test) row
    int npos[BM];       // its length (0 for rows that are not ranked)
    int urow[BM];       // batch-local user of the row (its candidate buffer, status, flags): the row itself unless ScoreSelectParams::umap
};

template <typename T>
struct SmemLayout {
    static constexpr int S = NumTraits<T>::STAGES, BK = NumTraits<T>::BK, BN = NumTraits<T>::BN;
    static constexpr size_t a_off = 0;                                            // [S][BK][BM]
    static constexpr size_t b_off = (size_t)S * BK * BM * sizeof(T);              // [S][BK][BN]
    static constexpr size_t rs_off = b_off + (size_t)S * BK * BN * sizeof(T);
    static constexpr size_t bar_off = rs_off + ((sizeof(RowState<T>) + 15) & ~size_t(15));   // full[S], empty[S]
    static constexpr size_t plain_bytes = bar_off + 2 * S * sizeof(u64);
    // AUC mode only
    static constexpr size_t blk_off = plain_bytes;                                // [NCWARPS][16][BN] score blocks
    static constexpr size_t pj_off = blk_off + (size_t)NCWARPS * 16 * BN * sizeof(T);    // [BM][16] thresholds
    static constexpr size_t cj_off = pj_off + (size_t)BM * 16 * sizeof(T);               // [BM][16] counters
    static constexpr size_t auc_bytes = cj_off + (size_t)BM * 16 * sizeof(unsigned);
};

template <typename T, bool AUC>
inline size_t score_select_smem_bytes() { return AUC ? SmemLayout<T>::auc_bytes : SmemLayout<T>::plain_bytes; }

extern __shared__ __align__(128) unsigned char smem_raw[];
template <typename T>
__device__ __forceinline__ RowState<T>* row_state() { return reinterpret_cast<RowState<T>*>(smem_raw + SmemLayout<T>::rs_off); }

// ------------------------------------------------------------------ selection: out-of-line slow paths
// One warp: cut the first nv (>= K) entries of a user's candidate buffer back to the best K
// (score descending, ties at the cut by ascending item id -- the order of ranks_before), left
// unordered at the head; publish the K-th best score (tau) and the new count.
template <typename T, int C>
__device__ __noinline__ void compact_user(T* cs, int* ci, const int nv, const int K, const int lane,
                                          T* tau_out, int* cnt_out)
{
    typedef typename NumTraits<T>::key_t key_t;
    constexpr int E = C / 32;
    if (nv < K) return;                       // nothing to drop yet; tau stays -inf
    key_t key[E];
    int it[E];
#pragma unroll
    for (int e = 0; e < E; e++) {
        const int idx = e * 32 + lane;
        const bool v = idx < nv;
        key[e] = v ? NumTraits<T>::key(cs[idx]) : (key_t)0;     // 0 sorts below every score (NaN is never stored)
        it[e] = v ? ci[idx] : INT_MAX;
    }
    __syncwarp();
    // K-th largest key, bit by bit from the top: the largest t with #(key >= t) >= K
    key_t t = 0;
    for (int b = NumTraits<T>::KEYBITS - 1; b >= 0; b--) {
        const key_t cand = t | ((key_t)1 << b);
        int c = 0;
#pragma unroll
        for (int e = 0; e < E; e++) c += (key[e] >= cand) ? 1 : 0;
        c = __reduce_add_sync(FULL, c);
        if (c >= K) t = cand;
    }
    int cgt = 0, cge = 0;
#pragma unroll
    for (int e = 0; e < E; e++) { cgt += (key[e] > t) ? 1 : 0; cge += (key[e] >= t) ? 1 : 0; }
    cgt = __reduce_add_sync(FULL, cgt);
    cge = __reduce_add_sync(FULL, cge);
    unsigned id_cut = 0x7fffffffu;            // keep tied entries with item id <= id_cut
    if (cge > K) {
        // more entries tie with the K-th score than fit: keep the (K - cgt) smallest item ids among them,
        // i.e. id_cut = max{x : #(tied, id < x) < need}
        const int need = K - cgt;
        unsigned x = 0;
        for (int b = 30; b >= 0; b--) {
            const unsigned cand = x | (1u << b);
            int c = 0;
#pragma unroll
            for (int e = 0; e < E; e++) c += (key[e] == t && (unsigned)it[e] < cand) ? 1 : 0;
            c = __reduce_add_sync(FULL, c);
            if (c < need) x = cand;
        }
        id_cut = x;
    }
    int base = 0;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const bool keep = (key[e] > t) || (key[e] == t && (unsigned)it[e] <= id_cut);
        const unsigned mask = __ballot_sync(FULL, keep);
        if (keep) {
            const int pos = base + __popc(mask & ((1u << lane) - 1u));
            cs[pos] = NumTraits<T>::from_orderable((u64)key[e]);
            ci[pos] = it[e];
        }
        base += __popc(mask);
    }
    if (lane == 0) { *tau_out = NumTraits<T>::from_orderable((u64)t); *cnt_out = base; }
    __syncwarp();
}

// is `item` in the sorted train row segment [lo, hi) ?  (hpp:494-495 moves those out of the pool)
__device__ __forceinline__ bool in_train_segment(const int* __restrict__ tri, int lo, int hi, const int item)
{
    const int end = hi;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tri[mid] < item) lo = mid + 1; else hi = mid;
    }
    return lo < end && tri[lo] == item;
}

// Slow path of the selection filter (taken by ~K ln(n/K) scores per user): candidate checks, then
// append to the user's buffer (cs / ci = the row's candidate buffer).  Everything arrives by value:
// the kernel parameters live in constant memory, which an out-of-line function could only reach
// through a generic pointer.
template <typename T>
__device__ __forceinline__ void insert_candidate(RowState<T>* rs, T* cs, int* ci, const int* __restrict__ tri, const int n,
                                                 const T s, const int row, const int item, const int item_end)
{
    if (item >= n) return;                                      // padding column
    if (rs->nxt_train[row] < item_end &&                        // the train row intersects this tile
        in_train_segment(tri, rs->cur_train[row], rs->end_train[row], item)) return;   // not a candidate
    if (s != s) { rs->nan[row] = 1; return; }                   // NaN candidate score => NaN row
    const int slot = atomicAdd(&rs->cnt[row], 1);
    // slot < C always: the buffer holds <= C-BN entries when a tile starts and a tile adds <= BN
    cs[slot] = s;
    ci[slot] = item;
}

// A thread's scores of one user row in one item tile, at least one of which is not below tau.
// item_lo = id of the first of the four columns.
template <typename T>
__device__ __noinline__ void row_slow(T* cs, int* ci, const int* __restrict__ tri, const int n, const T tau, const int row,
                                      const int item_lo, const int item_end,
                                      const T s0, const T s1, const T s2, const T s3)
{
    RowState<T>* rs = row_state<T>();
    if (tau == NumTraits<T>::inf()) return;                     // row is not ranked (padding / NaN-row user)
    if (!(s0 < tau)) insert_candidate<T>(rs, cs, ci, tri, n, s0, row, item_lo, item_end);
    if (!(s1 < tau)) insert_candidate<T>(rs, cs, ci, tri, n, s1, row, item_lo + 1, item_end);
    if (!(s2 < tau)) insert_candidate<T>(rs, cs, ci, tri, n, s2, row, item_lo + 2, item_end);
    if (!(s3 < tau)) insert_candidate<T>(rs, cs, ci, tri, n, s3, row, item_lo + 3, item_end);
}

// number of a / b / c / d strictly above p, subtracted as 0 / -1 masks (SASS: FSET + IADD3)
__device__ __forceinline__ unsigned gt_mask(const float a, const float p)
{
    unsigned r;
    asm("set.gt.u32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(p));
    return r;
}
__device__ __forceinline__ unsigned gt_mask(const double a, const double p)
{
    unsigned r;
    asm("set.gt.u32.f64 %0, %1, %2;" : "=r"(r) : "d"(a), "d"(p));
    return r;
}

// AUC counting of one staged row (BN scores at src) against up to 4 thresholds per lane.
template <typename T, int NT>
__device__ __forceinline__ void count_above(const T* __restrict__ src, const T (&thr)[NT], unsigned (&cnt)[NT])
{
    constexpr int BN = NumTraits<T>::BN;
    constexpr int VEC = 16 / (int)sizeof(T);
#pragma unroll 4
    for (int x = 0; x < BN; x += VEC) {
        T v[VEC];
        lds_vec(src + x, v);
#pragma unroll
        for (int q = 0; q < NT; q++)
#pragma unroll
            for (int e = 0; e < VEC; e++) cnt[q] -= gt_mask(v[e], thr[q]);
    }
}

template <typename T, int C, bool AUC>
__global__ void __launch_bounds__(NTHREADS, 1)
score_select_kernel(const __grid_constant__ ScoreSelectParams<T> P)
{
    typedef SmemLayout<T> L;
    constexpr int S = NumTraits<T>::STAGES;
    constexpr int BK = NumTraits<T>::BK;
    constexpr int BN = NumTraits<T>::BN;
    constexpr int NC = MicroTile<T>::NC;            // item columns per thread

    T* As = reinterpret_cast<T*>(smem_raw + L::a_off);
    T* Bs = reinterpret_cast<T*>(smem_raw + L::b_off);
    RowState<T>* rs = row_state<T>();
    const unsigned bar_full = smem_u32(smem_raw + L::bar_off), bar_empty = bar_full + 8 * S;
    T* pj_s = reinterpret_cast<T*>(smem_raw + L::pj_off);                 // AUC only
    unsigned* cj_s = reinterpret_cast<unsigned*>(smem_raw + L::cj_off);   // AUC only

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int tile_u0 = blockIdx.x * BM;            // first user (batch-local) of this CTA
    const int KC = (P.p_pad + BK - 1) / BK;
    const int NT = (P.n + BN - 1) / BN;
    const int total = NT * KC;

    // ---- one-time setup: per-row selection state, train cursors, barriers ----
    for (int r = tid; r < BM; r += NTHREADS) {
        const int ul0 = tile_u0 + r;
        const int ul = (ul0 < P.mb && P.umap != nullptr) ? P.umap[ul0] : ul0;
        const bool ranked = (ul0 < P.mb) && (P.ustatus[P.user0 + ul] == 0);
        rs->urow[r] = ul0 < P.mb ? ul : 0;
        rs->tau[r] = ranked ? -NumTraits<T>::inf() : NumTraits<T>::inf();
        rs->cnt[r] = 0;
        rs->nan[r] = 0;
        int cur = 0, end = 0, tp0 = 0, npos = 0;
        if (ranked) {
            const int u = P.user0 + ul;
            cur = P.trp[u]; end = P.trp[u + 1];
            tp0 = P.tep[u]; npos = P.tep[u + 1] - tp0;
        }
        rs->cur_train[r] = cur;
        rs->end_train[r] = end;
        rs->nxt_train[r] = cur < end ? P.tri[cur] : INT_MAX;
        rs->tp0[r] = tp0;
        rs->npos[r] = npos;
        if (AUC) {
            for (int j = 0; j < 16; j++) {
                pj_s[r * 16 + j] = j < npos ? P.pos_sorted[tp0 + j] : NumTraits<T>::inf();
                cj_s[r * 16 + j] = 0;
            }
        }
    }
    if (tid == 0) {
        for (int s = 0; s < S; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, NCWARPS); }
        mbar_fence_init();
    }
    __syncthreads();   // the only CTA-wide barrier

    if (warp == NCWARPS) {
        // ===================== producer: TMA ring over (item tile, k chunk) =====================
        if (lane == 0) {
            const T* gA = P.At + (size_t)blockIdx.x * P.p_pad * BM;
            int tile = 0, kc = 0;
            for (int it = 0; it < total; it++) {
                const int s = it % S;
                if (it >= S) mbar_wait(bar_empty + 8 * s, ((it / S) - 1) & 1);
                const int k0 = kc * BK;
                const int kcount = (P.p_pad - k0) < BK ? (P.p_pad - k0) : BK;
                const unsigned bytes_a = (unsigned)(kcount * BM * sizeof(T)), bytes_b = (unsigned)(kcount * BN * sizeof(T));
                mbar_arrive_expect_tx(bar_full + 8 * s, bytes_a + bytes_b);
                tma_bulk_g2s(smem_u32(As + (size_t)s * BK * BM), gA + (size_t)k0 * BM, bytes_a, bar_full + 8 * s);
                tma_bulk_g2s(smem_u32(Bs + (size_t)s * BK * BN), P.Bt + ((size_t)tile * P.p_pad + k0) * BN, bytes_b,
                             bar_full + 8 * s);
                if (++kc == KC) { kc = 0; tile++; }
            }
        }
        return;
    }

    // ===================== compute warps =====================
    const int ly = lane >> 4, lx = lane & 15;
    const int wrow0 = warp * 16;                    // first CTA row of this warp
    T* blk = reinterpret_cast<T*>(smem_raw + L::blk_off) + (size_t)warp * 16 * BN;   // AUC only: [16][BN]

    T rowmin[AUC ? 8 : 1];                          // AUC: smallest candidate score seen per thread row
    if (AUC) {
#pragma unroll
        for (int i = 0; i < 8; i++) rowmin[i] = NumTraits<T>::inf();
    }

    MicroTile<T> mt;
    typename MicroTile<T>::Frag frag[2];            // operand registers, double buffered across k
    int it = 0;
    unsigned ready = mbar_try(bar_full, 0);         // probe of the stage about to be consumed
    for (int tile = 0; tile < NT; tile++) {
        const int item0 = tile * BN;
        mt.zero();
        for (int kc = 0; kc < KC; kc++, it++) {
            const int s = it % S;
            if (!ready) mbar_wait(bar_full + 8 * s, (it / S) & 1);
            const int k0 = kc * BK;
            const int kcount = (P.p_pad - k0) < BK ? (P.p_pad - k0) : BK;
            const T* sA = As + (size_t)s * BK * BM + wrow0;
            const T* sB = Bs + (size_t)s * BK * BN;
#if RMB_SWPIPE
            MicroTile<T>::load(frag[0], sA, sB, ly, lx);
            for (int kk0 = 0; kk0 < kcount; kk0 += KPAD) {
                if (kk0 + KPAD >= kcount) {          // last block of the stage: probe the next stage now
                    const int nit = it + 1;
                    ready = (nit < total) ? mbar_try(bar_full + 8 * (nit % S), (nit / S) & 1) : 1u;
                }
#pragma unroll
                for (int kk = 0; kk < KPAD; kk++) {
                    if (kk + 1 < KPAD || kk0 + KPAD < kcount)
                        MicroTile<T>::load(frag[(kk + 1) & 1], sA + (kk0 + kk + 1) * BM, sB + (kk0 + kk + 1) * BN, ly, lx);
                    mt.compute(frag[kk & 1]);
                }
            }
#else
            for (int kk0 = 0; kk0 < kcount; kk0 += RMB_UNROLL) {
                if (RMB_EARLYTRY && kk0 + RMB_UNROLL >= kcount) {   // last block of the stage: probe the next stage now
                    const int nit = it + 1;
                    ready = (nit < total) ? mbar_try(bar_full + 8 * (nit % S), (nit / S) & 1) : 1u;
                }
#pragma unroll
                for (int kk = 0; kk < RMB_UNROLL; kk++) {
                    if (RMB_UNROLL > KPAD && kk0 + kk >= kcount) break;     // only the tail chunk of a k that is not a multiple of the unroll
                    MicroTile<T>::load(frag[0], sA + (kk0 + kk) * BM, sB + (kk0 + kk) * BN, ly, lx);
                    mt.compute(frag[0]);
                }
            }
            if (!RMB_EARLYTRY) ready = 0;
#endif
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * s);
        }

        // ---------------- epilogue of item tile `tile` (warp-private) ----------------
        T biasv[NC];
        if (P.bias != nullptr) {
#pragma unroll
            for (int c = 0; c < NC; c++) biasv[c] = P.bias[item0 + lx * 4 + (c & 3) + (c >> 2) * 64];
        }
        const bool tile_has_padding = item0 + BN > P.n;
        bool inserted = false;
        unsigned remask = 0;                         // AUC: thread rows whose minimum must be re-read after masking
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int row_l = ly * 4 + (i & 3) + (i >> 2) * 8;
            const int row = wrow0 + row_l;
            T s[NC];
            mt.row(i, s);
            if (P.bias != nullptr) {
#pragma unroll
                for (int c = 0; c < NC; c++) s[c] += biasv[c];
            }
            const T tau = rs->tau[row];
            T m = max_nan(max_nan(s[0], s[1]), max_nan(s[2], s[3]));
            if (NC == 8) m = max_nan(m, max_nan(max_nan(s[NC - 4], s[NC - 3]), max_nan(s[NC - 2], s[NC - 1])));
            if (!(m < tau)) {
                T* cs = P.cand_score + (size_t)rs->urow[row] * C;
                int* ci = P.cand_item + (size_t)rs->urow[row] * C;
                row_slow<T>(cs, ci, P.tri, P.n, tau, row, item0 + lx * 4, item0 + BN, s[0], s[1], s[2], s[3]);
                if (NC == 8)
                    row_slow<T>(cs, ci, P.tri, P.n, tau, row, item0 + 64 + lx * 4, item0 + BN, s[NC - 4], s[NC - 3], s[NC - 2], s[NC - 1]);
                inserted = true;
            }
            if (AUC) {
                if (tile_has_padding || rs->nxt_train[row] < item0 + BN) {
                    remask |= 1u << i;               // some of these are not candidates: exact minimum below
                } else {
                    T mn = fmin(fmin(s[0], s[1]), fmin(s[2], s[3]));      // fmin ignores NaN
                    if (NC == 8) mn = fmin(mn, fmin(fmin(s[NC - 4], s[NC - 3]), fmin(s[NC - 2], s[NC - 1])));
                    rowmin[i] = fmin(rowmin[i], mn);
                }
                T* dst = blk + (size_t)row_l * BN + lx * 4;
                sts4(dst, &s[0]);
                if (NC == 8) sts4(dst + 64, &s[NC - 4]);
            }
        }
        __syncwarp();

        if (AUC) {
            // mask (with NaN: never above a threshold, ignored by fmin) what is not a candidate in the
            // staged block: train items of rows whose train row intersects this tile, and the padding
            // columns of the last tile (lane r <-> row r)
            if (lane < 16) {
                const int row = wrow0 + lane;
                T* dst = blk + (size_t)lane * BN;
                if (rs->nxt_train[row] < item0 + BN) {
                    const int end = rs->end_train[row];
                    for (int c = rs->cur_train[row]; c < end; c++) {
                        const int item = P.tri[c];
                        if (item >= item0 + BN) break;
                        dst[item - item0] = NumTraits<T>::nan();
                    }
                }
                if (tile_has_padding)
                    for (int x = (P.n > item0 ? P.n - item0 : 0); x < BN; x++) dst[x] = NumTraits<T>::nan();
            }
            __syncwarp();
            if (__any_sync(FULL, remask != 0)) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (remask & (1u << i)) {
                        const int row_l = ly * 4 + (i & 3) + (i >> 2) * 8;
                        const T* src = blk + (size_t)row_l * BN + lx * 4;
                        T mn = NumTraits<T>::inf();
#pragma unroll
                        for (int c = 0; c < NC; c++) {
                            const T v = src[(c & 3) + (c >> 2) * 64];
                            mn = fmin(mn, v);                             // masked entries are NaN: ignored
                        }
                        rowmin[i] = fmin(rowmin[i], mn);
                    }
                }
            }
            // count: lane (ly, lx) owns held-out slots lx, lx+16, ... of rows 2*rp + ly
            for (int rp = 0; rp < 8; rp++) {
                const int row_l = 2 * rp + ly;
                const int row = wrow0 + row_l;
                const T* src = blk + (size_t)row_l * BN;
                {
                    const T thr[1] = {pj_s[row * 16 + lx]};
                    unsigned c[1] = {0};
                    count_above<T, 1>(src, thr, c);
                    cj_s[row * 16 + lx] += c[0];
                }
                // rows with more than 16 held-out items: further slots, 4 at a time, counters in global memory
                const int npos = rs->npos[row];
                if (npos > 16) {
                    const int tp0 = rs->tp0[row];
                    for (int j0 = 16 + lx; j0 < npos; j0 += 64) {
                        T thr[4];
                        unsigned c[4] = {0, 0, 0, 0};
#pragma unroll
                        for (int q = 0; q < 4; q++) thr[q] = (j0 + 16 * q < npos) ? P.pos_sorted[tp0 + j0 + 16 * q] : NumTraits<T>::inf();
                        count_above<T, 4>(src, thr, c);
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            if (c[q]) P.auc_cnt[(size_t)tp0 + j0 + 16 * q] += c[q];     // single owner: plain read-modify-write
                    }
                }
            }
            __syncwarp();
        }

        // advance the train cursors of the warp's rows past this tile (lanes 0..15, rare)
        if (lane < 16) {
            const int row = wrow0 + lane;
            int nxt = rs->nxt_train[row];
            if (nxt < item0 + BN) {
                int cur = rs->cur_train[row];
                const int end = rs->end_train[row];
                while (cur < end && (nxt = P.tri[cur]) < item0 + BN) cur++;
                rs->cur_train[row] = cur;
                rs->nxt_train[row] = cur < end ? nxt : INT_MAX;
            }
        }
        // cut back the buffers of this warp that passed the trigger (rare after the first few tiles)
        if (__any_sync(FULL, inserted)) {
            const int nv_l = lane < 16 ? rs->cnt[wrow0 + lane] : 0;
            unsigned need = __ballot_sync(FULL, nv_l > C - BN);
            while (need) {
                const int r = __ffs(need) - 1;
                need &= need - 1;
                const int row = wrow0 + r;
                const size_t base = (size_t)rs->urow[row] * C;
                compact_user<T, C>(P.cand_score + base, P.cand_item + base, rs->cnt[row], P.K, lane, &rs->tau[row], &rs->cnt[row]);
            }
        }
        __syncwarp();
    }

    // ---- leave the best min(cnt, K) candidates of every user at the head of its buffer ----
    for (int r = 0; r < 16; r++) {
        const int row = wrow0 + r;
        if (tile_u0 + row < P.mb) {
            const int ul = rs->urow[row];
            const size_t base = (size_t)ul * C;
            compact_user<T, C>(P.cand_score + base, P.cand_item + base, rs->cnt[row], P.K, lane, &rs->tau[row], &rs->cnt[row]);
            if (lane == 0) {
                P.cand_count[ul] = rs->cnt[row];
                if (rs->nan[row]) atomicOr(&P.uflags[P.user0 + ul], 1);
            }
        }
    }
    if (AUC) {
        for (int e = lane; e < 16 * 16; e += 32) {
            const int row = wrow0 + (e >> 4), j = e & 15;
            if (j < rs->npos[row]) P.auc_cnt[(size_t)rs->tp0[row] + j] = cj_s[row * 16 + j];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int row = wrow0 + ly * 4 + (i & 3) + (i >> 2) * 8;
            if (tile_u0 + row < P.mb && rs->npos[row] > 0 && rowmin[i] != NumTraits<T>::inf())
                atomicMin(&P.umin[P.user0 + rs->urow[row]], NumTraits<T>::orderable(rowmin[i]));
        }
    }
}

// One warp per user: order the (<= K <= 32*E) candidates score_select_kernel left at the head of the
// user's buffer (score descending, ties by ascending item id).
template <typename T, int E>
__global__ void rank_topk_kernel(T* __restrict__ cand_score, int* __restrict__ cand_item, const int* __restrict__ cand_count,
                                 const int C, const int mb)
{
    const int lane = threadIdx.x & 31;
    const int ul = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ul >= mb) return;
    const int nv = cand_count[ul];
    T* cs = cand_score + (size_t)ul * C;
    int* ci = cand_item + (size_t)ul * C;
    T s[E];
    int it[E];
#pragma unroll
    for (int e = 0; e < E; e++) {
        const int idx = e * 32 + lane;
        const bool v = idx < nv;
        s[e] = v ? cs[idx] : -NumTraits<T>::inf();
        it[e] = v ? ci[idx] : INT_MAX;
    }
    warp_sort_ranked<T, E>(s, it, lane);
#pragma unroll
    for (int e = 0; e < E; e++) {
        const int idx = e * 32 + lane;
        if (idx < nv) { cs[idx] = s[e]; ci[idx] = it[e]; }
    }
}

}  // namespace rmb
