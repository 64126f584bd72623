// filter_select.cuh -- tensor-core (tcgen05) candidate filter + exact re-scoring: the top-K path
// when no rank counting (ROC/PR-AUC) is requested.
//
// Same job as score_select.cuh (reference: candidate list /root/reference/src/recometrics.hpp:491-497,
// dot1 scoring :84-112/:499-512, partial_sort :537-548) with the work split differently:
//
//   1. FILTER  -- every (user, item) score is computed APPROXIMATELY on the 5th-generation tensor cores:
//      bf16 copies of the factors (item bias folded in as one more factor, as the reference's front-end
//      does, recometrics/__init__.py:548-551), fp32 accumulation in TMEM.  A CTA keeps a 128-user A tile
//      resident in shared memory and streams 128-item B tiles through a TMA ring; one elected thread
//      issues tcgen05.mma (M=128, N=128, K=16 per instruction) into one of two TMEM accumulator
//      buffers; four epilogue warps read the other buffer with tcgen05.ld -- thread t owns user row t
//      (TMEM lane t), so tau, counters and the candidate buffer of a user are private to one thread.
//      |approx - exact| <= margin_u = c * ||a_u|| * max_j ||b_j||  (bf16 rounding of both operands,
//      Cauchy-Schwarz; prep.cuh), so an item can only belong to the user's top K if
//      approx >= tau_u - margin_u, where tau_u is the EXACT K-th best score seen so far.  Everything
//      else (>= 99.9 % of the catalogue) is rejected with one compare on the approximate score.
//   2. EXACT   -- survivors are appended to the user's candidate buffer; when it fills, the warp
//      re-scores the new entries exactly (sequential fma chain over the original fp32 / fp64 factors:
//      the same chain, hence bit-identical scores, as score_select_kernel and score_entries_kernel),
//      drops train items / padding, and cuts the buffer back to the best K with the radix select of
//      score_select.cuh.  The final top-K, their order and their scores are therefore exactly those
//      of the FP32 (FP64) FMA path; the tensor cores only decide what is worth looking at.
//
// Shared-memory operand layout (no swizzle, K-major "interleaved"): [k/8][row][8 bf16] -- a core matrix
// is 8 rows x 16 bytes contiguous; descriptor LBO = 128 rows * 16 B (next k chunk), SBO = 128 B (next 8
// rows).  The pack kernel writes tiles in exactly this image, so one bulk copy per tile lands it.
// Encodings pinned on hardware by tools/ubench/umma_probe.cu.
#pragma once
#include <cuda_bf16.h>
#include "score_select.cuh"

namespace rmb {

constexpr int FN = 128;                  // items per MMA tile = TMEM columns per accumulator buffer
constexpr int F_EPI_WARPS = 4;           // 4 x 32 threads = 128 user rows = 128 TMEM lanes
constexpr int F_THREADS = (F_EPI_WARPS + 2) * 32;   // + TMA producer warp + MMA warp
constexpr int F_TMEM_COLS = 2 * FN;      // two accumulator buffers
constexpr int F_MAX_STAGES = 4;
constexpr int F_CHUNK = 32;              // TMEM columns per tcgen05.ld

template <typename T>
struct FilterParams {
    const __nv_bfloat16* __restrict__ Ab;   // [user tiles][KB/8][128][8]  bf16 user factors (+1.0 bias column)
    const __nv_bfloat16* __restrict__ Bb;   // [item tiles][KB/8][128][8]  bf16 item factors (+bias column)
    int KB;                                 // bf16 factors per row, multiple of 16
    int stages;                             // depth of the B ring
    int n, mb, user0;
    const T* __restrict__ At;               // exact user factors, tiled [user tile][p_pad][128]
    int p_pad, p;
    const T* __restrict__ Brow;             // exact item factors, row-major, leading dimension ldb
    size_t ldb;
    const T* __restrict__ bias;             // exact item biases or nullptr
    const float* __restrict__ anorm;        // [mb] ||a_u|| (with the 1.0 bias component), rounded up
    const unsigned* __restrict__ maxbn;     // float bits of max_j ||b_j|| (with the bias component)
    const int* __restrict__ trp;
    const int* __restrict__ tri;
    const int* __restrict__ ustatus;
    T* cand_score;                          // [mb_pad][C]
    int* cand_item;
    int* cand_count;
    int* uflags;
    int K;
};

inline size_t filter_smem_bytes(int KB, int stages, int p_pad, size_t elem)
{
    return (size_t)(1 + stages) * KB * 128 * 2 + (size_t)F_EPI_WARPS * p_pad * elem + 256 + 1024;
}

// ------------------------------------------------------------------ tcgen05 wrappers
__device__ __forceinline__ uint64_t umma_desc(const unsigned saddr)
{
    // K-major, no swizzle: LBO = 128 rows * 16 B = 2048 (>>4 = 128), SBO = 128 B (>>4 = 8), version 1
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)128 << 16) | ((uint64_t)8 << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_bf16(const unsigned tmem_d, const uint64_t adesc, const uint64_t bdesc,
                                          const unsigned idesc, const unsigned accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(const unsigned bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(const unsigned taddr, unsigned (&r)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float sub_down(const float tau, const float margin) { return __fsub_rd(tau, margin); }
__device__ __forceinline__ float sub_down(const double tau, const float margin) { return __double2float_rd(tau - (double)margin); }

// One warp: re-score the not yet exact entries [nproc, nv) of a user's candidate buffer, drop what is
// not a candidate (padding columns, train items: hpp:494-495; NaN scores raise the row's NaN flag),
// then cut the buffer back to the best K (score descending, ties by ascending item id).
// Returns the new entry count (all exact); tau is updated when at least K entries remain.
template <typename T, int C>
__device__ __noinline__ int filter_compact(T* cs, int* ci, const int nv, const int nproc, const int K, const int lane,
                                           T* a_sm, const T* __restrict__ At_user /* + k*128 */, const int p,
                                           const T* __restrict__ Brow, const size_t ldb, const T* __restrict__ bias, const int n,
                                           const int* __restrict__ tri, const int tr_lo, const int tr_hi,
                                           T* tau_io, int* nan_io)
{
    typedef typename NumTraits<T>::key_t key_t;
    constexpr int E = C / 32;
    for (int k = lane; k < p; k += 32) a_sm[k] = At_user[(size_t)k * BM];
    __syncwarp();
    key_t key[E];
    int it[E];
    int nanflag = 0;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const int idx = e * 32 + lane;
        key[e] = 0;                                             // 0 sorts below every score: "not an entry"
        it[e] = INT_MAX;
        if (idx < nv) {
            const int item = ci[idx];
            it[e] = item;
            if (idx < nproc) {
                key[e] = NumTraits<T>::key(cs[idx]);
            } else if (item < n && !in_train_segment(tri, tr_lo, tr_hi, item)) {
                const T* __restrict__ b = Brow + (size_t)item * ldb;
                T acc = (T)0;
#pragma unroll 8
                for (int k = 0; k < p; k++) acc = NumTraits<T>::fma(a_sm[k], b[k], acc);
                if (bias != nullptr) acc += bias[item];
                if (acc != acc) nanflag = 1;
                else key[e] = NumTraits<T>::key(acc);
            }
        }
    }
    nanflag = __any_sync(FULL, nanflag);
    int nvalid = 0;
#pragma unroll
    for (int e = 0; e < E; e++) nvalid += (key[e] != 0) ? 1 : 0;
    nvalid = __reduce_add_sync(FULL, nvalid);

    key_t t = 0;                  // keep key > t, and key == t with item id <= id_cut
    unsigned id_cut = 0x7fffffffu;
    if (nvalid >= K) {
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
        if (cge > K) {
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
    }
    __syncwarp();
    int base = 0;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const bool keep = (key[e] != 0) && ((key[e] > t) || (key[e] == t && (unsigned)it[e] <= id_cut));
        const unsigned mask = __ballot_sync(FULL, keep);
        if (keep) {
            const int pos = base + __popc(mask & ((1u << lane) - 1u));
            cs[pos] = NumTraits<T>::from_orderable((u64)key[e]);
            ci[pos] = it[e];
        }
        base += __popc(mask);
    }
    if (nvalid >= K) *tau_io = NumTraits<T>::from_orderable((u64)t);
    if (nanflag) *nan_io = 1;
    __syncwarp();
    return base;
}

template <typename T, int C>
__global__ void __launch_bounds__(F_THREADS, 1)
filter_select_kernel(const __grid_constant__ FilterParams<T> P)
{
    const int KB = P.KB, S = P.stages;
    const unsigned tile_bytes = (unsigned)KB * 128u * 2u;          // one operand tile (A or B)
    unsigned char* a_tile = smem_raw;                              // 1024-aligned dynamic shared memory
    unsigned char* b_ring = a_tile + tile_bytes;
    T* a_scratch = reinterpret_cast<T*>(b_ring + (size_t)S * tile_bytes);          // [F_EPI_WARPS][p_pad]
    u64* bars = reinterpret_cast<u64*>(reinterpret_cast<unsigned char*>(a_scratch) + (((size_t)F_EPI_WARPS * P.p_pad * sizeof(T) + 15) & ~size_t(15)));
    // barriers: full[S], empty[S], acc_full[2], acc_empty[2], a_full
    const unsigned bar_full = smem_u32(bars), bar_empty = bar_full + 8 * F_MAX_STAGES;
    const unsigned bar_accf = bar_empty + 8 * F_MAX_STAGES, bar_acce = bar_accf + 16, bar_a = bar_acce + 16;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 2 * F_MAX_STAGES + 5);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile_u0 = blockIdx.x * BM;
    const int NT = (P.n + FN - 1) / FN;

    if (tid == 0) {
        for (int s = 0; s < F_MAX_STAGES; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; b++) { mbar_init(bar_accf + 8 * b, 1); mbar_init(bar_acce + 8 * b, F_EPI_WARPS); }
        mbar_init(bar_a, 1);
        mbar_fence_init();
    }
    if (warp == F_EPI_WARPS + 1) {      // the MMA warp owns the tensor-memory allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((unsigned)F_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    if (warp == F_EPI_WARPS) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_a, tile_bytes);
            tma_bulk_g2s(smem_u32(a_tile), P.Ab + (size_t)blockIdx.x * KB * 128, tile_bytes, bar_a);
            for (int t = 0; t < NT; t++) {
                const int s = t % S;
                mbar_wait(bar_empty + 8 * s, ((t / S) & 1) ^ 1);          // first round passes immediately
                mbar_arrive_expect_tx(bar_full + 8 * s, tile_bytes);
                tma_bulk_g2s(smem_u32(b_ring + (size_t)s * tile_bytes), P.Bb + (size_t)t * KB * 128, tile_bytes, bar_full + 8 * s);
            }
        }
    } else if (warp == F_EPI_WARPS + 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            // D fp32 (bit 4), A bf16 (bit 7), B bf16 (bit 10), both K-major, N=128 (>>3 at bit 17), M=128 (>>4 at bit 24)
            const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(FN >> 3) << 17) | ((128u >> 4) << 24);
            mbar_wait(bar_a, 0);
            for (int t = 0; t < NT; t++) {
                const int s = t % S, b = t & 1;
                mbar_wait(bar_acce + 8 * b, ((t >> 1) & 1) ^ 1);          // accumulator buffer drained by the epilogue
                mbar_wait(bar_full + 8 * s, (t / S) & 1);                 // B tile landed
                tc_fence_after();
                const unsigned a0 = smem_u32(a_tile), b0 = smem_u32(b_ring + (size_t)s * tile_bytes);
                for (int ks = 0; ks < KB / 16; ks++)                      // K=16 per MMA = two 16-byte k chunks of 2048 B
                    umma_bf16(tmem_base + (unsigned)(b * FN), umma_desc(a0 + ks * 4096), umma_desc(b0 + ks * 4096), idesc, ks > 0 ? 1u : 0u);
                umma_commit(bar_empty + 8 * s);                           // stage reusable once these MMAs have read it
                umma_commit(bar_accf + 8 * b);                            // accumulator complete
            }
        }
    } else {
        // ===================== epilogue warps: thread <-> user row =====================
        const int row = warp * 32 + lane;
        const int ul = tile_u0 + row;
        const bool ranked = (ul < P.mb) && (P.ustatus[P.user0 + ul] == 0);
        T tau = -NumTraits<T>::inf();
        // |approx - exact| <= (2^-7 (1 + 2^-9) + k 2^-22) sum|a_k b_k| <= 0.0084 ||a|| ||b||: bf16 rounding of both
        // operands (relative 2^-8 each), fp32 accumulation in the tensor core, fp32 rounding of the exact chain
        const float margin = ranked ? 0.0084f * P.anorm[ul] * __uint_as_float(*P.maxbn) : 0.f;
        float thr = ranked ? -CUDART_INF_F : CUDART_INF_F;
        int cnt = 0, nproc = 0, nanrow = 0;
        T* cs = P.cand_score + (size_t)ul * C;
        int* ci = P.cand_item + (size_t)ul * C;
        int tr_lo = 0, tr_hi = 0;
        if (ranked) { tr_lo = P.trp[P.user0 + ul]; tr_hi = P.trp[P.user0 + ul + 1]; }
        T* a_sm = a_scratch + (size_t)warp * P.p_pad;

        auto compact_rows = [&](unsigned need) {
            while (need) {
                const int r = __ffs(need) - 1;
                need &= need - 1;
                const int ul_r = tile_u0 + warp * 32 + r;
                const int nv_r = __shfl_sync(FULL, cnt, r), np_r = __shfl_sync(FULL, nproc, r);
                const int lo_r = __shfl_sync(FULL, tr_lo, r), hi_r = __shfl_sync(FULL, tr_hi, r);
                T tau_r = __shfl_sync(FULL, tau, r);
                int nan_r = 0;
                const int kept = filter_compact<T, C>(P.cand_score + (size_t)ul_r * C, P.cand_item + (size_t)ul_r * C, nv_r, np_r, P.K, lane,
                                                      a_sm, P.At + (size_t)(ul_r / BM) * P.p_pad * BM + (ul_r % BM), P.p,
                                                      P.Brow, P.ldb, P.bias, P.n, P.tri, lo_r, hi_r, &tau_r, &nan_r);
                if (lane == r) {
                    cnt = kept; nproc = kept; tau = tau_r; nanrow |= nan_r;
                    thr = sub_down(tau, margin);
                }
            }
        };

        for (int t = 0; t < NT; t++) {
            const int b = t & 1;
            mbar_wait(bar_accf + 8 * b, (t >> 1) & 1);
            tc_fence_after();
            for (int c = 0; c < FN / F_CHUNK; c++) {
                unsigned v[32];
                tmem_ld32(tmem_base + ((unsigned)(warp * 32) << 16) + (unsigned)(b * FN + c * F_CHUNK), v);
                if (c == FN / F_CHUNK - 1) {            // accumulator buffer fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_acce + 8 * b);
                }
                float mx[16];
#pragma unroll
                for (int j = 0; j < 16; j++) mx[j] = max_nan(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
#pragma unroll
                for (int w = 8; w > 0; w >>= 1)
#pragma unroll
                    for (int j = 0; j < w; j++) mx[j] = max_nan(mx[j], mx[j + w]);
                if (__any_sync(FULL, !(mx[0] < thr))) {
                    const int item_base = t * FN + c * F_CHUNK;
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float s = __uint_as_float(v[j]);
                        if (!(s < thr) && ranked) { cs[cnt] = (T)s; ci[cnt] = item_base + j; cnt++; }
                    }
                    const unsigned need = __ballot_sync(FULL, cnt > C - F_CHUNK);
                    if (need) compact_rows(need);
                }
            }
        }
        // exact scores for whatever is still pending, best K at the head of every buffer
        compact_rows(__ballot_sync(FULL, ranked && cnt > 0));
        if (ul < P.mb) {
            P.cand_count[ul] = ranked ? cnt : 0;
            if (nanrow) atomicOr(&P.uflags[P.user0 + ul], 1);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == F_EPI_WARPS + 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((unsigned)F_TMEM_COLS));
}

}  // namespace rmb
