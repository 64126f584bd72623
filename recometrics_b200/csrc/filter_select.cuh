// filter_select.cuh -- tensor-core (tcgen05) candidate filter + exact re-scoring: the top-K path
// when no rank counting (ROC/PR-AUC) is requested.
//
// Same job as score_select.cuh (reference: candidate list /root/reference/src/recometrics.hpp:491-497,
// dot1 scoring :84-112/:499-512, partial_sort :537-548) with the work split differently:
//
//   1. FILTER (filter_select_kernel) -- every (user, item) score is computed APPROXIMATELY on the
//      5th-generation tensor cores: fp16 copies of the factors, scaled by powers of two so that every element
//      is at most 1 (prep.cuh; item bias folded in as one more factor, as the reference's front-end does,
//      recometrics/__init__.py:548-551), fp32 accumulation in TMEM.
//      A CTA keeps a 128-user A tile resident in shared memory and streams 128-item B tiles through a
//      TMA ring; one elected lane issues tcgen05.mma (M=128, N=128, K=16 per instruction) into one of
//      four TMEM accumulator buffers; sixteen epilogue warps read them back with tcgen05.ld: the four
//      warps of a TMEM lane quarter (32 user rows; they also share a warp scheduler) each take one
//      32-column chunk of every tile.  Everything a warp needs per row is PRIVATE to it -- its own region
//      of the row's candidate buffer with the counter in a register, its own cursor into the row's train
//      items -- so the four warps never wait for each other inside a tile: while one of them appends
//      candidates, the others run ahead (up to the four accumulator buffers).  They meet at a named
//      barrier only every 1..16 tiles (the distance adapts to the append rate) to cut the row's four
//      regions back together and raise the row's shared threshold.
//      With m_u = c * ||a_u|| * max_j ||b_j|| >= |approx - exact| (fp16 rounding of both operands,
//      Cauchy-Schwarz; filter_err_coef) and tau~ = the K-th best APPROXIMATE candidate score seen so far, every
//      member of the exact top K satisfies approx >= tau~ - 2 m_u (the K best approximate scores have
//      exact scores >= tau~ - m_u, so the exact K-th best is >= tau~ - m_u, and a member's approximate
//      score is within m_u of its exact one).  The kernel therefore keeps, per user, all candidates with
//      approx >= tau~ - 2 m_u: train items and padding columns are dropped when appended, the regions are
//      cut back with a histogram select on the approximate keys.  >= 99.9 % of the catalogue
//      is rejected with one compare.
//   2. EXACT (exact_topk_kernel) -- one warp per user re-scores the few hundred survivors exactly
//      (sequential fma chain over the original fp32 / fp64 factors: the same chain, hence bit-identical
//      scores, as score_select_kernel / score_entries_kernel) and keeps the best K.  The final top-K,
//      their order and their scores are exactly those of the FP32 (FP64) FMA path; the tensor cores only
//      decide what is worth looking at.  A user whose slack band does not fit the buffer is flagged and
//      the host re-runs that batch on the FMA path.
//
// Shared-memory operand layout (no swizzle, K-major "interleaved"): [k/8][row][8 halves] -- a core matrix
// is 8 rows x 16 bytes contiguous; descriptor LBO = 128 rows * 16 B (next k chunk), SBO = 128 B (next 8
// rows).  The pack kernel writes tiles in exactly this image, so one bulk copy per tile lands it.
// Encodings pinned on hardware by tools/ubench/umma_probe.cu.
#pragma once
#include <cuda_fp16.h>
#include "prep.cuh"
#include "score_select.cuh"
#include "tie_noise.cuh"

namespace rmb {

constexpr int FN = 128;                  // items per MMA tile = TMEM columns per accumulator buffer
constexpr int F_EPI_WARPS = 16;          // warp w reads TMEM lanes (user rows) 32*(w%4)..+31, column chunk w/4 of every tile
#ifndef RMB_F_MMA_WARPS
#define RMB_F_MMA_WARPS 2                // MMA-issuing warps, taking turns tile by tile: one warp's per-tile barrier waits and commits
#endif                                   // (~300 cycles in which it issues nothing) overlap with the other warp's MMAs
constexpr int F_MMA_WARPS = RMB_F_MMA_WARPS;
constexpr int F_THREADS = (F_EPI_WARPS + 1 + F_MMA_WARPS) * 32;   // + TMA producer warp + MMA warps
#ifndef RMB_F_ACCBUFS
#define RMB_F_ACCBUFS 4                  // TMEM accumulator buffers (2 or 4): slack between the MMA warp and the slowest epilogue warp
#endif
constexpr int F_ACCBUFS = RMB_F_ACCBUFS;
constexpr int F_ACCSHIFT = F_ACCBUFS == 4 ? 2 : 1;
constexpr int F_TMEM_COLS = F_ACCBUFS * FN;
constexpr int F_MAX_STAGES = 6;              // ring depth: as many B tiles as fit next to the A tile (5 at 128 factors)
constexpr int F_CHUNK = 32;              // TMEM columns per tcgen05.ld
#ifndef RMB_F_CUT_MARGIN
#define RMB_F_CUT_MARGIN 48               // a row's regions are cut back together once they hold this many more than the last cut kept
#endif
constexpr int F_CUT_MARGIN = RMB_F_CUT_MARGIN;
#ifndef RMB_F_MAX_MEET
#define RMB_F_MAX_MEET 64                 // most item tiles between two meetings of a quarter's warps
#endif
constexpr int F_MAX_MEET = RMB_F_MAX_MEET;

#ifndef RMB_F_DBG
#define RMB_F_DBG 0                      // developer build: honour FilterParams::dbg (env RMB200_DBG) inside the tile loop
#endif
#ifndef RMB_F_STATS
#define RMB_F_STATS 0                    // developer build: count appends / cuts / slow-path entries / cursor moves into FilterParams::retries[1..4]
#endif
#if RMB_F_STATS
#define F_STAT(i, v) atomicAdd(&rs->stat[i], (v))
#else
#define F_STAT(i, v) do { } while (0)
#endif

struct FilterParams {
    const __half* __restrict__ Ab;          // [user tiles][KB/8][128][8]  fp16 user factors (+1.0 bias column), each row scaled by pow2_scale_for(||a_u||)
    const __half* __restrict__ Bb;          // [item tiles][KB/8][128][8]  fp16 item factors (+bias column), all scaled by pow2_scale_for(max_j ||b_j||)
    int KB;                                 // fp16 factors per row, multiple of 16
    float err_coef;                         // c: |approx - exact| <= c ||a|| max||b|| (host: filter_err_coef)
    float noise_band;                       // break_ties_with_noise: how far the noise can move two scores apart (2e-12), else 0
    int stages;                             // depth of the B ring
    int n, mb, user0;
    const float* __restrict__ anorm;        // [mb] ||a_u|| (with the 1.0 bias component), rounded up
    const unsigned* __restrict__ maxbn;     // float bits of max_j ||b_j|| (with the bias component)
    const int* __restrict__ trp;
    const int* __restrict__ tri;
    const int* __restrict__ ustatus;
    float* cand_approx;                     // [mb_pad][C] approximate scores of the kept candidates
    int* cand_item;                         // [mb_pad][C]
    int* cand_count;                        // [mb_pad] kept candidates; -1 = slack band overflowed the buffer
    int* overflow;                          // number of users flagged -1
    int* uflags;                            // [m] bit0: a candidate score was NaN
    int K;
    int sample_tiles, sample_stride, sample_rank;   // pass 0 (see filter_pass_tile); sample_tiles == 0: no sampling
    int* retries;                           // rows that needed the retry pass (statistics; may be nullptr)
    int dbg;                                // developer switch (env RMB200_DBG): 1 = skip the scan (pipeline ceiling), 2 = no row is ranked (fast path only), 4 = no meetings, 8 = no tcgen05.ld, 16 = no TMA after the first ring round, 32 = half of the k steps (bits honoured only in -DRMB_F_DBG=1 builds)
};

struct FilterRowState {      // per user row of the CTA, shared by the four epilogue warps of its TMEM lane quarter
    float thr[BM], slack[BM], guess[BM];
    int cnt[BM];                     // candidates left contiguous at the head of the row's buffer when a pass ends
    int cnt4[4][BM];                 // candidates in each warp's region of the buffer, published at the meeting points
    int trig[BM];                    // total above which the regions are cut back together
    int flags[BM];                   // 1 = NaN candidate score, 2 = buffer overflow (slack band / burst), 4 = guess failed (retry pass)
    int nxt2_train[4][BM];           // per warp: the train item id after the cursor's next one (prefetched with cp.async)
    int retry;                       // some row of the CTA needs the retry pass
    int stat[4];                     // RMB_F_STATS: appends, cuts, slow-path entries (8-column groups), cursor moves
    unsigned hist[F_EPI_WARPS][32];  // bucket counters of cut_regions, one set per epilogue warp
};


// Error bound of the filter's approximate scores, as a multiple of ||a_u|| max_j ||b_j||:
//   fp16 rounding of both operands (scaled so that |x| <= 1; relative 2^-11 each, absolute 2^-25 below 6e-5):
//       sum_k |a_k b_k| (2^-10 + 2^-22) + 2^-25 sqrt(k) (||a|| + ||b||)   <=  (2^-10 (1 + 2^-10) + 4 sqrt(k) 2^-24) ||a|| ||b||
//       (Cauchy-Schwarz; the scaled norms are >= 0.5, hence the factor 4 on the absolute term)
//   fp32 accumulation inside the tensor core plus the rounding of the exact fp32 fma chain it is compared with: k 2^-22
// and 1 % on top.
inline float filter_err_coef(int KB)
{
    const double c = std::ldexp(1.0, -10) * (1.0 + std::ldexp(1.0, -10)) + 4.0 * std::sqrt((double)KB) * std::ldexp(1.0, -24) + KB * std::ldexp(1.0, -22);
    return (float)(1.01 * c);
}

inline size_t filter_smem_fixed_bytes() { return 256 + sizeof(FilterRowState); }      // barriers + row state
inline size_t filter_smem_bytes(int KB, int stages)
{
    return (size_t)(1 + stages) * KB * 128 * 2 + filter_smem_fixed_bytes();
}

// ------------------------------------------------------------------ tcgen05 wrappers
__device__ __forceinline__ uint64_t umma_desc(const unsigned saddr)
{
    // K-major, no swizzle: LBO = 128 rows * 16 B = 2048 (>>4 = 128), SBO = 128 B (>>4 = 8), version 1
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)128 << 16) | ((uint64_t)8 << 32) | (1ull << 46);
}
// the same descriptor `bytes` further into the tile: only the 14-bit address field (bytes >> 4) of the low word moves
// (shared-memory addresses stay below 2^18, so the field cannot carry into the LBO field)
__device__ __forceinline__ uint64_t umma_desc_advance(const uint64_t desc, const unsigned bytes) { return desc + (uint64_t)(bytes >> 4); }
__device__ __forceinline__ void umma_f16(const unsigned tmem_d, const uint64_t adesc, const uint64_t bdesc,
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
// one lane of a converged warp (elect.sync): the compiler keeps what the elected lane computes in uniform registers
__device__ __forceinline__ bool elect_one()
{
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
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

// ------------------------------------------------------------------ CTA pair (tcgen05 cta_group::2), see filter_select_kernel<C, true>
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared-memory address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ unsigned mapa_shared(const unsigned saddr, const unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(const unsigned cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ unsigned ld_shared_cluster_u32(const unsigned cluster_addr)
{
    unsigned v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(cluster_addr) : "memory");
    return v;
}
// descriptor of a 64-row half tile: LBO = 64 rows * 16 B = 1024 (>>4 = 64)
__device__ __forceinline__ uint64_t umma_desc_half(const unsigned saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)64 << 16) | ((uint64_t)8 << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_f16_pair(const unsigned tmem_d, const uint64_t adesc, const uint64_t bdesc,
                                               const unsigned idesc, const unsigned accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of the MMAs issued so far -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(const unsigned bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((unsigned short)3) : "memory");
}

// One warp: cut a row's candidates back to those that can still belong to the exact top K: approximate score >=
// (K-th best approximate score) - slack.  The row's buffer is four regions of C/4 entries (one per epilogue warp of the
// quarter), region s holding cnt4[s] entries; all of them are read into registers first, so the survivors can be written
// back in place: CONTIG ? packed at the head of the buffer : dealt round-robin over the four regions.
// The K-th best key is found by ROUNDS rounds of a 32-bucket histogram (shared-memory counters, one bucket per lane,
// suffix sums by shuffles) that narrow a window [lo, lo + span] of the key range around it.  The window's lower edge
// always satisfies #(key >= lo) >= K, so it is a valid (slightly low: window width = key range / 32^ROUNDS) stand-in
// for the K-th best approximate score; ROUNDS = 7 exhausts 32-bit keys and makes it exact.
// K <= total entries is required.  Returns the number kept; *tau_out = the (stand-in for the) K-th best score.
#ifndef RMB_F_HIST_ROUNDS
#define RMB_F_HIST_ROUNDS 2
#endif
template <int C, int ROUNDS, bool CONTIG>
__device__ __noinline__ int cut_regions(float* cs, int* ci, const int c0, const int c1, const int c2, const int c3, const int K,
                                        const float slack, const int lane, unsigned* hist, float* tau_out)
{
    constexpr int E = C / 32, RC = C / 4, EPR = E / 4;      // entries per lane, region capacity, per-lane entries per region
    unsigned key[E];
    int it[E];
    unsigned kmax = 0u, kmin = 0xffffffffu;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const int region = e / EPR, pos = (e % EPR) * 32 + lane;
        const int cr = region == 0 ? c0 : (region == 1 ? c1 : (region == 2 ? c2 : c3));
        const bool v = pos < cr;
        const int idx = region * RC + pos;
        key[e] = v ? NumTraits<float>::key(cs[idx]) : 0u;      // 0 sorts below every score
        it[e] = v ? ci[idx] : INT_MAX;
        if (v) { kmax = max(kmax, key[e]); kmin = min(kmin, key[e]); }
    }
    kmax = __reduce_max_sync(FULL, kmax);
    kmin = __reduce_min_sync(FULL, kmin);
    unsigned lo = kmin, span = kmax - kmin;     // window [lo, lo + span]; the need-th largest key inside it is wanted
    int need = K;
#pragma unroll 1
    for (int round = 0; round < ROUNDS; round++) {
        const int sh = max(0, 27 - __clz(span));               // (key - lo) >> sh < 32 inside the window
        hist[lane] = 0u;
        __syncwarp();
#pragma unroll
        for (int e = 0; e < E; e++) {
            const unsigned d = key[e] - lo;
            if (key[e] >= lo && d <= span) atomicAdd(&hist[d >> sh], 1u);
        }
        __syncwarp();
        unsigned suf = hist[lane];                             // -> number of window keys in buckets >= lane
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_down_sync(FULL, suf, o);
            if (lane + o < 32) suf += t;
        }
        const unsigned mask = __ballot_sync(FULL, (int)suf >= need);   // lane 0 always votes: the window holds >= need keys
        const int b = 31 - __clz(mask);
        const unsigned above = __shfl_sync(FULL, suf, (b + 1) & 31);
        if (b < 31) need -= (int)above;
        const unsigned off = (unsigned)b << sh;
        lo += off;
        span = min(span - off, (1u << sh) - 1u);
        if (sh == 0) break;
    }
    const float tau = NumTraits<float>::from_orderable((u64)lo);
    const float cutf = __fsub_rd(tau, slack);                   // NaN (tau or slack not finite): keep everything
    const unsigned cut = (cutf == cutf) ? NumTraits<float>::key(cutf) : 1u;
    int base = 0;
    __syncwarp();
#pragma unroll
    for (int e = 0; e < E; e++) {
        const bool keep = key[e] >= cut && key[e] != 0u;
        const unsigned mask = __ballot_sync(FULL, keep);
        if (keep) {
            const int k = base + __popc(mask & ((1u << lane) - 1u));
            const int pos = CONTIG ? k : (k & 3) * RC + (k >> 2);
            cs[pos] = NumTraits<float>::from_orderable((u64)key[e]);
            ci[pos] = it[e];
        }
        base += __popc(mask);
    }
    *tau_out = tau;
    __syncwarp();
    return base;
}

// Slow path of the filter, out of line and in ONE copy (the tile loop must stay small: sixteen warps share the
// instruction cache): the 8 scores of a column group of which at least one is not below the row's threshold.
// Padding columns and train items are dropped (hpp:494-495; [t_lo, t_hi) = the part of the row's sorted train items
// that can intersect this tile, empty when none does: the few inside the tile are scanned linearly), NaN scores raise
// flag 1 (hpp:195-197), the rest is appended to this warp's region of the row's candidate buffer (flag 2 when it is full).
// Returns the new fill of the region.
template <int RC>
__device__ __noinline__ int filter_append_group(const float s0, const float s1, const float s2, const float s3,
                                                const float s4, const float s5, const float s6, const float s7,
                                                const float thr, const int item0, const int n,
                                                const int* __restrict__ tri, const int t_lo, const int t_hi,
                                                float* cs, int* ci, int mycnt, int* flagp)
{
    // which of the 8 pass (branch-free), then one trip per passing score (almost always a single one)
    unsigned m = (s0 < thr ? 0u : 1u) | (s1 < thr ? 0u : 2u) | (s2 < thr ? 0u : 4u) | (s3 < thr ? 0u : 8u) |
                 (s4 < thr ? 0u : 16u) | (s5 < thr ? 0u : 32u) | (s6 < thr ? 0u : 64u) | (s7 < thr ? 0u : 128u);
    while (m) {
        const int jj = __ffs(m) - 1;
        m &= m - 1;
        const float lo4 = (jj & 2) ? ((jj & 1) ? s3 : s2) : ((jj & 1) ? s1 : s0);
        const float hi4 = (jj & 2) ? ((jj & 1) ? s7 : s6) : ((jj & 1) ? s5 : s4);
        const float s = (jj & 4) ? hi4 : lo4;
        const int item = item0 + jj;
        if (item >= n) break;                                   // padding columns (and all after them)
        bool in_train = false;
        for (int lo = t_lo; lo < t_hi; lo++) {
            const int v = tri[lo];
            if (v >= item) { in_train = (v == item); break; }
        }
        if (in_train) continue;
        if (s != s) atomicOr(flagp, 1);
        else if (mycnt < RC) { cs[mycnt] = s; ci[mycnt] = item; mycnt++; }
        else atomicOr(flagp, 2);
    }
    return mycnt;
}

// The item tiles a CTA walks, pass by pass (all roles -- TMA producer, MMA issuer, epilogue -- step through the same list):
//   pass 0  SAMPLE  every sample_stride-th tile (sample_tiles of them, ~1/16 of the catalogue; skipped when sample_tiles == 0).
//                   The rows run the same streaming selection with K = sample_rank and no slack; the sample_rank-th best
//                   approximate score of the sample becomes the row's GUESS g: with sample_rank chosen by the host so that
//                   P(fewer than K of ALL items reach the sample's sample_rank-th best) <= 1e-6 per row.
//   pass 1  MAIN    every tile, thresholds start at g - slack instead of -inf.  A streaming top-K appends K' (1 + ln(n / K'))
//                   candidates per row, more than half of them in the first percent of the catalogue while the threshold is
//                   still loose; starting from the guess leaves about (items >= g - slack) appends, 3x fewer at 1M items.
//                   The guess is VERIFIED: the pass is valid for a row iff at least K of its kept candidates reach g
//                   (then the K-th best approximate score tau~ >= g and everything >= tau~ - slack was kept).
//   pass 2  RETRY   only if some row of the CTA failed the check (or had too few sample candidates): every tile again,
//                   failed rows from -inf, the others closed.  Costs the CTA a second walk; never changes a result.
__device__ __forceinline__ int filter_pass_tiles(const FilterParams& P, const int pass, const int NT) { return pass == 0 ? P.sample_tiles : NT; }
__device__ __forceinline__ int filter_pass_tile(const FilterParams& P, const int pass, const int j) { return pass == 0 ? j * P.sample_stride : j; }

// PAIR (launched as clusters of two CTAs; EXPERIMENTAL, off unless RMB200_PAIR=1 -- written at the end of round 1 on top of
// tools/ubench/umma_pair_probe.cu and not yet tuned): the two CTAs of a cluster run ONE tcgen05.mma.cta_group::2 of 256 users x
// 128 items per k-step.  Each CTA keeps its own 128 user rows, its own accumulators (its TMEM holds its rows of D, so the
// epilogue below is unchanged) and only HALF of every item tile (CTA r: items r*64..r*64+63 of the tile, 16 KB bulk copy from a
// B image packed in halves): the L2 -> shared-memory traffic that bounds the pipeline is halved.  The leader (rank 0) issues
// the MMAs once both halves have landed (the peer's MMA warps forward their `full` barrier with a remote arrive) and both
// CTAs' epilogues have drained the accumulator buffer (the peer's epilogue warps arrive remotely on the leader's `acc_empty`);
// its commits are multicast to `empty` / `acc_full` of both CTAs.  The retry pass is taken by both CTAs if either needs it.
template <int C, bool PAIR>
__global__ void __launch_bounds__(F_THREADS, 1)
filter_select_kernel(const __grid_constant__ FilterParams P)
{
    const int KB = P.KB, S = P.stages;
    const unsigned tile_bytes = (unsigned)KB * 128u * 2u;          // one operand tile (A or B)
    const unsigned b_bytes = PAIR ? tile_bytes / 2u : tile_bytes;  // what one ring stage receives (the stage slots keep their size)
    const unsigned rank = PAIR ? cluster_ctarank() : 0u;
    unsigned char* a_tile = smem_raw;
    unsigned char* b_ring = a_tile + tile_bytes;
    u64* bars = reinterpret_cast<u64*>(b_ring + (size_t)S * tile_bytes);
    // barriers: full[4], empty[4], acc_full[4], acc_empty[4], a_full
    const unsigned bar_full = smem_u32(bars), bar_empty = bar_full + 8 * F_MAX_STAGES;
    const unsigned bar_accf = bar_empty + 8 * F_MAX_STAGES, bar_acce = bar_accf + 32, bar_a = bar_acce + 32;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 2 * F_MAX_STAGES + 9);
    const unsigned bar_pfull = bar_a + 16;                         // PAIR, leader: the peer's half of stage s has landed
    FilterRowState* rs = reinterpret_cast<FilterRowState*>(reinterpret_cast<unsigned char*>(bars) + 256);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile_u0 = blockIdx.x * BM;
    const int NT = (P.n + FN - 1) / FN;

    if (tid == 0) {
        for (int s = 0; s < F_MAX_STAGES; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < F_ACCBUFS; b++) { mbar_init(bar_accf + 8 * b, 1); mbar_init(bar_acce + 8 * b, PAIR ? 2 * F_EPI_WARPS : F_EPI_WARPS); }
        mbar_init(bar_a, 1);
        if (PAIR) for (int s = 0; s < F_MAX_STAGES; s++) mbar_init(bar_pfull + 8 * s, 1);
        mbar_fence_init();
        rs->retry = 0;
        for (int i = 0; i < 4; i++) rs->stat[i] = 0;
    }
    if (warp == F_EPI_WARPS + 1) {      // the MMA warp owns the tensor-memory allocation (PAIR: one warp of each CTA)
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((unsigned)F_TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((unsigned)F_TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();       // both CTAs: barriers initialised, tensor memory allocated
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    if (warp == F_EPI_WARPS && elect_one()) {
        mbar_arrive_expect_tx(bar_a, tile_bytes);
        tma_bulk_g2s(smem_u32(a_tile), P.Ab + (size_t)blockIdx.x * KB * 128, tile_bytes, bar_a);
    }

    int it0 = 0;                         // pipeline iterations before this pass (ring stage / accumulator buffer / parities follow it)
    for (int pass = P.sample_tiles > 0 ? 0 : 1; pass < 3; pass++) {
        const int ntiles = filter_pass_tiles(P, pass, NT);
        if (warp == F_EPI_WARPS) {
            // ===================== TMA producer (the warp stays converged, one elected lane issues) =====================
            int s = it0 % S, ph = (it0 / S) & 1;                             // ring stage and its phase parity
            for (int j = 0; j < ntiles; j++) {
                mbar_wait(bar_empty + 8 * s, ph ^ 1);                         // first round passes immediately
                if (elect_one()) {
#if RMB_F_DBG
                    if ((P.dbg & 16) && it0 + j >= S) { mbar_arrive(bar_full + 8 * s); } else       // 16: no TMA after the first round (stale tiles)
#endif
                    {
                    mbar_arrive_expect_tx(bar_full + 8 * s, b_bytes);
                    // (PAIR: the item matrix is packed in 64-item halves, [tile][half][k/8][64][8]; this CTA takes half `rank`)
                    tma_bulk_g2s(smem_u32(b_ring) + (unsigned)s * tile_bytes,
                                 P.Bb + (PAIR ? ((size_t)filter_pass_tile(P, pass, j) * 2 + rank) * KB * 64 : (size_t)filter_pass_tile(P, pass, j) * KB * 128),
                                 b_bytes, bar_full + 8 * s);
                    }
                }
                __syncwarp();
                if (++s == S) { s = 0; ph ^= 1; }
            }
        } else if (warp > F_EPI_WARPS) {
            // ===================== MMA issuers (converged warps, one elected lane issues; warp w takes iterations it % F_MMA_WARPS == w) =====================
            // D fp32 (bit 4), A and B fp16 (format 0 at bits 7 and 10), both K-major, N=128 (>>3 at bit 17), M=128 (>>4 at bit 24)
            const unsigned idesc = (1u << 4) | ((unsigned)(FN >> 3) << 17) | (((PAIR ? 256u : 128u) >> 4) << 24);
            const uint64_t adesc0 = umma_desc(smem_u32(a_tile)), bdesc0 = PAIR ? umma_desc_half(smem_u32(b_ring)) : umma_desc(smem_u32(b_ring));
            const unsigned b_kstep = PAIR ? 2048u : 4096u;                    // bytes between two K=16 steps of the B operand
            int ksteps = KB / 16;                                             // K=16 per MMA = two 16-byte k chunks of 2048 B
#if RMB_F_DBG
            if (P.dbg & 32) ksteps = ksteps / 2 > 0 ? ksteps / 2 : 1;         // 32: half of the k steps
#endif
            const int w = warp - (F_EPI_WARPS + 1);
            mbar_wait(bar_a, 0);
            int j = (w - it0 % F_MMA_WARPS + F_MMA_WARPS) % F_MMA_WARPS;      // first iteration of this pass that is this warp's
            int s = (it0 + j) % S, ph = ((it0 + j) / S) & 1;
            for (; j < ntiles; j += F_MMA_WARPS) {
                const int it = it0 + j, b = it & (F_ACCBUFS - 1);
                if (PAIR && rank != 0) {
                    // peer CTA: its half of the tile has landed -> tell the leader; the leader's MMAs do the rest
                    mbar_wait(bar_full + 8 * s, ph);
                    if (elect_one()) mbar_arrive_cluster(mapa_shared(bar_pfull + 8 * s, 0u));
                    __syncwarp();
                    s += F_MMA_WARPS;
                    while (s >= S) { s -= S; ph ^= 1; }
                    continue;
                }
                mbar_wait(bar_acce + 8 * b, ((it >> F_ACCSHIFT) & 1) ^ 1);     // accumulator buffer drained by the epilogue (PAIR: of both CTAs)
                mbar_wait(bar_full + 8 * s, ph);                              // B tile landed
                if (PAIR) mbar_wait(bar_pfull + 8 * s, ph);                   // ... and the peer's half
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t bdesc = umma_desc_advance(bdesc0, (unsigned)s * tile_bytes);
                    const unsigned d = tmem_base + (unsigned)(b * FN);
                    if (PAIR) {
                        umma_f16_pair(d, adesc0, bdesc, idesc, 0u);
#pragma unroll 7
                        for (int ks = 1; ks < ksteps; ks++)
                            umma_f16_pair(d, umma_desc_advance(adesc0, ks * 4096), umma_desc_advance(bdesc, ks * b_kstep), idesc, 1u);
                        umma_commit_pair(bar_empty + 8 * s);
                        umma_commit_pair(bar_accf + 8 * b);
                    } else {
                        umma_f16(d, adesc0, bdesc, idesc, 0u);
#pragma unroll 7
                        for (int ks = 1; ks < ksteps; ks++)
                            umma_f16(d, umma_desc_advance(adesc0, ks * 4096), umma_desc_advance(bdesc, ks * 4096), idesc, 1u);
                        umma_commit(bar_empty + 8 * s);                       // stage reusable once these MMAs have read it
                        umma_commit(bar_accf + 8 * b);                        // accumulator complete
                    }
                }
                __syncwarp();
                s += F_MMA_WARPS;
                while (s >= S) { s -= S; ph ^= 1; }
            }
        } else {
            // ===================== epilogue warps =====================
            constexpr int RC = C / 4;                       // capacity of a warp's region of the row's candidate buffer
            const int q = warp & 3, slot = warp >> 2;       // TMEM lane quarter, 32-column chunk of the tile
            const int row = q * 32 + lane;                  // user row of this thread (shared with the 3 other slots)
            const int ul = tile_u0 + row;
            const unsigned qbar = 1 + q;                    // named barrier of the quarter's four warps
            const int Kp = pass == 0 ? P.sample_rank : P.K; // how many best candidates the pass keeps track of
            float* cs = P.cand_approx + (size_t)ul * C + slot * RC;      // this warp's region of the row's buffer
            int* ci = P.cand_item + (size_t)ul * C + slot * RC;
            if (slot == 0) {
                const bool ranked = (ul < P.mb) && (P.ustatus[P.user0 + ul] == 0) && !(RMB_F_DBG && (P.dbg & 2));
                // |approx - exact| <= c ||a|| max||b|| (c = P.err_coef, see filter_err_coef); the approximate scores live in the
                // scaled units of the operand image (row scale x matrix scale, both powers of two), and so does the slack
                float slack = 0.f;
                if (ranked) {
                    const float an = P.anorm[ul], bn = __uint_as_float(*P.maxbn);
                    const float sa = pow2_scale_for(an), sb = pow2_scale_for(bn);
                    slack = 2.f * P.err_coef * (an * sa) * (bn * sb) + P.noise_band * sa * sb;
                }
                float thr0 = CUDART_INF_F;                  // approx < thr: cannot be in the exact top K; +inf: the row is closed
                if (pass == 0) {
                    rs->flags[row] = 0;
                    rs->guess[row] = -CUDART_INF_F;
                    if (ranked) thr0 = -CUDART_INF_F;
                } else if (pass == 1) {
                    if (P.sample_tiles == 0) { rs->flags[row] = 0; rs->guess[row] = -CUDART_INF_F; }
                    if (ranked) { thr0 = __fsub_rd(rs->guess[row], slack); if (!(thr0 == thr0)) thr0 = -CUDART_INF_F; }
                } else {
                    if (rs->flags[row] & 4) thr0 = -CUDART_INF_F;
                }
                rs->slack[row] = pass == 0 ? 0.f : slack;
                rs->thr[row] = thr0;
                rs->trig[row] = Kp + F_CUT_MARGIN;
                if (thr0 != CUDART_INF_F || pass < 2) rs->cnt[row] = 0;   // (a row closed in the retry pass keeps what pass 1 found)
            }
            asm volatile("bar.sync %0, 128;" ::"r"(qbar) : "memory");
            // private per-thread row state: region fill, cursor into the row's sorted train items (hpp:494-495 takes those out
            // of the pool): t_nxt = the first train item id >= the current tile, the one after it prefetched in shared memory
            int mycnt = 0, t_cur = 0, t_end = 0, t_nxt = INT_MAX;
            if (rs->thr[row] != CUDART_INF_F) {
                t_cur = P.trp[P.user0 + ul]; t_end = P.trp[P.user0 + ul + 1];
                if (t_cur < t_end) t_nxt = P.tri[t_cur];
                rs->nxt2_train[slot][row] = t_cur + 1 < t_end ? P.tri[t_cur + 1] : INT_MAX;
            }
            int tot_prev = 0;                               // row total after the previous meeting (the same in all four warps)
            int meet_in = 0, interval = 1;                  // tiles until the next meeting; tiles since the previous one
            // loop-carried pipeline state, kept incrementally (no per-tile index arithmetic): accumulator buffer and its phase
            // parity, first item id of this warp's 32-column chunk, the item step between two tiles of the pass
            int buf = it0 & (F_ACCBUFS - 1);
            unsigned par = (unsigned)(it0 >> F_ACCSHIFT) & 1u;
            int item_base = slot * F_CHUNK;
            const int item_step = (pass == 0 ? P.sample_stride : 1) * FN;
            const unsigned taddr0 = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(slot * F_CHUNK);
            int* const flag_p = &rs->flags[row];
            float thr = rs->thr[row];                       // the row's threshold only moves at the meetings: kept in a register in between

#pragma unroll 1
            for (int left = ntiles; left > 0; left--, item_base += item_step) {
                const int tile_end = item_base - slot * F_CHUNK + FN;
                mbar_wait(bar_accf + 8 * buf, par);
                tc_fence_after();
                unsigned v[32];
#if RMB_F_DBG
                if (!(P.dbg & 8))
#endif
                tmem_ld32(taddr0 + (unsigned)(buf * FN), v);
                tc_fence_before();                      // this warp's part of the accumulator is in registers
                __syncwarp();
                if (lane == 0) {
                    if (PAIR && rank != 0) mbar_arrive_cluster(mapa_shared(bar_acce + 8 * buf, 0u));    // the leader counts both CTAs' warps
                    else mbar_arrive(bar_acce + 8 * buf);
                }
                if (++buf == F_ACCBUFS) { buf = 0; par ^= 1u; }
#if RMB_F_DBG
                if (P.dbg & 1) continue;
#endif
                const bool has_train = t_nxt < tile_end;
                // NaN-propagating max of each group of 8 columns; one compare for the whole chunk, groups only when it passes
                float gm[4];
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const float m01 = max_nan(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1]));
                    const float m23 = max_nan(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3]));
                    const float m45 = max_nan(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5]));
                    const float m67 = max_nan(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7]));
                    gm[g] = max_nan(max_nan(m01, m23), max_nan(m45, m67));
                }
                if (!(max_nan(max_nan(gm[0], gm[1]), max_nan(gm[2], gm[3])) < thr)) {
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        if (!(gm[g] < thr)) {
                            F_STAT(2, 1);
                            const int before = mycnt;
                            mycnt = filter_append_group<RC>(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1]), __uint_as_float(v[8 * g + 2]),
                                                            __uint_as_float(v[8 * g + 3]), __uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5]),
                                                            __uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7]), thr, item_base + 8 * g, P.n,
                                                            P.tri, t_cur, has_train ? t_end : t_cur, cs, ci, mycnt, flag_p);
                            F_STAT(0, mycnt - before);
                            (void)before;
                        }
                    }
                }
                // move the train cursor past this tile: the id after t_nxt was prefetched into shared memory when the cursor last moved
                if (has_train) {
                    F_STAT(3, 1);
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    int nxt = rs->nxt2_train[slot][row];
                    t_cur++;
                    while (nxt < tile_end) { t_cur++; nxt = t_cur < t_end ? P.tri[t_cur] : INT_MAX; }    // several train items in one tile
                    t_nxt = nxt;
                    if (t_cur + 1 < t_end)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&rs->nxt2_train[slot][row])), "l"(P.tri + t_cur + 1) : "memory");
                    else rs->nxt2_train[slot][row] = INT_MAX;
                }
                if (--meet_in >= 0 && left > 1) continue;
#if RMB_F_DBG
                if (P.dbg & 4) continue;
#endif

                // ---- meeting point of the quarter's four warps: publish the region fills, cut rows that grew past their trigger
                rs->cnt4[slot][row] = mycnt;
                asm volatile("bar.sync %0, 128;" ::"r"(qbar) : "memory");
                int c0 = rs->cnt4[0][row], c1 = rs->cnt4[1][row], c2 = rs->cnt4[2][row], c3 = rs->cnt4[3][row];
                int tot = c0 + c1 + c2 + c3;
                const int grown = tot - tot_prev;                              // appends to the row since the previous meeting
                unsigned need = __ballot_sync(FULL, tot > rs->trig[row] && !(rs->flags[row] & 2));
                if (left == 1) need = 0;                                       // the pass ends with the exact cut below
                if (need) {                                                    // (the same mask in all four warps)
                    while (need) {
                        const int r = __ffs(need) - 1;
                        need &= need - 1;
                        if ((r & 3) != slot) continue;                         // the quarter's warps share the work
                        const int rr = q * 32 + r;
                        const size_t base = (size_t)(tile_u0 + rr) * C;
                        float tau_r;
                        const int kept = cut_regions<C, RMB_F_HIST_ROUNDS, false>(P.cand_approx + base, P.cand_item + base, rs->cnt4[0][rr], rs->cnt4[1][rr],
                                                                                  rs->cnt4[2][rr], rs->cnt4[3][rr], Kp, rs->slack[rr], lane, rs->hist[warp], &tau_r);
                        if (lane == 0) {
                            F_STAT(1, 1);
                            float thr_r = __fsub_rd(tau_r, rs->slack[rr]);
                            if (!(thr_r == thr_r)) thr_r = -CUDART_INF_F;       // non-finite bound: keep everything
                            thr_r = fmaxf(thr_r, rs->thr[rr]);                  // an earlier (valid) bound may be the sharper one
                            if (kept > 4 * (RC - 32)) { rs->flags[rr] |= 2; thr_r = CUDART_INF_F; }     // the slack band does not fit
                            rs->thr[rr] = thr_r;
                            rs->trig[rr] = kept + F_CUT_MARGIN;
#pragma unroll
                            for (int sgm = 0; sgm < 4; sgm++) rs->cnt4[sgm][rr] = (kept + 3 - sgm) >> 2;
                        }
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(qbar) : "memory");
                    c0 = rs->cnt4[0][row]; c1 = rs->cnt4[1][row]; c2 = rs->cnt4[2][row]; c3 = rs->cnt4[3][row];
                    tot = c0 + c1 + c2 + c3;
                    mycnt = slot == 0 ? c0 : (slot == 1 ? c1 : (slot == 2 ? c2 : c3));
                }
                thr = rs->thr[row];
                if (rs->flags[row] & 2) { mycnt = 0; thr = CUDART_INF_F; if (slot == 0) rs->thr[row] = CUDART_INF_F; }     // the row is closed: whatever it holds is void
                tot_prev = tot;
                // next meeting: far enough that barriers are rare, close enough that no region can fill up at twice the
                // append rate just seen (all of it into one region); a burst beyond that closes the row (flag 2)
                const int headroom = RC - max(max(c0, c1), max(c2, c3));
                int nxt_iv = grown > 0 ? (headroom * interval) / (2 * grown) : F_MAX_MEET;
                nxt_iv = __reduce_min_sync(FULL, nxt_iv);
                interval = nxt_iv < 1 ? 1 : (nxt_iv > F_MAX_MEET ? F_MAX_MEET : nxt_iv);
                meet_in = interval - 1;
            }
            // end of the pass (all four warps are past the last meeting): the exact Kp-th best approximate score of what each
            // row kept, the survivors packed at the head of the row's buffer
            for (int r = slot; r < 32; r += 4) {
                const int rr = q * 32 + r;
                if (rs->thr[rr] == CUDART_INF_F && !(rs->flags[rr] & 2)) continue;    // closed row (not ranked / settled in pass 1)
                const int c0 = rs->cnt4[0][rr], c1 = rs->cnt4[1][rr], c2 = rs->cnt4[2][rr], c3 = rs->cnt4[3][rr];
                const int nv = c0 + c1 + c2 + c3;
                float tau_r = -CUDART_INF_F;
                const bool have_k = nv >= Kp && !(rs->flags[rr] & 2);
                if (nv > 0 && !(rs->flags[rr] & 2)) {
                    const size_t base = (size_t)(tile_u0 + rr) * C;
                    const int kept = cut_regions<C, 7, true>(P.cand_approx + base, P.cand_item + base, c0, c1, c2, c3, have_k ? Kp : nv, rs->slack[rr], lane,
                                                             rs->hist[warp], &tau_r);
                    if (lane == 0 && pass > 0) rs->cnt[rr] = kept;      // last cut: the exact stage gets only what can still be in the top K
                }
                if (lane == 0) {
                    if (pass == 0) {
                        rs->guess[rr] = have_k ? tau_r : -CUDART_INF_F;                 // too small a sample / overflow: no guess
                        rs->flags[rr] &= ~2;
                    } else if (pass == 1 && !(rs->flags[rr] & 2)) {
                        const float g = rs->guess[rr];
                        const bool ok = (g == -CUDART_INF_F) || (have_k && tau_r >= g);  // >= K kept candidates reach the guess
                        if (!ok) { rs->flags[rr] |= 4; rs->retry = 1; }
                    }
                }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(qbar) : "memory");
        }
        it0 += ntiles;
        if (pass == 0) continue;         // the sample's result is consumed by the epilogue warps alone
        __syncthreads();
        unsigned again = rs->retry != 0 ? 1u : 0u;
        if (PAIR) {                                                      // both CTAs walk the same passes: retry if either needs it
            cluster_sync_all();
            again |= ld_shared_cluster_u32(mapa_shared(smem_u32(&rs->retry), rank ^ 1u)) != 0u ? 1u : 0u;
        }
        if (pass == 2 || !__any_sync(FULL, again != 0u)) break;          // (a vote: the compiler keeps the pass loop warp-uniform)
    }

    if (warp < F_EPI_WARPS && (warp >> 2) == 0) {
        const int row = (warp & 3) * 32 + lane, ul = tile_u0 + row;
        if (ul < P.mb) {
            const int fl = rs->flags[row];
            P.cand_count[ul] = (fl & 2) ? -1 : rs->cnt[row];
            if (fl & 2) atomicAdd(P.overflow, 1);
            if (fl & 1) atomicOr(&P.uflags[P.user0 + ul], 1);
            if ((fl & 4) && P.retries) atomicAdd(P.retries, 1);
        }
    }

    tc_fence_before();
    __syncthreads();
#if RMB_F_STATS
    if (tid < 4 && P.retries) atomicAdd(P.retries + 1 + tid, rs->stat[tid]);
#endif
    if (PAIR) cluster_sync_all();       // neither CTA leaves (or frees tensor memory) while the pair's MMAs or remote arrivals are in flight
    if (warp == F_EPI_WARPS + 1) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((unsigned)F_TMEM_COLS));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((unsigned)F_TMEM_COLS));
    }
}

// One warp per user: exact scores of the candidates the filter kept (sequential fma chain over the original
// factors, + bias: the chain of score_select_kernel / score_entries_kernel, so the values are bit-identical
// to the FMA path), NaN scores raise the user's NaN flag (hpp:195-197), best K (score descending, ties by
// ascending item id) left unordered at the head of cand_score / cand_item.
// MODE 1 / 2 (staged): the item-factor rows of 32 candidates at a time are fetched with cp.async into shared memory --
// all 32 rows in flight at once, no registers in between -- then every lane runs the chain of its own candidate from
// there.  MODE 2 moves 16 bytes per cp.async and reads the rows back with 16-byte loads (row stride p_pad + 16 bytes:
// conflict-free for 16-byte accesses); it needs factor rows that are 16-byte aligned and a multiple of 16 bytes long.
// MODE 1 moves single elements (row stride p_pad + 1).  MODE 0 reads the rows straight from global memory.
constexpr int EXACT_WARPS = 4;
__device__ __forceinline__ void cp_async_elem(float* dst, const float* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_elem(double* dst, const double* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// per warp: 32 rows of p_pad + 16 bytes (either staged mode fits) + the user's factors
inline size_t exact_topk_smem_bytes(int p_pad, size_t elem) { return (size_t)EXACT_WARPS * (32 * (size_t)(p_pad + 16 / elem) + p_pad) * elem; }

// Best K of the warp's scored candidates (key: order-preserving integer image of the score, 0 = empty slot; ties at the
// cut by ascending item id), written unordered to the head of the user's buffer.  Only the first EU of the E register
// slots per lane can be occupied.  Returns the number kept.
template <typename T, int E, int EU>
__device__ __forceinline__ int exact_select(const typename NumTraits<T>::key_t (&key)[E], const int (&it)[E], const int K, const int lane,
                                            T* cs, int* cio)
{
    typedef typename NumTraits<T>::key_t key_t;
    int nvalid = 0;
#pragma unroll
    for (int e = 0; e < EU; e++) nvalid += (key[e] != 0) ? 1 : 0;
    nvalid = __reduce_add_sync(FULL, nvalid);
    key_t t = 0;                  // keep key > t, and key == t with item id <= id_cut
    unsigned id_cut = 0x7fffffffu;
    if (nvalid > K) {
        for (int b = NumTraits<T>::KEYBITS - 1; b >= 0; b--) {
            const key_t cand = t | ((key_t)1 << b);
            int c = 0;
#pragma unroll
            for (int e = 0; e < EU; e++) c += (key[e] >= cand) ? 1 : 0;
            c = __reduce_add_sync(FULL, c);
            if (c >= K) t = cand;
        }
        int cgt = 0, cge = 0;
#pragma unroll
        for (int e = 0; e < EU; e++) { cgt += (key[e] > t) ? 1 : 0; cge += (key[e] >= t) ? 1 : 0; }
        cgt = __reduce_add_sync(FULL, cgt);
        cge = __reduce_add_sync(FULL, cge);
        if (cge > K) {
            const int need = K - cgt;
            unsigned x = 0;
            for (int b = 30; b >= 0; b--) {
                const unsigned cand = x | (1u << b);
                int c = 0;
#pragma unroll
                for (int e = 0; e < EU; e++) c += (key[e] == t && (unsigned)it[e] < cand) ? 1 : 0;
                c = __reduce_add_sync(FULL, c);
                if (c < need) x = cand;
            }
            id_cut = x;
        }
    }
    __syncwarp();
    int base = 0;
#pragma unroll
    for (int e = 0; e < EU; e++) {
        const bool keep = (key[e] != 0) && ((key[e] > t) || (key[e] == t && (unsigned)it[e] <= id_cut));
        const unsigned mask = __ballot_sync(FULL, keep);
        if (keep) {
            const int pos = base + __popc(mask & ((1u << lane) - 1u));
            cs[pos] = NumTraits<T>::from_orderable((u64)key[e]);
            cio[pos] = it[e];
        }
        base += __popc(mask);
    }
    return base;
}

template <typename T, int C, int MODE>
__global__ void __launch_bounds__(EXACT_WARPS * 32)
exact_topk_kernel(const float* __restrict__ cand_approx, T* __restrict__ cand_score, int* __restrict__ cand_item,
                  int* __restrict__ cand_count, const int mb, const int user0,
                  const T* __restrict__ At, const int p_pad, const int p,
                  const T* __restrict__ Brow, const size_t ldb, const T* __restrict__ bias,
                  int* __restrict__ uflags, const int K,
                  const int noise, const unsigned long long seed_user0, const size_t mt_off,
                  const int* __restrict__ trp, const int* __restrict__ tri, const int n)
{
    typedef typename NumTraits<T>::key_t key_t;
    constexpr int E = C / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ul = blockIdx.x * EXACT_WARPS + warp;
    if (ul >= mb) return;
    const int nv = cand_count[ul];
    if (nv <= 0) return;
    const T* __restrict__ a = At + (size_t)(ul / BM) * p_pad * BM + (ul % BM);      // + k * BM
    const int* ci = cand_item + (size_t)ul * C;
    constexpr bool STAGED = MODE != 0;
    constexpr int V = 16 / (int)sizeof(T);                                                         // elements per 16 bytes
    const int RS = p_pad + (MODE == 2 ? V : 1);                                                    // row stride of the staged rows
    T* rows = reinterpret_cast<T*>(smem_raw) + (size_t)warp * (32 * (size_t)(p_pad + V) + p_pad);   // [32][RS]
    T* a_sm = rows + 32 * (size_t)(p_pad + V);                                                     // [p_pad]
    if (STAGED) {
        for (int k = lane; k < p; k += 32) a_sm[k] = a[(size_t)k * BM];
        __syncwarp();
    }
    key_t key[E];
    int it[E];
    int nanflag = 0;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const int idx = e * 32 + lane;
        key[e] = 0;
        it[e] = INT_MAX;
        if (e * 32 < nv) {                       // warp-uniform
            const bool valid = idx < nv;
            const int item = valid ? ci[idx] : 0;
            if (valid) it[e] = item;
            T acc = (T)0;
            if (STAGED) {
                // candidate j of this round: its factor row, 32 lanes wide, straight into shared memory (cp.async: all
                // 32 rows are in flight at once, no registers in between)
                for (int j = 0; j < 32; j++) {
                    const int item_j = __shfl_sync(FULL, item, j);
                    const T* __restrict__ b = Brow + (size_t)item_j * ldb;
                    T* dst = rows + (size_t)j * RS;
                    if (MODE == 2) { for (int k = lane * V; k < p; k += 32 * V) cp_async_16(dst + k, b + k); }
                    else { for (int k = lane; k < p; k += 32) cp_async_elem(dst + k, b + k); }
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncwarp();
                const T* mine = rows + (size_t)lane * RS;
                if (MODE == 2) {
#pragma unroll 4
                    for (int k = 0; k < p; k += V) {
                        T av[V], bv[V];
                        lds_vec(a_sm + k, av);
                        lds_vec(mine + k, bv);
#pragma unroll
                        for (int x = 0; x < V; x++) acc = NumTraits<T>::fma(av[x], bv[x], acc);
                    }
                } else {
#pragma unroll 8
                    for (int k = 0; k < p; k++) acc = NumTraits<T>::fma(a_sm[k], mine[k], acc);
                }
                __syncwarp();
            } else if (valid) {
                const T* __restrict__ b = Brow + (size_t)item * ldb;
#pragma unroll 8
                for (int k = 0; k < p; k++) acc = NumTraits<T>::fma(a[(size_t)k * BM], b[k], acc);
            }
            if (valid) {
                if (bias != nullptr) acc += bias[item];
                if (acc != acc) nanflag = 1;
                else key[e] = NumTraits<T>::key(acc);
            }
        }
    }
    if (__any_sync(FULL, nanflag)) { if (lane == 0) atomicOr(&uflags[user0 + ul], 1); }
    if (noise) {
        // break_ties_with_noise (hpp:531-534, tie_noise.cuh): candidate number ix of the user -- candidates in ascending item
        // order, i.e. item id minus the train items before it -- gets draw ix of mt19937(seed + user).  Only scores the
        // noise can change need their draw (every double score; float scores below 2^-14 in magnitude).
        bool want = false;
        int pos[E];
        const int t0 = trp[user0 + ul], t1 = trp[user0 + ul + 1];
        {
            // validity rule of the noise branch (hpp:524-527): all candidates equal => NaN row.  The filter keeps whatever can
            // reach the K-th best score: if it kept fewer than all candidates some score is lower than the rest; if it kept
            // them all, they are all here.
            key_t kmax = 0, kmin = ~(key_t)0;
            int cntv = 0;
#pragma unroll
            for (int e = 0; e < E; e++)
                if (key[e] != 0) { kmax = key[e] > kmax ? key[e] : kmax; kmin = key[e] < kmin ? key[e] : kmin; cntv++; }
            cntv = __reduce_add_sync(FULL, cntv);
            // warp-wide extremes (64-bit keys for double: reduce through shuffles)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const key_t omax = __shfl_xor_sync(FULL, kmax, o), omin = __shfl_xor_sync(FULL, kmin, o);
                kmax = omax > kmax ? omax : kmax;
                kmin = omin < kmin ? omin : kmin;
            }
            if (cntv == n - (t1 - t0) && cntv > 0 && kmax == kmin && lane == 0) atomicOr(&uflags[user0 + ul], 4);
        }
#pragma unroll
        for (int e = 0; e < E; e++) {
            pos[e] = -1;
            if (key[e] != 0 && TieNoise<T>::can_change(NumTraits<T>::from_orderable((u64)key[e]))) {
                int lo = t0, hi = t1;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (tri[mid] < it[e]) lo = mid + 1; else hi = mid; }
                pos[e] = it[e] - (lo - t0);
                want = true;
            }
        }
        if (__any_sync(FULL, want)) {
            int maxpos = -1;
#pragma unroll
            for (int e = 0; e < E; e++) maxpos = max(maxpos, pos[e]);
            maxpos = __reduce_max_sync(FULL, maxpos);
            unsigned* x = reinterpret_cast<unsigned*>(smem_raw + mt_off) + warp * MT_N;
            mt_seed_warp(x, seed_user0 + (unsigned long long)ul, lane);
            const int nblocks = (int)(((long long)(maxpos + 1) * TieNoise<T>::WORDS + MT_N - 1) / MT_N);
            for (int blk = 0; blk < nblocks; blk++) {
                mt_twist_warp(x, lane);
#pragma unroll
                for (int e = 0; e < E; e++) {
                    if (pos[e] < 0) continue;
                    const long long w = (long long)pos[e] * TieNoise<T>::WORDS - (long long)blk * MT_N;     // first word of the draw, inside this block?
                    if (w >= 0 && w < MT_N) {
                        const unsigned w0 = mt_temper(x[w]);
                        const unsigned w1 = TieNoise<T>::WORDS == 2 ? mt_temper(x[w + 1]) : 0u;
                        const T noisy = NumTraits<T>::from_orderable((u64)key[e]) + TieNoise<T>::draw(w0, w1);
                        key[e] = NumTraits<T>::key(noisy);
                    }
                }
                __syncwarp();
            }
        }
    }
    T* cs = cand_score + (size_t)ul * C;
    int* cio = cand_item + (size_t)ul * C;
    int base;
    if (E > 4 && nv <= 128) base = exact_select<T, E, (E < 4 ? E : 4)>(key, it, K, lane, cs, cio);      // (the filter rarely keeps more: K' ~ 1.2 K)
    else if (E > 8 && nv <= 256) base = exact_select<T, E, (E < 8 ? E : 8)>(key, it, K, lane, cs, cio);
    else base = exact_select<T, E, E>(key, it, K, lane, cs, cio);
    if (lane == 0) cand_count[ul] = base;
}

}  // namespace rmb
