// prep.cuh -- layout and pre-pass kernels around the fused scoring kernel.
//
//   pack_tiles_kernel      row-major factors (the reference's layout, hpp:213-227) -> 128-wide k-major slabs
//   user_status_kernel     eligibility filter of /root/reference/src/recometrics.hpp:439-448, :479-486
//   score_entries_kernel   scores of the held-out (test) items, same FMA order as the tile kernel
//   sort_positives_kernel  per-user ascending order of those scores (rank by counting)
#pragma once
#include <cuda_fp16.h>
#include "score_select.cuh"

namespace rmb {

// Re-tile a row-major factor matrix src[rows][cols] (leading dimension ld_src) into W-row slabs,
// k-major inside a slab and zero padded:  dst[(r / W) * cols_pad + c][r % W] = src[r][c].
// One slab chunk of `kcount` factors is then a contiguous block of kcount*W elements (one TMA bulk copy).
template <typename T, int W>
__global__ void pack_tiles_kernel(const T* __restrict__ src, const size_t ld_src, const int rows, const int cols,
                                  T* __restrict__ dst, const int rows_pad, const int cols_pad)
{
    __shared__ T tile[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;      // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < rows && c < cols) ? src[(size_t)r * ld_src + c] : (T)0;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (c < cols_pad && r < rows_pad)
            dst[((size_t)(r / W) * cols_pad + c) * W + (r % W)] = tile[tx][i];
    }
}

// The same for a LIST of rows: slab row r of dst = src row base + rowmap[r] (r < rows; the rest is zero).  Used to re-run
// the users the tensor-core filter handed back on the FMA tiles (api.cu).
template <typename T, int W>
__global__ void pack_tiles_gather_kernel(const T* __restrict__ src, const size_t ld_src, const int* __restrict__ rowmap, const int rows,
                                         const int cols, T* __restrict__ dst, const int rows_pad, const int cols_pad)
{
    __shared__ T tile[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;      // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < rows && c < cols) ? src[(size_t)rowmap[r] * ld_src + c] : (T)0;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (c < cols_pad && r < rows_pad)
            dst[((size_t)(r / W) * cols_pad + c) * W + (r % W)] = tile[tx][i];
    }
}

// break_ties_with_noise + rank counting: mark (-1) the users of the batch with a held-out item that has another candidate within
// the noise's reach (near[e] counts the item itself as well): the noise decides one of their ranks.
__global__ void mark_noise_users_kernel(const int* __restrict__ tep, const unsigned* __restrict__ near, const int* __restrict__ ustatus,
                                        const int user0, const int nb, int* __restrict__ mark)
{
    const int ul = blockIdx.x * blockDim.x + threadIdx.x;
    if (ul >= nb) return;
    const int u = user0 + ul;
    int m = 0;
    if (ustatus[u] == 0)
        for (int e = tep[u]; e < tep[u + 1]; e++)
            if (near[e] >= 2u) { m = -1; break; }
    mark[ul] = m;
}

// Users of a batch the filter flagged (cand_count == -1), in ascending order: list[0 .. *count).  One block; the batch
// has at most a few hundred thousand users and flagged ones are rare.
__global__ void collect_flagged_kernel(const int* __restrict__ cand_count, const int nb, int* __restrict__ list, int* __restrict__ count)
{
    __shared__ int warp_tot[32];
    __shared__ int base_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int u0 = 0; u0 < nb; u0 += blockDim.x) {
        const int u = u0 + threadIdx.x;
        const bool f = u < nb && cand_count[u] < 0;
        const unsigned mask = __ballot_sync(FULL, f);
        if (lane == 0) warp_tot[warp] = __popc(mask);
        __syncthreads();
        int off = base_s;
        for (int w = 0; w < warp; w++) off += warp_tot[w];
        if (f) list[off + __popc(mask & ((1u << lane) - 1u))] = u;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < nw; w++) t += warp_tot[w]; base_s += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base_s;
}

template <typename T>
__global__ void pad_copy_kernel(const T* __restrict__ src, const int n, T* __restrict__ dst, const int n_pad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) dst[i] = i < n ? src[i] : (T)0;
}

struct StatusParams {
    const int* trp; const int* tep;
    int n, K, user_begin, user_end;
    int has_ndcg, has_rescue;      // rescue = roc || pr || ap || tap || rr  (hpp:485)
    int consider_cold_start, min_items_pool, min_pos_test;
    int* ustatus;                  // [m]
    int* uflags;                   // [m] cleared
    unsigned long long* umin;      // [m] or nullptr, set to ~0
};

// hpp:439-448 (status 1) and hpp:483-486 (status 2); 0 = the user gets ranked.
__global__ void user_status_kernel(const StatusParams P)
{
    const int u = P.user_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= P.user_end) return;
    const int ntrain = P.trp[u + 1] - P.trp[u];
    const int npos = P.tep[u + 1] - P.tep[u];
    int st = 0;
    if (npos <= 0 || ((ntrain + npos) >= P.n && !P.has_ndcg) || (P.n - ntrain) < P.min_items_pool ||
        (!P.consider_cold_start && ntrain == 0) || npos < P.min_pos_test)
        st = 1;
    else if ((P.n - ntrain) <= P.K && !P.has_rescue)
        st = 2;
    P.ustatus[u] = st;
    P.uflags[u] = 0;
    if (P.umin) P.umin[u] = ~0ull;
}

// One warp per user: score every held-out item of the user.  The accumulation is the sequential
// fma chain over k = 0..p_pad-1 (then + bias) that score_select_kernel performs for the same
// (user,item) pair, so both kernels produce bit-identical values.  At / Bt are the tiled slabs.
template <typename T, int W>      // W: width of the user-factor tiles (BM, or AUC_UM for the rank-counting kernel's copy)
__global__ void score_entries_kernel(const T* __restrict__ At, const T* __restrict__ Bt,
                                     const T* __restrict__ bias, const int p_pad, const int user0, const int mb,
                                     const int* __restrict__ tep, const int* __restrict__ tei,
                                     const int* __restrict__ ustatus, T* __restrict__ pos_raw)
{
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int ul = blockIdx.x * warps_per_block + (threadIdx.x >> 5); ul < mb; ul += gridDim.x * warps_per_block) {
        const int u = user0 + ul;
        if (ustatus[u] != 0) continue;
        const int e0 = tep[u], e1 = tep[u + 1];
        constexpr int BN = NumTraits<T>::BN;
        const T* a = At + (size_t)(ul / W) * p_pad * W + (ul % W);
        for (int e = e0 + lane; e < e1; e += 32) {
            const int item = tei[e];
            const T* b = Bt + (size_t)(item / BN) * p_pad * BN + (item % BN);
            T acc = (T)0;
            for (int k = 0; k < p_pad; k++)
                acc = NumTraits<T>::fma(a[(size_t)k * W], b[(size_t)k * BN], acc);
            if (bias != nullptr) acc += bias[item];
            pos_raw[e] = acc;
        }
    }
}

// One warp per user: rank-by-counting sort of the user's held-out scores, ascending; among equal scores the LARGER
// item id first, so that a walk from the top meets tied items in ascending item id -- the order of the top-K selection
// (ranks_before) and of the rank counts (score_select_kernel).
// pos_perm[sorted position] = entry offset inside the row, pos_item[sorted position] = its item id.
template <typename T>
__global__ void sort_positives_kernel(const int user0, const int mb, const int* __restrict__ tep, const int* __restrict__ tei,
                                      const int* __restrict__ ustatus, const T* __restrict__ pos_raw,
                                      T* __restrict__ pos_sorted, int* __restrict__ pos_perm, int* __restrict__ pos_item)
{
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int ul = blockIdx.x * warps_per_block + (threadIdx.x >> 5); ul < mb; ul += gridDim.x * warps_per_block) {
        const int u = user0 + ul;
        if (ustatus[u] != 0) continue;
        const int e0 = tep[u];
        const int npos = tep[u + 1] - e0;
        for (int e = lane; e < npos; e += 32) {
            const T v = pos_raw[e0 + e];
            int rank = 0;
            for (int j = 0; j < npos; j++) {
                const T w = pos_raw[e0 + j];
                rank += (w < v) || (w == v && j > e);          // (the row's item ids ascend with the entry offset)
            }
            pos_sorted[e0 + rank] = v;
            pos_perm[e0 + rank] = e;
            pos_item[e0 + rank] = tei[e0 + e];
        }
    }
}


// Power-of-two scale that brings a positive norm into [0.5, 1): 2^-e with nrm = f * 2^e (frexp).  Scaling by it is
// exact; 1 for zero / non-finite norms (such rows produce non-finite scores and are handled as NaN rows).
__device__ __forceinline__ float pow2_scale_for(const float nrm)
{
    if (!(nrm > 0.f) || nrm == CUDART_INF_F) return 1.f;
    int e;
    frexpf(nrm, &e);
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    return ldexpf(1.f, -e);
}

// fp16 operand image for the tensor-core filter (filter_select.cuh): 128-row tiles, [tile][k/8][row][8 halves],
// KB factors per row (multiple of 16).  Column `cols` holds extra[row] (the item bias) or, for the user side,
// 1.0 (ones_col) -- the bias travels as one more factor, as in recometrics/__init__.py:548-551; the rest is 0.
// Rows are scaled by a power of two so that every element is at most 1 in magnitude (fp16 keeps 11 significant bits
// from 6e-5 up): the user side by its own row norm (row_norm[r]), the item side by ONE scale for the whole matrix
// (*global_norm_bits = max_j ||b_j||), so that a user's approximate scores stay comparable across items.
// One thread per (row, 8-factor chunk): consecutive threads write consecutive 16-byte rows of a chunk.
template <typename T>
__global__ void pack_f16_kernel(const T* __restrict__ src, const size_t ld, const int rows, const int cols,
                                const T* __restrict__ extra, const int ones_col,
                                const float* __restrict__ row_norm, const unsigned* __restrict__ global_norm_bits,
                                __half* __restrict__ dst, const int rows_pad, const int KB, const int split_halves = 0)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int chunks = KB / 8;
    const long long total = (long long)rows_pad * chunks;
    if (idx >= total) return;
    const int tile = (int)(idx / ((long long)chunks * 128));
    const int rem = (int)(idx % ((long long)chunks * 128));
    const int c = rem / 128, rl = rem % 128;
    const int r = tile * 128 + rl;
    float scale = 1.f;
    if (r < rows) scale = pow2_scale_for(row_norm != nullptr ? row_norm[r] : __uint_as_float(*global_norm_bits));
    __align__(16) __half v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int k = c * 8 + e;
        float x = 0.f;
        if (r < rows) {
            if (k < cols) x = (float)src[(size_t)r * ld + k];
            else if (k == cols) x = extra != nullptr ? (float)extra[r] : (ones_col ? 1.f : 0.f);
        }
        v[e] = __float2half_rn(x * scale);
    }
    // split_halves (item matrix for the CTA-pair filter): a tile is stored as two 64-row halves, [tile][half][k/8][64][8]
    const size_t out = split_halves ? ((((size_t)tile * 2 + (size_t)(rl / 64)) * chunks + c) * 64 + (size_t)(rl % 64)) : (size_t)idx;
    *reinterpret_cast<uint4*>(dst + out * 8) = *reinterpret_cast<const uint4*>(v);
}

// Euclidean norm of every row (with the extra bias / 1.0 component), rounded up a little.  A block of 8 warps takes 32
// consecutive rows (4 per warp).  The row is scaled by a power of two taken from its largest element before it is
// squared, so tiny rows do not underflow to 0 and huge ones do not overflow: the norm is exactly 0 iff every element
// is 0, NaN iff an element is NaN, +inf iff an element is infinite (or the norm leaves the float range).
// Optional outputs: the maximum over all rows (float bits compare like unsigned ints for non-negative values; NaN
// wins) and the maximum over each block's 32 rows (chunk_max[blockIdx.x]: the error bound of the tensor-core filter
// is taken per 32-item chunk, filter_select.cuh).
constexpr int NORM_ROWS_PER_BLOCK = 32;
template <typename T>
__global__ void __launch_bounds__(256)
row_norm_kernel(const T* __restrict__ src, const size_t ld, const int rows, const int cols,
                const T* __restrict__ extra, const int ones_col,
                float* __restrict__ norm_out, unsigned* __restrict__ max_bits, float* __restrict__ chunk_max)
{
    __shared__ unsigned wmax[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned mybits = 0u;
    for (int i = 0; i < 4; i++) {
        const int r = blockIdx.x * NORM_ROWS_PER_BLOCK + warp * 4 + i;
        if (r >= rows) break;                                       // (warp-uniform)
        const T* __restrict__ row = src + (size_t)r * ld;
        const T e = extra != nullptr ? extra[r] : (ones_col ? (T)1 : (T)0);
        T am = (T)0;
        int has_nan = 0;
        for (int k = lane; k < cols; k += 32) { const T x = row[k]; has_nan |= (x != x); const T ax = x < 0 ? -x : x; am = ax > am ? ax : am; }
        if (lane == 0) { has_nan |= (e != e); const T ae = e < 0 ? -e : e; am = ae > am ? ae : am; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const T w = __shfl_xor_sync(FULL, am, o); am = w > am ? w : am; }
        has_nan = __any_sync(FULL, has_nan);
        float nrm;
        if (has_nan) nrm = CUDART_NAN_F;
        else if (am == (T)0) nrm = 0.f;
        else if (am > (T)3.0e38) nrm = CUDART_INF_F;                // infinite element (or beyond what a float norm can hold)
        else {
            int ex;
            frexp((double)am, &ex);                                 // am = f * 2^ex, f in [0.5, 1)
            const double s = ldexp(1.0, -ex);                       // (double: 2^-ex can leave the float range for denormal rows)
            float acc = 0.f;
            for (int k = lane; k < cols; k += 32) { const float x = (float)((double)row[k] * s); acc = fmaf(x, x, acc); }
            if (lane == 0) { const float x = (float)((double)e * s); acc = fmaf(x, x, acc); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
            const double d = ldexp((double)(sqrtf(acc) * 1.001f), ex);
            nrm = d > 3.0e38 ? CUDART_INF_F : (d < 1.2e-38 ? 1.2e-38f : __double2float_ru(d));
        }
        if (lane == 0 && norm_out) norm_out[r] = nrm;
        const unsigned b = __float_as_uint(nrm);
        mybits = b > mybits ? b : mybits;
    }
    if (lane == 0) wmax[warp] = mybits;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned b = 0u;
        for (int w = 0; w < 8; w++) b = wmax[w] > b ? wmax[w] : b;
        if (chunk_max) chunk_max[blockIdx.x] = __uint_as_float(b);
        if (max_bits) atomicMax(max_bits, b);
    }
}

// Register-resident FMA chains on every SM: measured FP32 / FP64 FMA peak (roofline denominator).
template <typename T>
__global__ void fma_peak_kernel(T* out, const int iters, const T seed)
{
    T a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = seed + (T)(threadIdx.x + i);
    const T x = (T)1.0000001, y = (T)1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 8; rep++)
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = NumTraits<T>::fma(a[i], x, y);
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    if (s == (T)123456789) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // keep the chains alive
}

// The same with packed fma.rn.f32x2 (FFMA2): 2 fp32 FMAs per instruction, the form the scoring kernel uses.
__global__ void fma2_peak_kernel(float* out, const int iters, const float seed)
{
    u64 a[16];
    const u64 x = pack2(1.0000001f, 1.0000001f), y = pack2(1e-9f, 1e-9f);
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = pack2(seed + (float)(threadIdx.x + i), seed - (float)i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++)
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(x), "l"(y));
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= a[i];
    if (s == 123456789ull) out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}

}  // namespace rmb
