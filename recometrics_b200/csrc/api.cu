// api.cu -- C-ABI of librecometrics_b200.so and the host orchestration of one call.
//
// Stands where /root/reference/src/recometrics_instantiated.cpp:43-143 stands in the reference: the
// non-template entry points its bindings call (calc_metrics_float / calc_metrics_double), here
// driving the GPU instead of an OpenMP loop.  The host prologue mirrors
// /root/reference/src/recometrics.hpp:387-393 (clamps) and :426 / :488 / :964 (SIGINT latch).
#include "../../include/recometrics_b200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "metrics.cuh"
#include "prep.cuh"
#include "score_select.cuh"
#include "filter_select.cuh"
#include "full_order.h"

namespace {

thread_local std::string g_err;
std::atomic<int> g_interrupt{0};
std::mutex g_call_mutex;   // one call at a time per process (the SIGINT latch is process-wide)
constexpr int MAX_DEVICES = 16;   // GPUs one call can be spread over

void set_err(const char* what, const char* detail = nullptr)
{
    g_err = what;
    if (detail) { g_err += ": "; g_err += detail; }
}

}  // namespace
namespace rmb {
void set_last_error(const char* what, const char* detail) { set_err(what, detail); }   // for the library's other translation units (split.cu)
}
namespace {

extern "C" void rmb200_sigint_handler(int) { g_interrupt.store(1); }

// /root/reference/src/recometrics.hpp:126-174 (SignalSwitcher): swap the SIGINT handler for the
// duration of the call, poll the latch between user batches, restore + re-raise afterwards.
struct SignalLatch {
    void (*old_handler)(int) = SIG_DFL;
    bool active = false;
    explicit SignalLatch(bool engage = true)
    {
        if (!engage) return;
        g_interrupt.store(0);
        old_handler = std::signal(SIGINT, rmb200_sigint_handler);
        active = (old_handler != SIG_ERR);
    }
    void restore()
    {
        if (active) { std::signal(SIGINT, old_handler); active = false; }
    }
    ~SignalLatch() { restore(); }
};

// Device workspace cache: a call needs a few GB of scratch (operand images, candidate buffers); cudaMalloc / cudaFree
// of those cost tens of milliseconds per call and synchronise the device, so released blocks are kept for the next
// call (rmb200_release_workspace() or RMB200_NO_POOL=1 gives them back).
struct PoolEntry { void* p; size_t bytes; int dev; bool in_use; };
std::vector<PoolEntry> g_pool;
std::mutex g_pool_mutex;

bool pool_enabled()
{
    static const bool on = []() { const char* e = std::getenv("RMB200_NO_POOL"); return !(e && e[0] == '1'); }();
    return on;
}

void pool_trim_locked(int dev_only)
{
    for (size_t i = 0; i < g_pool.size();) {
        if (!g_pool[i].in_use && (dev_only < 0 || g_pool[i].dev == dev_only)) {
            cudaFree(g_pool[i].p);
            g_pool[i] = g_pool.back();
            g_pool.pop_back();
        } else i++;
    }
}

cudaError_t pool_alloc(void** out, size_t bytes)
{
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    if (pool_enabled()) {
        int best = -1;
        for (size_t i = 0; i < g_pool.size(); i++) {
            const PoolEntry& e = g_pool[i];
            if (e.in_use || e.dev != dev || e.bytes < bytes || e.bytes > bytes + bytes / 4 + (1u << 20)) continue;
            if (best < 0 || e.bytes < g_pool[best].bytes) best = (int)i;
        }
        if (best >= 0) { g_pool[best].in_use = true; *out = g_pool[best].p; return cudaSuccess; }
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {           // give cached blocks back and try once more
        cudaGetLastError();
        pool_trim_locked(dev);
        e = cudaMalloc(out, bytes);
    }
    if (e == cudaSuccess && pool_enabled()) g_pool.push_back(PoolEntry{*out, bytes, dev, true});
    return e;
}

void pool_free(void* p)
{
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    for (auto& e : g_pool)
        if (e.p == p) { e.in_use = false; return; }
    cudaFree(p);
}

}  // namespace
namespace rmb {   // the same cache for the library's other translation units (split.cu)
cudaError_t workspace_alloc(void** out, size_t bytes) { return pool_alloc(out, bytes); }
void workspace_free(void* p) { pool_free(p); }
}
namespace {

struct DevBuf {
    void* p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) { pool_free(p); p = nullptr; } }
    cudaError_t alloc(size_t bytes)
    {
        release();
        if (bytes == 0) bytes = 16;
        return pool_alloc(&p, bytes);
    }
    template <typename U> U* as() const { return reinterpret_cast<U*>(p); }
};

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            set_err(#call, cudaGetErrorString(e_));                                      \
            cudaGetLastError();                                                          \
            return (e_ == cudaErrorMemoryAllocation) ? RMB200_ERR_OOM : RMB200_ERR_CUDA; \
        }                                                                                \
    } while (0)

inline int round_up(int x, int q) { return (x + q - 1) / q * q; }

struct PhaseTimer {
    cudaStream_t st;
    cudaEvent_t a = nullptr, b = nullptr;
    explicit PhaseTimer(cudaStream_t s) : st(s) { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~PhaseTimer() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    void start() { cudaEventRecord(a, st); }
    void stop(double& acc)
    {
        cudaEventRecord(b, st);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        acc += ms;
    }
};

__global__ void rebase_indptr_kernel(const int* __restrict__ src, const int lo, int* __restrict__ dst, const int cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) dst[i] = src[i] - lo;
}

// One call spread over several GPUs (rmb200_extra_t::devices / RMB200_DEVICES): every device runs run_call() on its own
// contiguous block of users from its own host thread.  The only thing the devices share is the item-factor matrix: device
// g uploads the g-th slice of its rows over its own PCIe link and fetches the other slices from its peers' memory
// (NVLink), instead of G uploads of the whole matrix.  The threads meet twice: slices uploaded / slices fetched.
struct MultiCtx {
    int G = 0;
    std::mutex mu;
    std::condition_variable cv;
    int arrived[2] = {0, 0};
    int departed = 0;
    bool failed = false;
    void* Bptr[MAX_DEVICES] = {};
    int Bdev[MAX_DEVICES] = {};
    // false: a peer failed (the caller gives up as well)
    bool meet(int phase)
    {
        std::unique_lock<std::mutex> lk(mu);
        arrived[phase]++;
        cv.notify_all();
        cv.wait(lk, [&] { return arrived[phase] + departed >= G; });
        return !failed;
    }
    void depart(bool fail)          // the device's thread is done (a thread that fails early stands in for its later meetings)
    {
        std::lock_guard<std::mutex> lk(mu);
        departed++;
        if (fail) failed = true;
        cv.notify_all();
    }
};

template <typename T>
struct CallArgs {
    MultiCtx* mc = nullptr; int mc_rank = 0;
    const T* A; size_t lda; const T* B; size_t ldb;
    int m, n, k;
    const int32_t *trp, *tri, *tep, *tei; const T* tev;
    int K, cumulative, noise;
    T* out[10];   // p tp r ap tap ndcg hit rr roc pr
    int consider_cold_start, min_items_pool, min_pos_test;
    uint64_t seed;
    const T* bias;
    const rmb200_extra_t* ex;
};

template <typename T, int C, bool AUC>
cudaError_t launch_score_select_inst(const rmb::ScoreSelectParams<T>& P, int n_rows, int slices, cudaStream_t st)
{
    auto kern = rmb::score_select_kernel<T, C, AUC>;
    const size_t smem = rmb::score_select_smem_bytes<T, AUC>();
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<dim3((n_rows + rmb::BM - 1) / rmb::BM, slices), AUC ? rmb::AUC_THREADS : rmb::NTHREADS, smem, st>>>(P);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_score_select(const rmb::ScoreSelectParams<T>& P, int C, bool auc, int n_rows, int slices, cudaStream_t st)
{
    if (C == 256) return auc ? launch_score_select_inst<T, 256, true>(P, n_rows, slices, st)
                             : launch_score_select_inst<T, 256, false>(P, n_rows, slices, st);
    if (C == 512) return auc ? launch_score_select_inst<T, 512, true>(P, n_rows, slices, st)
                             : launch_score_select_inst<T, 512, false>(P, n_rows, slices, st);
    return auc ? launch_score_select_inst<T, 1024, true>(P, n_rows, slices, st)
               : launch_score_select_inst<T, 1024, false>(P, n_rows, slices, st);
}

template <int C>
cudaError_t launch_filter_inst(const rmb::FilterParams& P, int n_user_tiles, cudaStream_t st)
{
    const size_t smem = rmb::filter_smem_bytes(P.KB, P.stages);
    auto kern = rmb::filter_select_kernel<C>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<n_user_tiles, rmb::F_THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

inline cudaError_t launch_filter_select(const rmb::FilterParams& P, int C, int n_user_tiles, cudaStream_t st)
{
    if (C == 256) return launch_filter_inst<256>(P, n_user_tiles, st);
    if (C == 512) return launch_filter_inst<512>(P, n_user_tiles, st);
    return launch_filter_inst<1024>(P, n_user_tiles, st);
}

struct NoiseArgs { int on; unsigned long long seed_user0; const int* trp; const int* tri; int n; };

template <typename T, int C>
cudaError_t launch_exact_topk_inst(const uint2* capx, T* cs, int* ci, int* cc, int nb, int user0, const T* At, int p_pad, int p,
                                   const T* Brow, size_t ldb, const T* bias, int* uflags, int K, const NoiseArgs& nz,
                                   const rmb::FilterErrStat& es, cudaStream_t st)
{
    const int blocks = (nb + rmb::EXACT_WARPS - 1) / rmb::EXACT_WARPS;
    const size_t staging = (rmb::exact_topk_smem_bytes(p_pad, sizeof(T)) + 15) & ~size_t(15);
    const size_t mt_bytes = nz.on ? (size_t)rmb::EXACT_WARPS * rmb::MT_N * sizeof(unsigned) : 0;   // generator states (tie_noise.cuh)
    if (staging + mt_bytes <= 100 * 1024) {       // two blocks per SM
        // 16-byte copies when every factor row starts on a 16-byte boundary and is a multiple of 16 bytes long
        constexpr int V = 16 / (int)sizeof(T);
        bool vec = (p % V == 0) && (ldb % V == 0) && ((reinterpret_cast<uintptr_t>(Brow) & 15) == 0);
        if (const char* env = std::getenv("RMB200_EXACT_VEC")) vec = vec && std::atoi(env) != 0;      // developer: element-wise staging
        auto kern = vec ? rmb::exact_topk_kernel<T, C, 2> : rmb::exact_topk_kernel<T, C, 1>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(staging + mt_bytes));
        if (e != cudaSuccess) return e;
        kern<<<blocks, rmb::EXACT_WARPS * 32, staging + mt_bytes, st>>>(capx, cs, ci, cc, nb, user0, At, p_pad, p, Brow, ldb, bias, uflags, K,
                                                                        nz.on, nz.seed_user0, staging, nz.trp, nz.tri, nz.n, es);
    } else {
        rmb::exact_topk_kernel<T, C, 0><<<blocks, rmb::EXACT_WARPS * 32, mt_bytes, st>>>(capx, cs, ci, cc, nb, user0, At, p_pad, p, Brow, ldb, bias, uflags, K,
                                                                                             nz.on, nz.seed_user0, (size_t)0, nz.trp, nz.tri, nz.n, es);
    }
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_exact_topk(const uint2* capx, T* cs, int* ci, int* cc, int C, int nb, int user0, const T* At, int p_pad, int p,
                              const T* Brow, size_t ldb, const T* bias, int* uflags, int K, const NoiseArgs& nz,
                              const rmb::FilterErrStat& es, cudaStream_t st)
{
    if (C == 256) return launch_exact_topk_inst<T, 256>(capx, cs, ci, cc, nb, user0, At, p_pad, p, Brow, ldb, bias, uflags, K, nz, es, st);
    if (C == 512) return launch_exact_topk_inst<T, 512>(capx, cs, ci, cc, nb, user0, At, p_pad, p, Brow, ldb, bias, uflags, K, nz, es, st);
    return launch_exact_topk_inst<T, 1024>(capx, cs, ci, cc, nb, user0, At, p_pad, p, Brow, ldb, bias, uflags, K, nz, es, st);
}

// order the <= K survivors of every user (warp per user, bitonic network sized to K)
template <typename T>
cudaError_t launch_rank_topk(T* cs, int* ci, const int* cc, int C, int nb, int K, cudaStream_t st)
{
    const int blocks = (nb + 7) / 8;
    if (K <= 32) rmb::rank_topk_kernel<T, 1><<<blocks, 256, 0, st>>>(cs, ci, cc, C, nb);
    else if (K <= 64) rmb::rank_topk_kernel<T, 2><<<blocks, 256, 0, st>>>(cs, ci, cc, C, nb);
    else if (K <= 128) rmb::rank_topk_kernel<T, 4><<<blocks, 256, 0, st>>>(cs, ci, cc, C, nb);
    else if (K <= 256) rmb::rank_topk_kernel<T, 8><<<blocks, 256, 0, st>>>(cs, ci, cc, C, nb);
    else rmb::rank_topk_kernel<T, 16><<<blocks, 256, 0, st>>>(cs, ci, cc, C, nb);
    return cudaGetLastError();
}

// Copy a row-major host/device matrix slab [rows][cols] (leading dimension ld) into a compact
// device buffer when it lives on the host; on-device inputs are used in place.

// ---------------------------------------------------------------- host -> device uploads
// The reference's callers hand over ordinary (pageable) numpy / R memory (wrapper.pyx:208-224).  cudaMemcpy from
// pageable memory goes through the driver's single staging buffer at ~11 GB/s (measured: 62 ms for the 0.67 GB of a
// 151K-user cfg4 call); a PCIe 5 x16 link takes 55 GB/s from pinned memory.  Large pageable sources therefore go through
// our own pipeline: UP_THREADS host threads each copy their share of the rows into pinned bounce buffers (two per
// thread, UP_CHUNK bytes) and queue the DMA from there on their own stream -- the memcpy of one chunk overlaps the DMA
// of the previous one and the threads overlap each other.  Pinned (or registered) sources and small copies take the
// plain cudaMemcpyAsync path.  RMB200_UPLOAD_THREADS=0 switches the pipeline off.
constexpr size_t UP_CHUNK = 8u << 20;
constexpr int UP_MAX_THREADS = 16;
struct UploadLane {
    cudaStream_t st = nullptr;
    void* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};      // recorded behind the DMA that reads buf[i]
    bool recorded[2] = {false, false};           // ... at least once: the buffer may still be in flight from an earlier upload
};
struct UploadPool {
    int dev = -1;
    UploadLane lane[UP_MAX_THREADS];
    cudaEvent_t done[UP_MAX_THREADS] = {};
    cudaEvent_t start = nullptr;
    int lanes = 0;
    bool ok = false;
};
UploadPool g_ups[MAX_DEVICES];   // one per device: calls are serialised by the library's mutex, and inside a multi-GPU call
                                 // every device has its own host thread
thread_local int g_upload_share = 1;   // a multi-GPU call divides the host's copy threads among its devices

int upload_threads()
{
    int v = (int)std::thread::hardware_concurrency() / 2;      // measured on a 16-core box: 60 ms plain, 32 ms with 4 threads, 25 ms with 8
    if (v < 1) v = 1;
    if (const char* e = std::getenv("RMB200_UPLOAD_THREADS")) v = std::atoi(e);
    if (v > 1 && g_upload_share > 1) { v /= g_upload_share; if (v < 1) v = 1; }
    return v < 0 ? 0 : (v > UP_MAX_THREADS ? UP_MAX_THREADS : v);
}

void upload_pool_release(UploadPool& g_up)
{
    if (g_up.dev < 0) return;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(g_up.dev);
    for (int t = 0; t < UP_MAX_THREADS; t++) {
        UploadLane& L = g_up.lane[t];
        for (int b = 0; b < 2; b++) { if (L.buf[b]) cudaFreeHost(L.buf[b]); if (L.ev[b]) cudaEventDestroy(L.ev[b]); L.buf[b] = nullptr; L.ev[b] = nullptr; L.recorded[b] = false; }
        if (L.st) cudaStreamDestroy(L.st);
        L.st = nullptr;
        if (g_up.done[t]) cudaEventDestroy(g_up.done[t]);
        g_up.done[t] = nullptr;
    }
    if (g_up.start) cudaEventDestroy(g_up.start);
    g_up.start = nullptr;
    g_up.dev = -1; g_up.ok = false; g_up.lanes = 0;
    cudaSetDevice(cur);
}

void upload_pool_release_all()
{
    for (int d = 0; d < MAX_DEVICES; d++) upload_pool_release(g_ups[d]);
}

bool upload_pool_ready(int dev, int nthreads)
{
    if (dev < 0 || dev >= MAX_DEVICES) return false;
    UploadPool& g_up = g_ups[dev];
    if (g_up.ok && g_up.dev == dev && g_up.lanes >= nthreads) return true;
    upload_pool_release(g_up);
    g_up.dev = dev;
    g_up.lanes = nthreads;
    for (int t = 0; t < nthreads; t++) {
        UploadLane& L = g_up.lane[t];
        if (cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); upload_pool_release(g_up); return false; }
        for (int b = 0; b < 2; b++) {
            if (cudaHostAlloc(&L.buf[b], UP_CHUNK, cudaHostAllocDefault) != cudaSuccess ||
                cudaEventCreateWithFlags(&L.ev[b], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); upload_pool_release(g_up); return false; }
        }
        if (cudaEventCreateWithFlags(&g_up.done[t], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); upload_pool_release(g_up); return false; }
    }
    if (cudaEventCreateWithFlags(&g_up.start, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); upload_pool_release(g_up); return false; }
    g_up.ok = true;
    return true;
}

bool host_pointer_is_pageable(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

// rows x row_bytes from host memory (row pitch src_pitch) to a packed device buffer, ordered on stream `st`
cudaError_t upload_rows(void* dst, const void* src, size_t src_pitch, size_t row_bytes, size_t rows, cudaStream_t st)
{
    const size_t total = row_bytes * rows;
    const int nt = upload_threads();
    int dev = 0;
    cudaGetDevice(&dev);
    // a contiguous source is a byte stream (chunks of UP_CHUNK bytes); a pitched one is cut at row boundaries
    const bool contiguous = (src_pitch == row_bytes) || rows == 1;
    size_t min_bytes = 32u << 20;
    if (const char* env = std::getenv("RMB200_UPLOAD_MIN_BYTES")) min_bytes = (size_t)std::atoll(env);      // developer / tests
    if (total < min_bytes || total == 0 || nt == 0 || (!contiguous && row_bytes > UP_CHUNK) || !host_pointer_is_pageable(src) ||
        !upload_pool_ready(dev, nt)) {
        if (src_pitch == row_bytes) return cudaMemcpyAsync(dst, src, total, cudaMemcpyHostToDevice, st);
        return cudaMemcpy2DAsync(dst, row_bytes, src, src_pitch, row_bytes, rows, cudaMemcpyHostToDevice, st);
    }
    UploadPool& g_up = g_ups[dev];
    // what is already queued on `st` (e.g. kernels still reading dst's previous contents) comes first
    cudaEvent_t start = g_up.start;
    cudaError_t e = cudaEventRecord(start, st);
    if (e != cudaSuccess) return e;
    const size_t unit = contiguous ? 1 : row_bytes;                              // bytes per indivisible piece
    const size_t units = contiguous ? total : rows;
    const size_t rows_per_chunk = contiguous ? UP_CHUNK : UP_CHUNK / row_bytes;   // pieces per chunk
    const size_t nchunks = (units + rows_per_chunk - 1) / rows_per_chunk;
    cudaError_t errs[UP_MAX_THREADS];
    std::vector<std::thread> workers;
    for (int t = 0; t < nt; t++) {
        errs[t] = cudaSuccess;
        workers.emplace_back([&, t]() {
            cudaError_t er = cudaSetDevice(dev);
            UploadLane& L = g_up.lane[t];
            if (er == cudaSuccess) er = cudaStreamWaitEvent(L.st, start, 0);
            int b = 0;
            for (size_t c = (size_t)t; c < nchunks && er == cudaSuccess; c += (size_t)nt, b ^= 1) {
                const size_t r0 = c * rows_per_chunk, nr = (units - r0) < rows_per_chunk ? (units - r0) : rows_per_chunk;
                if (L.recorded[b]) er = cudaEventSynchronize(L.ev[b]);     // the DMA that last read this bounce buffer (this upload's or an earlier one's)
                if (er != cudaSuccess) break;
                const unsigned char* sp = static_cast<const unsigned char*>(src) + r0 * (contiguous ? 1 : src_pitch);
                if (contiguous) std::memcpy(L.buf[b], sp, nr);
                else for (size_t r = 0; r < nr; r++) std::memcpy(static_cast<unsigned char*>(L.buf[b]) + r * row_bytes, sp + r * src_pitch, row_bytes);
                er = cudaMemcpyAsync(static_cast<unsigned char*>(dst) + r0 * unit, L.buf[b], nr * unit, cudaMemcpyHostToDevice, L.st);
                if (er == cudaSuccess) { er = cudaEventRecord(L.ev[b], L.st); L.recorded[b] = true; }
            }
            errs[t] = er;
        });
    }
    for (auto& w : workers) w.join();
    for (int t = 0; t < nt; t++) if (errs[t] != cudaSuccess) return errs[t];
    // `st` continues once every lane's copies have landed (the bounce buffers are reused only after their own events)
    for (int t = 0; t < nt; t++) {
        e = cudaEventRecord(g_up.done[t], g_up.lane[t].st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, g_up.done[t], 0);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// The mirror image of upload_rows for one contiguous block: device memory to ordinary (pageable) host memory through the
// lanes' pinned bounce buffers, every lane alternating between the DMA into one buffer and the memcpy out of the other.  The
// driver's own path stages through one buffer and copies out on one thread (12 GB/s measured); this reaches what the host's
// memory system gives several threads.  Work queued on `st` before the call is waited for; returns when the bytes are in dst.
cudaError_t download_bytes(void* dst, const void* src, size_t total, cudaStream_t st)
{
    const int nt = upload_threads();
    int dev = 0;
    cudaGetDevice(&dev);
    size_t min_bytes = 32u << 20;
    if (const char* env = std::getenv("RMB200_UPLOAD_MIN_BYTES")) min_bytes = (size_t)std::atoll(env);      // developer / tests
    if (total < min_bytes || nt == 0 || !host_pointer_is_pageable(dst) || !upload_pool_ready(dev, nt)) {
        const cudaError_t e = cudaMemcpyAsync(dst, src, total, cudaMemcpyDeviceToHost, st);
        return e != cudaSuccess ? e : cudaStreamSynchronize(st);
    }
    UploadPool& g_up = g_ups[dev];
    cudaEvent_t start = g_up.start;
    cudaError_t e = cudaEventRecord(start, st);
    if (e != cudaSuccess) return e;
    const size_t nchunks = (total + UP_CHUNK - 1) / UP_CHUNK;
    cudaError_t errs[UP_MAX_THREADS];
    std::vector<std::thread> workers;
    for (int t = 0; t < nt; t++) {
        errs[t] = cudaSuccess;
        workers.emplace_back([&, t]() {
            cudaError_t er = cudaSetDevice(dev);
            UploadLane& L = g_up.lane[t];
            if (er == cudaSuccess) er = cudaStreamWaitEvent(L.st, start, 0);
            // (an earlier UPLOAD may still be reading the bounce buffers: its events come first)
            for (int b = 0; b < 2 && er == cudaSuccess; b++) if (L.recorded[b]) er = cudaEventSynchronize(L.ev[b]);
            auto span = [&](size_t c, size_t* off) { *off = c * UP_CHUNK; return (total - *off) < UP_CHUNK ? (total - *off) : UP_CHUNK; };
            auto fetch = [&](size_t c, int b) {
                size_t off; const size_t nb = span(c, &off);
                cudaError_t x = cudaMemcpyAsync(L.buf[b], static_cast<const unsigned char*>(src) + off, nb, cudaMemcpyDeviceToHost, L.st);
                if (x == cudaSuccess) { x = cudaEventRecord(L.ev[b], L.st); L.recorded[b] = true; }
                return x;
            };
            int b = 0;
            if (er == cudaSuccess && (size_t)t < nchunks) er = fetch((size_t)t, 0);
            for (size_t c = (size_t)t; c < nchunks && er == cudaSuccess; c += (size_t)nt, b ^= 1) {
                if (c + (size_t)nt < nchunks) er = fetch(c + (size_t)nt, b ^ 1);       // the next chunk's DMA runs under this chunk's memcpy
                if (er == cudaSuccess) er = cudaEventSynchronize(L.ev[b]);
                if (er != cudaSuccess) break;
                size_t off; const size_t nb = span(c, &off);
                std::memcpy(static_cast<unsigned char*>(dst) + off, L.buf[b], nb);
            }
            errs[t] = er;
        });
    }
    for (auto& w : workers) w.join();
    for (int t = 0; t < nt; t++) if (errs[t] != cudaSuccess) return errs[t];
    return cudaSuccess;
}

}  // namespace
namespace rmb {   // the pageable-memory copy pipelines for the library's other translation units (split.cu).  The lanes belong
                  // to whoever holds the call mutex: a split that runs beside an evaluation call takes the driver's plain path.
cudaError_t upload_bytes_pageable(void* dst, const void* src, size_t bytes, cudaStream_t st, int share)
{
    std::unique_lock<std::mutex> lk(g_call_mutex, std::try_to_lock);
    if (!lk.owns_lock()) return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess;
    g_upload_share = share > 1 ? share : 1;
    const cudaError_t e = bytes ? upload_rows(dst, src, bytes, bytes, 1, st) : cudaSuccess;
    g_upload_share = 1;
    return e;
}
cudaError_t download_bytes_pageable(void* dst, const void* src, size_t bytes, cudaStream_t st, int share)
{
    if (!bytes) return cudaSuccess;
    std::unique_lock<std::mutex> lk(g_call_mutex, std::try_to_lock);
    if (!lk.owns_lock()) {
        const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
        return e != cudaSuccess ? e : cudaStreamSynchronize(st);
    }
    g_upload_share = share > 1 ? share : 1;
    const cudaError_t e = download_bytes(dst, src, bytes, st);
    g_upload_share = 1;
    return e;
}
}
namespace {

// The user factors are uploaded one user batch ahead, on a stream of their own, into two alternating staging buffers:
// the upload of batch b+1 runs while the scoring kernel of batch b does.  ready[s]: the upload into buffer s has landed;
// freed[s]: the kernels that read buffer s have run.
struct PrefetchStream {
    cudaStream_t s = nullptr;
    cudaEvent_t ready[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
    cudaEvent_t test_rows = nullptr;                 // the held-out CSR arrays (read first by the metrics kernel) have landed
    bool freed_valid[2] = {false, false}, test_rows_pending = false;
    cudaError_t init()
    {
        cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        for (int i = 0; i < 2 && e == cudaSuccess; i++) {
            e = cudaEventCreateWithFlags(&ready[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&freed[i], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&test_rows, cudaEventDisableTiming);
        return e;
    }
    ~PrefetchStream()
    {
        if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }     // nothing may still be writing the staging buffers
        for (int i = 0; i < 2; i++) { if (ready[i]) cudaEventDestroy(ready[i]); if (freed[i]) cudaEventDestroy(freed[i]); }
        if (test_rows) cudaEventDestroy(test_rows);
    }
};

// The metric rows of a user batch go back to the host on a stream of their own, out of two alternating sets of staging
// buffers: the copy of batch b runs while batch b+1 is scored (cumulative rows are m x K values per metric: 400 MB at cfg5).
// ready[s]: the kernels writing set s have run; copied[s]: its copies have landed (the set may be written again).
struct ResultStream {
    cudaStream_t s = nullptr;
    cudaEvent_t ready[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    bool copied_valid[2] = {false, false};
    cudaError_t init()
    {
        cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        for (int i = 0; i < 2 && e == cudaSuccess; i++) {
            e = cudaEventCreateWithFlags(&ready[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming);
        }
        return e;
    }
    ~ResultStream()
    {
        if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }       // nothing may still be reading the staging buffers
        for (int i = 0; i < 2; i++) { if (ready[i]) cudaEventDestroy(ready[i]); if (copied[i]) cudaEventDestroy(copied[i]); }
    }
};

template <typename T>
int stage_rows(const T* src, size_t ld, int rows, int cols, bool on_dev, DevBuf& staging,
               const T** dev_src, size_t* dev_ld, cudaStream_t st, rmb200_timing_t& tm)
{
    if (on_dev) { *dev_src = src; *dev_ld = ld; return RMB200_OK; }
    CK(upload_rows(staging.p, src, ld * sizeof(T), (size_t)cols * sizeof(T), (size_t)rows, st));
    tm.h2d_bytes += (int64_t)rows * cols * (int64_t)sizeof(T);
    *dev_src = staging.as<T>();
    *dev_ld = (size_t)cols;
    return RMB200_OK;
}

template <typename T>
int run_call(const CallArgs<T>& a)
{
    using namespace rmb;
    const auto t_begin = std::chrono::steady_clock::now();
    const rmb200_extra_t* ex = a.ex;
    const bool on_dev = ex && ex->inputs_on_device;
    rmb200_timing_t tm;
    std::memset(&tm, 0, sizeof(tm));

    // ---- argument checks (shape errors are the binding's job in the reference,
    //      recometrics/wrapper.pyx:261-268; here they become RMB200_ERR_BAD_ARG) ----
    if (a.m <= 0 || a.n <= 0 || a.k <= 0 || a.K <= 0) { set_err("bad argument", "m, n, k and k_metrics must be positive"); return RMB200_ERR_BAD_ARG; }
    if (!a.A || !a.B || !a.trp || !a.tep) { set_err("bad argument", "A, B, Xtrain_csr_p and Xtest_csr_p are required"); return RMB200_ERR_BAD_ARG; }
    if (a.lda < (size_t)a.k || a.ldb < (size_t)a.k) { set_err("bad argument", "lda/ldb smaller than k"); return RMB200_ERR_BAD_ARG; }
    if (a.out[5] && !a.tev) { set_err("bad argument", "NDCG requested but Xtest_csr (values) is NULL"); return RMB200_ERR_BAD_ARG; }
    if (ex && ex->struct_size != (int32_t)sizeof(rmb200_extra_t)) { set_err("bad argument", "rmb200_extra_t::struct_size mismatch"); return RMB200_ERR_BAD_ARG; }
    bool any_out = false;
    for (int q = 0; q < 10; q++) any_out |= (a.out[q] != nullptr);
    const bool want_means = ex && (ex->metric_means || ex->metric_counts);
    const bool want_extras = ex && (ex->topk_items || ex->topk_scores || ex->pos_rank || ex->status);
    const bool want_filter_stats = (ex && ex->filter_stats) || std::getenv("RMB200_FILTER_STATS") != nullptr;
    if (!any_out && !want_extras && !want_means) return RMB200_OK;

    int ub = 0, ue = a.m;
    if (ex && ex->user_end < 0) return RMB200_OK;                        // explicit empty block
    if (ex && (ex->user_begin != 0 || ex->user_end != 0)) { ub = ex->user_begin; ue = ex->user_end; }
    if (ub < 0 || ue > a.m || ub > ue) { set_err("bad argument", "user range outside [0, m]"); return RMB200_ERR_BAD_ARG; }
    const int mr = ue - ub;   // users of this call ("shard"); device-side rows are shard-local
    if (mr == 0) return RMB200_OK;

    // ---- device ----
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        set_err("no CUDA device", "librecometrics_b200 has no CPU fallback");
        return RMB200_ERR_NO_DEVICE;
    }
    int dev = -1;
    if (ex && ex->device >= 0) dev = ex->device;
    else if (const char* env = std::getenv("RMB200_DEVICE")) dev = std::atoi(env);
    if (dev < 0) CK(cudaGetDevice(&dev));
    if (dev >= ndev) { set_err("bad argument", "device ordinal out of range"); return RMB200_ERR_BAD_ARG; }
    CK(cudaSetDevice(dev));

    const bool nested = a.mc != nullptr;                   // one device's share of a multi-GPU call: run_multi() holds the mutex and the latch
    std::unique_lock<std::mutex> lock(g_call_mutex, std::defer_lock);
    if (!nested) lock.lock();
    SignalLatch latch(!nested);
    cudaStream_t st = nullptr;   // legacy default stream: ordered after the caller's default-stream work
    PhaseTimer pt(st), pk(st);   // phases; the dominant scoring kernel alone

    // hpp:391-393
    int mip = a.min_items_pool < a.K ? a.K : a.min_items_pool;
    if (mip < 2) mip = 2;
    int mpt = a.min_pos_test;
    if (!(ex && ex->strict_min_pos_test)) mpt = mpt < 1 ? mpt : 1;   // std::min(min_pos_test, 1): quirk Q1

    // what a "NaN" written into the outputs looks like: a quiet NaN, or the caller's bit pattern (R: NA_REAL, hpp:75-80)
    T nan_value = std::numeric_limits<T>::quiet_NaN();
    if (ex && ex->has_nan_bits) {
        if (sizeof(T) == 8) { const uint64_t b = ex->nan_bits; std::memcpy(&nan_value, &b, 8); }
        else { const uint32_t b = (uint32_t)ex->nan_bits; std::memcpy(&nan_value, &b, 4); }
        if (nan_value == nan_value) { set_err("bad argument", "rmb200_extra_t::nan_bits is not a NaN"); return RMB200_ERR_BAD_ARG; }
    }

    const int K = a.K;
    // candidate buffer: K kept + 128 appended per tile + head-room between two cuts (the tensor-core filter keeps
    // the slack band below the K-th best as well: about 2.5 K entries on Gaussian scores, so it gets more room)
    int C = (K <= 64) ? 256 : (K <= 192 ? 512 : 1024);
    const bool want_roc = a.out[8] != nullptr, want_pr = a.out[9] != nullptr;
    const bool count_ranks = want_roc || want_pr || (ex && ex->pos_rank);
    const int p_pad = round_up(a.k, KPAD);
    constexpr int BN = NumTraits<T>::BN;
    const int n_pad = round_up(a.n, BN);
    const size_t rs = a.cumulative ? (size_t)K : 1;

    // ---- which scoring path: tensor-core filter + exact re-scoring (top-K only), or FMA tiles ----
    int path_req = ex ? ex->scoring_path : 0;                       // 0 auto, 1 fma, 2 tensor
    if (path_req == 0) if (const char* env = std::getenv("RMB200_PATH")) {
        if (!std::strcmp(env, "fma")) path_req = 1; else if (!std::strcmp(env, "tensor")) path_req = 2; else if (!std::strcmp(env, "full")) path_req = 3;
    }
    // k_metrics beyond what the selection kernels' candidate buffers hold: the full-order path (full_order.cu)
    bool use_full = path_req == 3 || K > RMB200_MAX_K;
    if (use_full && ex && (ex->scoring_path == 1 || ex->scoring_path == 2)) { set_err("unsupported", "scoring_path fma / tensor need k_metrics <= RMB200_MAX_K (384)"); return RMB200_ERR_UNSUPPORTED; }
    if (use_full) path_req = 3;                                      // (an RMB200_PATH preference does not apply)
    const int KB = round_up(a.k + (a.bias ? 1 : 0), 16);            // fp16 factors per row (bias = one more factor)
    int f_stages = 0;
    {
        // A tile + ring stages + barriers + row state within the 227 KB a CTA can have
        const long long tile = (long long)KB * 256, budget = 227 * 1024 - (long long)filter_smem_fixed_bytes();
        f_stages = (int)(budget / tile) - 1;
        if (f_stages > F_MAX_STAGES) f_stages = F_MAX_STAGES;
        if (const char* env = std::getenv("RMB200_STAGES")) { const int v = std::atoi(env); if (v >= 2 && v < f_stages) f_stages = v; }   // developer: shallower ring
    }
    // With rank counting (ROC/PR-AUC) every score has to be exact: the FMA tiles do everything.  Exception: with the
    // reference's tie-breaking noise the ranked top-K comes from the tensor path as well (its exact stage is where the
    // noise is applied, tie_noise.cuh), after the FMA pass has counted the ranks.
    const bool tensor_shape_ok = f_stages >= 2 && K <= 256 && !use_full;
    const bool tensor_ok = tensor_shape_ok && (!count_ranks || a.noise);
    if (path_req == 2 && !tensor_ok) { set_err("unsupported", "scoring_path=tensor needs no ROC/PR-AUC (rank counting), k <= ~400 and k_metrics <= 256"); return RMB200_ERR_UNSUPPORTED; }
    const bool use_tensor = tensor_ok && path_req != 1;
    // break_ties_with_noise (the reference's default) where the tensor path's exact stage cannot apply it -- shapes only the FMA
    // tiles handle (k_metrics 257..384, more than ~400 factors) or an RMB200_PATH=fma preference: the full-order path, which
    // adds every candidate's draw before sorting.  Only an explicit scoring_path = 1 keeps the FMA tiles (noise-free order).
    if (a.noise && !use_tensor && !use_full && !(ex && ex->scoring_path == 1)) { use_full = true; path_req = 3; }
    const bool fma_counts_first = use_tensor && count_ranks;
    // ... and inside the rank counts (ROC/PR-AUC with the noise): the FMA tiles count on the noise-free scores and report, per
    // held-out item, the candidates within the noise's reach; users for whom the noise decides a rank are handed to the full-order path
    const bool noise_handback = a.noise && count_ranks && !use_full;
    tm.scoring_path = use_full ? 3 : (use_tensor ? 2 : 1);
#ifndef RMB_F_CMID
#define RMB_F_CMID 512
#endif
    if (use_tensor) C = (K <= 32) ? 256 : (K <= 128 ? RMB_F_CMID : 1024);
    if (use_tensor) if (const char* env = std::getenv("RMB200_FILTER_C")) { const int v = std::atoi(env); if ((v == 256 || v == 512 || v == 1024) && v >= C) C = v; }   // developer
    if (use_full) C = round_up(K < a.n ? K : a.n, 32);              // row pitch of the ranked lists

    // ---- CSR slices of users [ub, ue), index pointers re-based to the slice ----
    int lo_hi[4];   // trp[ub], trp[ue], tep[ub], tep[ue]
    if (!on_dev) {
        lo_hi[0] = a.trp[ub]; lo_hi[1] = a.trp[ue]; lo_hi[2] = a.tep[ub]; lo_hi[3] = a.tep[ue];
    } else {
        CK(cudaMemcpy(&lo_hi[0], a.trp + ub, sizeof(int), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&lo_hi[1], a.trp + ue, sizeof(int), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&lo_hi[2], a.tep + ub, sizeof(int), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&lo_hi[3], a.tep + ue, sizeof(int), cudaMemcpyDeviceToHost));
    }
    if (lo_hi[1] < lo_hi[0] || lo_hi[3] < lo_hi[2]) { set_err("bad argument", "CSR index pointers decrease"); return RMB200_ERR_BAD_ARG; }
    const size_t nnz_tr = (size_t)(lo_hi[1] - lo_hi[0]);
    const size_t nnz_te = (size_t)(lo_hi[3] - lo_hi[2]);
    if ((nnz_tr && !a.tri) || (nnz_te && !a.tei)) { set_err("bad argument", "CSR indices missing"); return RMB200_ERR_BAD_ARG; }

    DevBuf d_trp, d_tri, d_tep, d_tei, d_tev, d_Arow[2];
    PrefetchStream pf;                                     // (declared after the buffers it fills: destroyed, i.e. drained, before them)
    if (!on_dev) CK(pf.init());
    CK(d_trp.alloc((size_t)(mr + 1) * sizeof(int)));
    CK(d_tep.alloc((size_t)(mr + 1) * sizeof(int)));
    const int* tri_d = nullptr; const int* tei_d = nullptr; const T* tev_d = nullptr;
    pt.start();
    if (!on_dev) {
        std::vector<int> tmp1((size_t)mr + 1), tmp2((size_t)mr + 1);
        for (int i = 0; i <= mr; i++) { tmp1[i] = a.trp[ub + i] - lo_hi[0]; tmp2[i] = a.tep[ub + i] - lo_hi[2]; }
        CK(cudaMemcpyAsync(d_trp.p, tmp1.data(), tmp1.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_tep.p, tmp2.data(), tmp2.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(d_tri.alloc(nnz_tr * sizeof(int)));
        CK(d_tei.alloc(nnz_te * sizeof(int)));
        if (nnz_tr) CK(upload_rows(d_tri.p, a.tri + lo_hi[0], nnz_tr * sizeof(int), nnz_tr * sizeof(int), 1, st));
        // the held-out rows are not needed before the first batch's metrics (rank counting: its first scoring step): they
        // go up on the prefetch stream, behind the item factors on the link instead of in front of them
        if (nnz_te) CK(upload_rows(d_tei.p, a.tei + lo_hi[2], nnz_te * sizeof(int), nnz_te * sizeof(int), 1, pf.s));
        if (a.tev) {
            CK(d_tev.alloc(nnz_te * sizeof(T)));
            if (nnz_te) CK(upload_rows(d_tev.p, a.tev + lo_hi[2], nnz_te * sizeof(T), nnz_te * sizeof(T), 1, pf.s));
            tev_d = d_tev.as<T>();
        }
        CK(cudaEventRecord(pf.test_rows, pf.s));
        pf.test_rows_pending = true;
        CK(cudaStreamSynchronize(st));   // tmp1/tmp2 go out of scope
        tri_d = d_tri.as<int>(); tei_d = d_tei.as<int>();
        tm.h2d_bytes += (int64_t)(2 * (size_t)(mr + 1) * sizeof(int) + (nnz_tr + nnz_te) * sizeof(int) + (a.tev ? nnz_te * sizeof(T) : 0));
        pt.stop(tm.h2d_ms);
    } else {
        rebase_indptr_kernel<<<(mr + 1 + 255) / 256, 256, 0, st>>>(a.trp + ub, lo_hi[0], d_trp.as<int>(), mr + 1);
        rebase_indptr_kernel<<<(mr + 1 + 255) / 256, 256, 0, st>>>(a.tep + ub, lo_hi[2], d_tep.as<int>(), mr + 1);
        CK(cudaGetLastError());
        tm.kernel_launches += 2;
        tri_d = a.tri ? a.tri + lo_hi[0] : nullptr;
        tei_d = a.tei ? a.tei + lo_hi[2] : nullptr;
        tev_d = a.tev ? a.tev + lo_hi[2] : nullptr;
        pt.stop(tm.prep_ms);
    }
    const int* trp_d = d_trp.as<int>();
    const int* tep_d = d_tep.as<int>();

    // ---- item factors: row-major staging copy, then the operand image of the chosen path ----
    DevBuf d_Bt, d_bias, d_Brow, d_Bb, d_maxbn, d_chunkn;
    const T* Bsrc = nullptr; size_t Bld = 0;
    if (!on_dev) {
        CK(d_Brow.alloc((size_t)a.n * a.k * sizeof(T)));
        pt.start();
        if (!nested) {
            int rc = stage_rows<T>(a.B, a.ldb, a.n, a.k, false, d_Brow, &Bsrc, &Bld, st, tm);
            if (rc) return rc;
        } else {
            // this device's slice of the rows from the host, the other slices from the peers' copies (see MultiCtx)
            MultiCtx& mc = *a.mc;
            const int G = mc.G, g = a.mc_rank;
            auto slice = [&](int h) { return (long long)a.n * h / G; };
            const size_t row_bytes = (size_t)a.k * sizeof(T);
            unsigned char* base = static_cast<unsigned char*>(d_Brow.p);
            if (slice(g + 1) > slice(g))
                CK(upload_rows(base + (size_t)slice(g) * row_bytes, a.B + (size_t)slice(g) * a.ldb, a.ldb * sizeof(T), row_bytes,
                               (size_t)(slice(g + 1) - slice(g)), st));
            tm.h2d_bytes += (int64_t)(slice(g + 1) - slice(g)) * (int64_t)row_bytes;
            CK(cudaStreamSynchronize(st));
            mc.Bptr[g] = d_Brow.p; mc.Bdev[g] = dev;
            if (!mc.meet(0)) { set_err("multi-GPU call", "another device failed"); return RMB200_ERR_CUDA; }
            for (int h = 1; h < G; h++) {
                const int src = (g + h) % G;
                const size_t off = (size_t)slice(src) * row_bytes, bytes = (size_t)(slice(src + 1) - slice(src)) * row_bytes;
                if (bytes) CK(cudaMemcpyPeerAsync(base + off, dev, static_cast<unsigned char*>(mc.Bptr[src]) + off, mc.Bdev[src], bytes, st));
            }
            CK(cudaStreamSynchronize(st));
            if (!mc.meet(1)) { set_err("multi-GPU call", "another device failed"); return RMB200_ERR_CUDA; }   // nobody frees its copy before everybody has fetched
            Bsrc = d_Brow.as<T>(); Bld = (size_t)a.k;
        }
        pt.stop(tm.h2d_ms);
    } else { Bsrc = a.B; Bld = a.ldb; }
    bool have_Bt = false;
    auto pack_Bt = [&]() -> int {          // operand image of the FMA path (the tensor path needs it only as a fall-back)
        if (have_Bt) return RMB200_OK;
        CK(d_Bt.alloc((size_t)p_pad * n_pad * sizeof(T)));
        pt.start();
        dim3 grid(n_pad / 32, (p_pad + 31) / 32), block(32, 8);
        pack_tiles_kernel<T, BN><<<grid, block, 0, st>>>(Bsrc, Bld, a.n, a.k, d_Bt.as<T>(), n_pad, p_pad);
        CK(cudaGetLastError());
        tm.kernel_launches++;
        pt.stop(tm.prep_ms);   // (synchronises)
        have_Bt = true;
        return RMB200_OK;
    };
    if (!use_tensor || fma_counts_first) {
        int rc = pack_Bt();
        if (rc) return rc;
        if (!use_tensor) {
            d_Brow.release();      // the FMA path reads only the tiled copy
            if (!on_dev) Bsrc = nullptr;
        }
    }
    const T* bias_d = nullptr;
    if (a.bias) {
        CK(d_bias.alloc((size_t)n_pad * sizeof(T)));
        DevBuf d_braw;
        const T* bsrc = a.bias;
        if (!on_dev) {
            CK(d_braw.alloc((size_t)a.n * sizeof(T)));
            pt.start();
            CK(cudaMemcpyAsync(d_braw.p, a.bias, (size_t)a.n * sizeof(T), cudaMemcpyHostToDevice, st));
            tm.h2d_bytes += (int64_t)a.n * (int64_t)sizeof(T);
            pt.stop(tm.h2d_ms);
            bsrc = d_braw.as<T>();
        }
        pt.start();
        pad_copy_kernel<T><<<(n_pad + 255) / 256, 256, 0, st>>>(bsrc, a.n, d_bias.as<T>(), n_pad);
        CK(cudaGetLastError());
        tm.kernel_launches++;
        pt.stop(tm.prep_ms);
        bias_d = d_bias.as<T>();
    }
    const int n_pad128 = round_up(a.n, 128);
    if (use_tensor) {
        CK(d_Bb.alloc((size_t)n_pad128 * KB * sizeof(__half)));
        CK(d_maxbn.alloc(4 * sizeof(unsigned)));                    // [0] max_j ||b_j||, [1] largest observed error / bound ratio (statistic)
        CK(d_chunkn.alloc((size_t)(n_pad128 / 32) * sizeof(float)));
        pt.start();
        CK(cudaMemsetAsync(d_maxbn.p, 0, 4 * sizeof(unsigned), st));
        CK(cudaMemsetAsync(d_chunkn.p, 0, (size_t)(n_pad128 / 32) * sizeof(float), st));
        const long long total = (long long)n_pad128 * (KB / 8);
        row_norm_kernel<T><<<(a.n + NORM_ROWS_PER_BLOCK - 1) / NORM_ROWS_PER_BLOCK, 256, 0, st>>>(Bsrc, Bld, a.n, a.k, bias_d, 0, nullptr,
                                                                                              d_maxbn.as<unsigned>(), d_chunkn.as<float>());
        CK(cudaGetLastError());
        pack_f16_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Bsrc, Bld, a.n, a.k, bias_d, 0, nullptr, d_maxbn.as<unsigned>(),
                                                                            d_Bb.as<__half>(), n_pad128, KB);
        CK(cudaGetLastError());
        tm.kernel_launches += 2;
        pt.stop(tm.prep_ms);
    }

    // ---- per-user state ----
    DevBuf d_status, d_flags, d_umin, d_pos_raw, d_pos_sorted, d_pos_perm, d_pos_item, d_auc, d_log2, d_pos_rank;
    CK(d_status.alloc((size_t)mr * sizeof(int)));
    CK(d_flags.alloc((size_t)mr * sizeof(int)));
    if (count_ranks) {
        CK(d_umin.alloc((size_t)mr * sizeof(unsigned long long)));
        CK(d_pos_raw.alloc(nnz_te * sizeof(T)));
        CK(d_pos_sorted.alloc(nnz_te * sizeof(T)));
        CK(d_pos_perm.alloc(nnz_te * sizeof(int)));
        CK(d_pos_item.alloc(nnz_te * sizeof(int)));
        CK(d_auc.alloc(nnz_te * sizeof(unsigned int)));
        CK(cudaMemsetAsync(d_auc.p, 0, nnz_te * sizeof(unsigned int) + (nnz_te ? 0 : 16), st));
        CK(cudaMemsetAsync(d_pos_sorted.p, 0, nnz_te * sizeof(T) + (nnz_te ? 0 : 16), st));
        CK(cudaMemsetAsync(d_pos_perm.p, 0, nnz_te * sizeof(int) + (nnz_te ? 0 : 16), st));
        CK(cudaMemsetAsync(d_pos_item.p, 0, nnz_te * sizeof(int) + (nnz_te ? 0 : 16), st));
    }
    long long* pos_rank_d = nullptr;
    if (ex && ex->pos_rank) {
        if (on_dev) pos_rank_d = reinterpret_cast<long long*>(ex->pos_rank) + lo_hi[2];
        else { CK(d_pos_rank.alloc(nnz_te * sizeof(long long))); pos_rank_d = d_pos_rank.as<long long>(); }
        CK(cudaMemsetAsync(pos_rank_d, 0, nnz_te * sizeof(long long), st));
    }
    {
        std::vector<double> l2((size_t)K);
        for (int i = 0; i < K; i++) l2[i] = std::log2((double)(i + 2));   // hpp:620 std::log2(ix+2)
        CK(d_log2.alloc((size_t)K * sizeof(double)));
        CK(cudaMemcpy(d_log2.p, l2.data(), (size_t)K * sizeof(double), cudaMemcpyHostToDevice));
    }
    pt.start();
    {
        StatusParams sp;
        sp.trp = trp_d; sp.tep = tep_d; sp.n = a.n; sp.K = K; sp.user_begin = 0; sp.user_end = mr;
        sp.has_ndcg = a.out[5] != nullptr;
        sp.has_rescue = (a.out[8] || a.out[9] || a.out[3] || a.out[4] || a.out[7]) ? 1 : 0;   // hpp:485
        sp.consider_cold_start = a.consider_cold_start; sp.min_items_pool = mip; sp.min_pos_test = mpt;
        sp.ustatus = d_status.as<int>(); sp.uflags = d_flags.as<int>();
        sp.umin = count_ranks ? d_umin.as<unsigned long long>() : nullptr;
        user_status_kernel<<<(mr + 255) / 256, 256, 0, st>>>(sp);
        CK(cudaGetLastError());
        tm.kernel_launches++;
    }
    pt.stop(tm.prep_ms);

    // ---- user batches ----
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int wave_users = nsm * BM;                       // one CTA (128 users) per SM
    int UB = 8 * wave_users;                               // users per batch: 8 full waves ...
    // ... 4 when the filter's item image is much larger than the L2 (256 MB at 1M items x 128 factors): the CTAs of a longer
    // launch drift further apart and re-stream the item tiles from HBM (cfg4: 3.16 M users/s at 4 waves, 3.10 at 8, 2.96 at
    // 16, 2.76 with all 1M users in one launch; smaller catalogues prefer 8; profiles/r02_batch_size.txt)
    if (use_tensor && (size_t)round_up(a.n, 128) * KB * sizeof(__half) > (size_t)160 << 20) UB = 4 * wave_users;
    if (const char* env = std::getenv("RMB200_BATCH_USERS")) { const int v = std::atoi(env); if (v > 0) UB = round_up(v, BM); }
    int fo_chunk = 0; size_t fo_bytes = 0;
    if (use_full) {
        // ranked lists (K entries per user) and cumulative rows (K columns per metric) of a batch within ~2 GB
        const long long per_user = (long long)C * (long long)(sizeof(T) + 4) + (a.cumulative ? 8ll * K * (long long)sizeof(T) : 0);
        long long cap = (2ll << 30) / (per_user > 0 ? per_user : 1);
        cap = cap / BM * BM;
        if (cap < BM) cap = BM;
        if (UB > cap) UB = (int)cap;
    }
    if (UB > round_up(mr, BM)) UB = round_up(mr, BM);
    DevBuf d_fo, d_near, d_nz_mark, d_nz_list, d_nz_cnt, d_At_nz;
    if (use_full) {
        full_order_plan(a.n, (int)sizeof(T), UB, &fo_chunk, &fo_bytes);
        CK(d_fo.alloc(fo_bytes));
    }
    if (noise_handback) {
        CK(d_near.alloc(nnz_te * sizeof(unsigned)));
        CK(cudaMemsetAsync(d_near.p, 0, nnz_te * sizeof(unsigned) + (nnz_te ? 0 : 16), st));
        CK(d_nz_mark.alloc((size_t)UB * sizeof(int)));
        CK(d_nz_list.alloc((size_t)UB * sizeof(int)));
        CK(d_nz_cnt.alloc(16));
    }

    DevBuf d_At, d_cs, d_ci, d_cc, d_out[2][10], d_tki[2], d_tks[2], d_stat_out[2], d_Ab, d_anorm;
    ResultStream rsx;                                     // (declared after the buffers it reads: drained before they are freed)
    if (!on_dev) CK(rsx.init());
    CK(d_At.alloc((size_t)p_pad * UB * sizeof(T)));
    DevBuf d_capx, d_overflow, d_fb_list, d_At_fb;
    // sampled threshold guess of the filter (filter_select.cuh, pass 0): every stride-th item tile, about 7/K of the
    // catalogue (1/16 .. 1/8); the guess is the r-th best of the sample with r = the smallest rank for which
    // P(Poisson(K * sample / n) >= r) <= 1e-6, i.e. fewer than K items of the whole catalogue reach it about once
    // in a million users (those users' CTAs walk the catalogue a second time; results never depend on the guess)
    int f_sample_tiles = 0, f_sample_stride = 1, f_sample_rank = 0;
    if (use_tensor) {
        const int NT = (a.n + 127) / 128;
        int min_tiles = 512;
        if (const char* env = std::getenv("RMB200_SAMPLE_MIN_TILES")) { const int v = std::atoi(env); if (v >= 16) min_tiles = v; }   // developer
        bool on = NT >= min_tiles;
        if (const char* env = std::getenv("RMB200_SAMPLE")) on = on && std::atoi(env) != 0;
        if (on) {
            double frac = 7.0 / K;
            if (frac < 1.0 / 16) frac = 1.0 / 16;
            if (frac > 1.0 / 8) frac = 1.0 / 8;
            int stride = (int)(1.0 / frac);
            if (const char* env = std::getenv("RMB200_SAMPLE_STRIDE")) { const int v = std::atoi(env); if (v >= 2) stride = v; }
            const int ns = NT / stride;                                  // tiles 0, stride, ..., (ns-1)*stride: all full tiles
            const double lam = (double)K * ((double)ns * 128.0) / (double)a.n;
            double term = std::exp(-lam), cdf = 0.0;                     // P(X <= r-1), X ~ Poisson(lam)
            int r = 0;
            while (r < 4096) { cdf += term; r++; term *= lam / r; if (1.0 - cdf <= 1e-6) break; }
            if (const char* env = std::getenv("RMB200_SAMPLE_RANK")) { const int v = std::atoi(env); if (v >= 1) r = v; }
            // (the guess is the r-th largest of up to min(C / 4, 128) group maxima of the sample, filter_select.cuh)
            const int groups = (C / 4 < 128 ? C / 4 : 128) < ns ? (C / 4 < 128 ? C / 4 : 128) : ns;
            if (ns >= 8 && 2 * r <= groups) { f_sample_tiles = ns; f_sample_stride = stride; f_sample_rank = r; }
        }
    }
    if (use_tensor) {
        CK(d_Ab.alloc((size_t)UB * KB * sizeof(__half)));
        CK(d_anorm.alloc((size_t)UB * sizeof(float)));
        CK(d_capx.alloc((size_t)UB * C * sizeof(uint2)));
        CK(d_overflow.alloc(40 * sizeof(int)));    // [0] users handed back to the FMA path, [1] users that took the retry pass, [6] length of the hand-back list, [8..] cycle counters (stats builds)
        CK(d_fb_list.alloc((size_t)UB * sizeof(int)));
    }
    // upload of users [b0, b0 + UB) into staging buffer `slot`, on the prefetch stream
    auto upload_users = [&](int b0, int slot) -> int {
        const int nb = (mr - b0) < UB ? (mr - b0) : UB;
        if (pf.freed_valid[slot]) CK(cudaStreamWaitEvent(pf.s, pf.freed[slot], 0));
        CK(upload_rows(d_Arow[slot].p, a.A + (size_t)(ub + b0) * a.lda, a.lda * sizeof(T), (size_t)a.k * sizeof(T), (size_t)nb, pf.s));
        tm.h2d_bytes += (int64_t)nb * a.k * (int64_t)sizeof(T);
        CK(cudaEventRecord(pf.ready[slot], pf.s));
        return RMB200_OK;
    };
    if (!on_dev) {
        CK(d_Arow[0].alloc((size_t)UB * a.k * sizeof(T)));
        if (UB < mr) CK(d_Arow[1].alloc((size_t)UB * a.k * sizeof(T)));
        int rc = upload_users(0, 0);
        if (rc) return rc;
    }
    // A call with fewer user tiles than SMs (small m) cuts the catalogue into item ranges on the FMA tiles: one CTA per (user tile,
    // range), every range keeping its own best K, joined afterwards (merge_slices_kernel); rank counts are added atomically.
    int fma_slices = 1;
    if (!use_full) {
        const int tiles = UB / BM, NT_all = (a.n + BN - 1) / BN;
        int sl = tiles > 0 ? nsm / tiles : 1;
        if (sl > NT_all / 8) sl = NT_all / 8;                       // (at least eight item tiles per range)
        if (sl > C / K) sl = C / K;                                 // the joined heads fit one row ...
        if (sl > 512 / K) sl = 512 / K;                             // ... and the final sort
        if (const char* env = std::getenv("RMB200_SLICES")) { const int v = std::atoi(env); if (v >= 1 && v < sl) sl = v; }      // developer
        if (sl > 1) fma_slices = sl;
    }
    CK(d_cs.alloc((size_t)UB * fma_slices * C * sizeof(T)));
    CK(d_ci.alloc((size_t)UB * fma_slices * C * sizeof(int)));
    CK(d_cc.alloc((size_t)UB * fma_slices * sizeof(int)));
    if (!on_dev) {
        for (int set = 0; set < (UB < mr ? 2 : 1); set++) {
            for (int q = 0; q < 10; q++)
                if (a.out[q]) CK(d_out[set][q].alloc((size_t)UB * (q < 8 ? rs : 1) * sizeof(T)));
            if (ex && ex->topk_items) CK(d_tki[set].alloc((size_t)UB * K * sizeof(int)));
            if (ex && ex->topk_scores) CK(d_tks[set].alloc((size_t)UB * K * sizeof(T)));
            if (ex && ex->status) CK(d_stat_out[set].alloc((size_t)UB * sizeof(int)));
        }
    }

    // per-metric means over the users of the call (extension): partial sums per 256 users, added up in order at the end
    DevBuf d_psum, d_pcnt, d_means, d_mcounts;
    const int mean_W = a.cumulative ? K : 1;
    int mean_blocks_total = 0, mean_block0 = 0;
    if (want_means) {
        for (int b0 = 0; b0 < mr; b0 += UB) mean_blocks_total += (((mr - b0) < UB ? (mr - b0) : UB) + MEAN_CHUNK - 1) / MEAN_CHUNK;
        CK(d_psum.alloc((size_t)mean_blocks_total * 10 * mean_W * sizeof(double)));
        CK(d_pcnt.alloc((size_t)mean_blocks_total * 10 * mean_W * sizeof(int)));
    }

    for (int b0 = 0; b0 < mr; b0 += UB) {
        if (g_interrupt.load()) break;                     // hpp:488-489
        const int nb = (mr - b0) < UB ? (mr - b0) : UB;
        const int nb_pad = round_up(nb, BM);
        const int a_slot = (b0 / UB) & 1;                  // staging buffer holding this batch's user factors (host inputs)
        const int r_set = (b0 / UB) & 1;                   // set of result staging buffers of this batch (host outputs)
        // the kernels reading the staged rows are queued: mark the buffer; once the scoring kernel is queued as well, the
        // next batch's rows go up into the other buffer (pageable sources keep this thread busy copying meanwhile)
        auto staged_rows_consumed = [&]() -> int {
            if (on_dev) return RMB200_OK;
            CK(cudaEventRecord(pf.freed[a_slot], st));
            pf.freed_valid[a_slot] = true;
            return RMB200_OK;
        };
        bool prefetched = false;
        auto prefetch_next_users = [&]() -> int {
            if (on_dev || prefetched || b0 + UB >= mr) return RMB200_OK;
            prefetched = true;
            return upload_users(b0 + UB, a_slot ^ 1);
        };

        // user factors of the batch -> k-major
        const T* Asrc = nullptr; size_t Ald = 0;
        if (!on_dev) {
            pt.start();
            CK(cudaStreamWaitEvent(st, pf.ready[a_slot], 0));     // uploaded one batch ahead: only what is still exposed is timed
            Asrc = d_Arow[a_slot].as<T>(); Ald = (size_t)a.k;
            pt.stop(tm.h2d_ms);
        } else { Asrc = a.A + (size_t)(ub + b0) * a.lda; Ald = a.lda; }
        pt.start();
        {
            dim3 grid(nb_pad / 32, (p_pad + 31) / 32), block(32, 8);
            pack_tiles_kernel<T, BM><<<grid, block, 0, st>>>(Asrc, Ald, nb, a.k, d_At.as<T>(), nb_pad, p_pad);
            CK(cudaGetLastError());
            tm.kernel_launches++;

            if (!use_tensor) { int rc = staged_rows_consumed(); if (rc) return rc; }
        }
        if (count_ranks && !use_full) {
            if (pf.test_rows_pending) { CK(cudaStreamWaitEvent(st, pf.test_rows, 0)); pf.test_rows_pending = false; }
            const int blocks = (nb + 7) / 8 < 8 * nsm ? (nb + 7) / 8 : 8 * nsm;
            score_entries_kernel<T, BM><<<blocks, 256, 0, st>>>(d_At.as<T>(), d_Bt.as<T>(), bias_d, p_pad, b0, nb,
                                                             tep_d, tei_d, d_status.as<int>(), d_pos_raw.as<T>());
            CK(cudaGetLastError());
            sort_positives_kernel<T><<<blocks, 256, 0, st>>>(b0, nb, tep_d, tei_d, d_status.as<int>(), d_pos_raw.as<T>(),
                                                              d_pos_sorted.as<T>(), d_pos_perm.as<int>(), d_pos_item.as<int>());
            CK(cudaGetLastError());
            tm.kernel_launches += 2;
        }
        pt.stop(tm.prep_ms);

        bool pk_pending = false;
        // fused score / exclude / select (/ rank counting) on the FMA pipe: the whole batch, or (umap) the listed users only
        auto run_fma = [&](const T* At_ptr, int n_rows, const int* umap, bool with_counts, bool timed) -> int {
            ScoreSelectParams<T> sp = ScoreSelectParams<T>();       // (every field defined: optional ones stay null)
            sp.At = At_ptr; sp.Bt = d_Bt.as<T>(); sp.bias = bias_d;
            sp.p_pad = p_pad; sp.n = a.n; sp.mb = n_rows; sp.user0 = b0;
            sp.trp = trp_d; sp.tri = tri_d; sp.tep = tep_d; sp.ustatus = d_status.as<int>();
            sp.cand_score = d_cs.as<T>(); sp.cand_item = d_ci.as<int>(); sp.cand_count = d_cc.as<int>();
            sp.uflags = d_flags.as<int>(); sp.K = K;
            sp.pos_sorted = with_counts ? d_pos_sorted.as<T>() : nullptr;
            sp.auc_cnt = with_counts ? d_auc.as<unsigned int>() : nullptr;
            sp.pos_item = with_counts ? d_pos_item.as<int>() : nullptr;
            sp.auc_near = (with_counts && noise_handback) ? d_near.as<unsigned>() : nullptr;
            sp.umin = with_counts ? d_umin.as<unsigned long long>() : nullptr;
            sp.umap = umap;
            const int slices = (umap == nullptr && !use_tensor) ? fma_slices : 1;
            sp.slice_rows = slices > 1 ? UB : 0;
            if (const char* env = std::getenv("RMB200_AUC_DBG")) sp.dbg = std::atoi(env);      // developer: timing experiments (wrong results)
            if (timed) cudaEventRecord(pk.a, st);
            CK(launch_score_select<T>(sp, C, with_counts, n_rows, slices, st));
            if (timed) { cudaEventRecord(pk.b, st); pk_pending = true; }
            tm.kernel_launches++;
            if (slices > 1) {
                merge_slices_kernel<T><<<(n_rows + 7) / 8, 256, 0, st>>>(d_cs.as<T>(), d_ci.as<int>(), d_cc.as<int>(), C, n_rows, slices, UB);
                CK(cudaGetLastError());
                tm.kernel_launches++;
            }
            return RMB200_OK;
        };
        auto run_fma_batch = [&]() -> int {
            int rc = run_fma(d_At.as<T>(), nb, nullptr, count_ranks, true);
            if (rc) return rc;
            return prefetch_next_users();
        };
        if (fma_counts_first) {            // rank counts (and the smallest candidate score) from the FMA tiles; its top-K is replaced below
            pt.start();
            int rc = run_fma_batch();
            if (rc) return rc;
            pt.stop(tm.score_select_ms);
        }
        if (use_tensor) {
            // norms + fp16 operand image of the batch's users, then the tensor-core filter and the exact re-scoring
            pt.start();
            const long long total = (long long)nb_pad * (KB / 8);
            row_norm_kernel<T><<<(nb + NORM_ROWS_PER_BLOCK - 1) / NORM_ROWS_PER_BLOCK, 256, 0, st>>>(Asrc, Ald, nb, a.k, (const T*)nullptr, a.bias ? 1 : 0,
                                                                                               d_anorm.as<float>(), nullptr, nullptr);
            CK(cudaGetLastError());
            pack_f16_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Asrc, Ald, nb, a.k, (const T*)nullptr, a.bias ? 1 : 0,
                                                                                d_anorm.as<float>(), nullptr, d_Ab.as<__half>(), nb_pad, KB);
            CK(cudaGetLastError());
            CK(cudaMemsetAsync(d_overflow.p, 0, 40 * sizeof(int), st));
            tm.kernel_launches += 2;
            pt.stop(tm.prep_ms);
            pt.start();
            const FilterErrCoef ec = filter_err_coefs(KB);
            FilterParams fp;
            fp.Ab = d_Ab.as<__half>(); fp.Bb = d_Bb.as<__half>(); fp.KB = KB; fp.stages = f_stages;
            fp.c_rel = ec.c_rel; fp.c_abs = ec.c_abs; fp.c_const = ec.c_const;
            fp.n = a.n; fp.mb = nb; fp.user0 = b0;
            fp.anorm = d_anorm.as<float>(); fp.maxbn = d_maxbn.as<unsigned>(); fp.chunk_norm = d_chunkn.as<float>();
            fp.trp = trp_d; fp.tri = tri_d; fp.ustatus = d_status.as<int>();
            fp.cand = d_capx.as<uint2>(); fp.cand_count = d_cc.as<int>();
            fp.overflow = d_overflow.as<int>(); fp.uflags = d_flags.as<int>(); fp.K = K;
            fp.sample_tiles = f_sample_tiles; fp.sample_stride = f_sample_stride; fp.sample_rank = f_sample_rank;
            fp.retries = d_overflow.as<int>() + 1;
            fp.noise_band = a.noise ? 2.02e-12f : 0.f;
            fp.dbg = 0; if (const char* env = std::getenv("RMB200_DBG")) fp.dbg = std::atoi(env);
            cudaEventRecord(pk.a, st);
            CK(launch_filter_select(fp, C, nb_pad / BM, st));
            cudaEventRecord(pk.b, st);
            pk_pending = true;
            tm.kernel_launches++;
            // the exact stage reads the staged user factors (through d_At) and the filter is queued: the next batch's rows can go up
            { int rc = staged_rows_consumed(); if (rc) return rc; }
            { int rc = prefetch_next_users(); if (rc) return rc; }
            NoiseArgs nz;
            nz.on = a.noise; nz.seed_user0 = (unsigned long long)a.seed + (unsigned long long)(ub + b0); nz.trp = trp_d; nz.tri = tri_d; nz.n = a.n;
            FilterErrStat es;
            es.anorm = d_anorm.as<float>(); es.maxbn = d_maxbn.as<unsigned>(); es.chunk_norm = d_chunkn.as<float>();
            es.c_rel = ec.c_rel; es.c_abs = ec.c_abs; es.c_const = ec.c_const;
            es.max_ratio_bits = want_filter_stats ? d_maxbn.as<unsigned>() + 1 : nullptr;
            CK(launch_exact_topk<T>(d_capx.as<uint2>(), d_cs.as<T>(), d_ci.as<int>(), d_cc.as<int>(), C, nb, b0, d_At.as<T>(), p_pad, a.k,
                                    Bsrc, Bld, bias_d, d_flags.as<int>(), K, nz, es, st));
            tm.kernel_launches++;
            int over_retry[40] = {0};
            CK(cudaMemcpyAsync(over_retry, d_overflow.p, 40 * sizeof(int), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
#if RMB_F_STATS
            std::fprintf(stderr, "[rmb200 stats] users %d: appends/user %.1f cuts/user %.2f slow-path entries/user %.1f cursor moves/user %.1f retries %d handed back %d (sample tiles %d stride %d rank %d)\n",
                         nb, over_retry[2] / (double)nb, over_retry[3] / (double)nb, over_retry[4] / (double)nb, over_retry[5] / (double)nb,
                         over_retry[1], over_retry[0], f_sample_tiles, f_sample_stride, f_sample_rank);
            {
                const unsigned long long* ck = reinterpret_cast<const unsigned long long*>(over_retry + 8);
                const double nc = (double)(nb_pad / BM);
                for (int ps = 0; ps < 2; ps++)
                    std::fprintf(stderr, "[rmb200 stats]   pass %d, one epilogue warp, kcycles per CTA: wait acc %.0f, ld+fast %.0f, append %.0f, cursor %.0f, meet+cut %.0f, total %.0f\n",
                                 ps, ck[ps * 6 + 0] / nc / 1e3, ck[ps * 6 + 1] / nc / 1e3, ck[ps * 6 + 2] / nc / 1e3, ck[ps * 6 + 3] / nc / 1e3, ck[ps * 6 + 4] / nc / 1e3, ck[ps * 6 + 5] / nc / 1e3);
            }
#endif
            const int n_over = over_retry[0];
            tm.filter_retry_rows += over_retry[1];
            if (n_over > 0) {
                // some users' kept band did not fit their candidate buffer (near-constant scores) or their factors cannot be scaled
                // into fp16 range: THOSE users are re-run on the FMA tiles (same results), everybody else keeps the filter's result
                tm.filter_fallback_batches++;
                tm.filter_fallback_users += n_over;
                pt.stop(tm.score_select_ms);
                int rc = pack_Bt();
                if (rc) return rc;
                pt.start();
                const int n_over_pad = round_up(n_over, BM);
                CK(d_At_fb.alloc((size_t)p_pad * n_over_pad * sizeof(T)));
                collect_flagged_kernel<<<1, 1024, 0, st>>>(d_cc.as<int>(), nb, d_fb_list.as<int>(), d_overflow.as<int>() + 6);
                CK(cudaGetLastError());
                dim3 grid(n_over_pad / 32, (p_pad + 31) / 32), block(32, 8);
                pack_tiles_gather_kernel<T, BM><<<grid, block, 0, st>>>(Asrc, Ald, d_fb_list.as<int>(), n_over, a.k, d_At_fb.as<T>(), n_over_pad, p_pad);
                CK(cudaGetLastError());
                tm.kernel_launches += 2;
                { int rc2 = staged_rows_consumed(); if (rc2) return rc2; }      // (the gather read the staged rows once more)
                if (n_over <= 2048 && a.n >= 32768) {
                    // a few users of a large catalogue: the full-order path (every candidate scored by item-parallel blocks, one
                    // segmented sort) -- on the FMA tiles ONE CTA walked the whole catalogue for them (78 ms at 1M items for a
                    // single user, profiles/r02v: two of four ranks of a strong-scaling run ran at half speed for it)
                    int chunk = 0; size_t bytes = 0;
                    full_order_plan(a.n, (int)sizeof(T), n_over, &chunk, &bytes);
                    CK(d_fo.alloc(bytes));
                    FullOrderArgs<T> fo;
                    fo.At = d_At_fb.as<T>(); fo.Bt = d_Bt.as<T>(); fo.bias = bias_d; fo.p_pad = p_pad; fo.n = a.n; fo.K = K; fo.C = C;
                    fo.user0 = b0; fo.nb = n_over; fo.umap = d_fb_list.as<int>();
                    fo.trp = trp_d; fo.tri = tri_d; fo.tep = tep_d; fo.tei = tei_d; fo.ustatus = d_status.as<int>(); fo.uflags = d_flags.as<int>();
                    fo.cand_score = d_cs.as<T>(); fo.cand_item = d_ci.as<int>(); fo.cand_count = d_cc.as<int>();
                    fo.umin = count_ranks ? d_umin.as<unsigned long long>() : nullptr;
                    fo.auc_cnt = count_ranks ? d_auc.as<unsigned int>() : nullptr;
                    fo.pos_perm = count_ranks ? d_pos_perm.as<int>() : nullptr;
                    fo.noise = a.noise; fo.seed_user0 = (unsigned long long)a.seed + (unsigned long long)(ub + b0);
                    fo.scratch = d_fo.p; fo.scratch_bytes = bytes; fo.chunk_users = chunk;
                    if (pf.test_rows_pending && count_ranks) { CK(cudaStreamWaitEvent(st, pf.test_rows, 0)); pf.test_rows_pending = false; }
                    long long nl = 0;
                    CK(full_order_run<T>(fo, st, &nl));
                    tm.kernel_launches += nl;
                } else {
                    rc = run_fma(d_At_fb.as<T>(), n_over, d_fb_list.as<int>(), false, false);
                    if (rc) return rc;
                }
            }
        } else if (use_full) {
            // every candidate scored, (noise,) sorted: ranked top-K, smallest candidate score and held-out ranks off the sorted lists
            pt.start();
            if (pf.test_rows_pending) { CK(cudaStreamWaitEvent(st, pf.test_rows, 0)); pf.test_rows_pending = false; }
            FullOrderArgs<T> fo;
            fo.At = d_At.as<T>(); fo.Bt = d_Bt.as<T>(); fo.bias = bias_d; fo.p_pad = p_pad; fo.n = a.n; fo.K = K; fo.C = C;
            fo.user0 = b0; fo.nb = nb; fo.umap = nullptr;
            fo.trp = trp_d; fo.tri = tri_d; fo.tep = tep_d; fo.tei = tei_d; fo.ustatus = d_status.as<int>(); fo.uflags = d_flags.as<int>();
            fo.cand_score = d_cs.as<T>(); fo.cand_item = d_ci.as<int>(); fo.cand_count = d_cc.as<int>();
            fo.umin = count_ranks ? d_umin.as<unsigned long long>() : nullptr;
            fo.auc_cnt = count_ranks ? d_auc.as<unsigned int>() : nullptr;
            fo.pos_perm = count_ranks ? d_pos_perm.as<int>() : nullptr;
            fo.noise = a.noise; fo.seed_user0 = (unsigned long long)a.seed + (unsigned long long)(ub + b0);
            fo.scratch = d_fo.p; fo.scratch_bytes = fo_bytes; fo.chunk_users = fo_chunk;
            cudaEventRecord(pk.a, st);
            long long nl = 0;
            CK(full_order_run<T>(fo, st, &nl));
            cudaEventRecord(pk.b, st);
            pk_pending = true;
            tm.kernel_launches += nl;
            { int rc = prefetch_next_users(); if (rc) return rc; }
        } else {
            pt.start();
            int rc = run_fma_batch();
            if (rc) return rc;
        }
        if (!use_full) {
            CK(launch_rank_topk<T>(d_cs.as<T>(), d_ci.as<int>(), d_cc.as<int>(), C, nb, use_tensor ? K : K * fma_slices, st));
            tm.kernel_launches++;
        }
        if (noise_handback) {
            // users for whom the tie-breaking noise decides a rank: their ranked top-K and ranks from the full-order path
            mark_noise_users_kernel<<<(nb + 255) / 256, 256, 0, st>>>(tep_d, d_near.as<unsigned>(), d_status.as<int>(), b0, nb, d_nz_mark.as<int>());
            CK(cudaGetLastError());
            CK(cudaMemsetAsync(d_nz_cnt.p, 0, 16, st));
            collect_flagged_kernel<<<1, 1024, 0, st>>>(d_nz_mark.as<int>(), nb, d_nz_list.as<int>(), d_nz_cnt.as<int>());
            CK(cudaGetLastError());
            tm.kernel_launches += 2;
            int n_nz = 0;
            CK(cudaMemcpyAsync(&n_nz, d_nz_cnt.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (n_nz > 0) {
                tm.noise_handback_users += n_nz;
                int rc = pack_Bt();
                if (rc) return rc;
                const int n_pad_u = round_up(n_nz, BM);
                CK(d_At_nz.alloc((size_t)p_pad * n_pad_u * sizeof(T)));
                dim3 grid(n_pad_u / 32, (p_pad + 31) / 32), block(32, 8);
                pack_tiles_gather_kernel<T, BM><<<grid, block, 0, st>>>(Asrc, Ald, d_nz_list.as<int>(), n_nz, a.k, d_At_nz.as<T>(), n_pad_u, p_pad);
                CK(cudaGetLastError());
                tm.kernel_launches++;
                { int rc2 = staged_rows_consumed(); if (rc2) return rc2; }
                int chunk = 0; size_t bytes = 0;
                full_order_plan(a.n, (int)sizeof(T), n_nz, &chunk, &bytes);
                CK(d_fo.alloc(bytes));
                FullOrderArgs<T> fo;
                fo.At = d_At_nz.as<T>(); fo.Bt = d_Bt.as<T>(); fo.bias = bias_d; fo.p_pad = p_pad; fo.n = a.n; fo.K = K; fo.C = C;
                fo.user0 = b0; fo.nb = n_nz; fo.umap = d_nz_list.as<int>();
                fo.trp = trp_d; fo.tri = tri_d; fo.tep = tep_d; fo.tei = tei_d; fo.ustatus = d_status.as<int>(); fo.uflags = d_flags.as<int>();
                fo.cand_score = d_cs.as<T>(); fo.cand_item = d_ci.as<int>(); fo.cand_count = d_cc.as<int>();
                fo.umin = d_umin.as<unsigned long long>(); fo.auc_cnt = d_auc.as<unsigned int>(); fo.pos_perm = d_pos_perm.as<int>();
                fo.noise = 1; fo.seed_user0 = (unsigned long long)a.seed + (unsigned long long)(ub + b0);
                fo.scratch = d_fo.p; fo.scratch_bytes = bytes; fo.chunk_users = chunk;
                long long nl = 0;
                CK(full_order_run<T>(fo, st, &nl));
                tm.kernel_launches += nl;
            }
        }
        if (a.noise && !use_tensor && !use_full && !count_ranks) {
            all_equal_check_kernel<T><<<nb, 128, 0, st>>>(d_cs.as<T>(), d_cc.as<int>(), C, b0, K, a.n, trp_d, tri_d, d_status.as<int>(),
                                                           d_At.as<T>(), d_Bt.as<T>(), bias_d, p_pad, d_flags.as<int>());
            CK(cudaGetLastError());
            tm.kernel_launches++;
        }
        pt.stop(tm.score_select_ms);
        if (pk_pending) { float ms = 0; cudaEventElapsedTime(&ms, pk.a, pk.b); tm.dominant_kernel_ms += ms; }

        // per-user metrics
        pt.start();
        {
            MetricsParams<T> mp;
            mp.n = a.n; mp.K = K; mp.C = C; mp.user0 = b0; mp.mb = nb; mp.cumulative = a.cumulative;
            mp.want_roc = want_roc; mp.want_pr = want_pr; mp.count_ranks = count_ranks; mp.noise = a.noise;
            mp.trp = trp_d; mp.tep = tep_d; mp.tei = tei_d; mp.tev = tev_d;
            mp.ustatus = d_status.as<int>(); mp.uflags = d_flags.as<int>();
            mp.cand_score = d_cs.as<T>(); mp.cand_item = d_ci.as<int>(); mp.cand_count = d_cc.as<int>();
            mp.auc_cnt = d_auc.as<unsigned int>(); mp.umin = d_umin.as<unsigned long long>();
            mp.pos_perm = d_pos_perm.as<int>(); mp.log2tab = d_log2.as<double>();
            mp.nan_value = nan_value;
            T* outs[10];
            if (!on_dev && rsx.copied_valid[r_set]) CK(cudaStreamWaitEvent(st, rsx.copied[r_set], 0));      // the set's previous rows have left
            for (int q = 0; q < 10; q++) {
                const size_t stride = q < 8 ? rs : 1;
                if (!a.out[q]) outs[q] = nullptr;
                else if (on_dev) outs[q] = a.out[q] + (size_t)(ub + b0) * stride;
                else outs[q] = d_out[r_set][q].as<T>();
            }
            mp.p = outs[0]; mp.tp = outs[1]; mp.r = outs[2]; mp.ap = outs[3]; mp.tap = outs[4];
            mp.ndcg = outs[5]; mp.hit = outs[6]; mp.rr = outs[7]; mp.roc = outs[8]; mp.pr = outs[9];
            mp.status_out = nullptr; mp.topk_items = nullptr; mp.topk_scores = nullptr;
            if (ex && ex->status) mp.status_out = on_dev ? ex->status + ub + b0 : d_stat_out[r_set].as<int>();
            if (ex && ex->topk_items) mp.topk_items = on_dev ? ex->topk_items + (size_t)(ub + b0) * K : d_tki[r_set].as<int>();
            if (ex && ex->topk_scores) mp.topk_scores = on_dev ? reinterpret_cast<T*>(ex->topk_scores) + (size_t)(ub + b0) * K : d_tks[r_set].as<T>();
            mp.pos_rank = pos_rank_d;
            if (pf.test_rows_pending) { CK(cudaStreamWaitEvent(st, pf.test_rows, 0)); pf.test_rows_pending = false; }
            user_metrics_kernel<T><<<(nb + METRICS_WARPS - 1) / METRICS_WARPS, METRICS_WARPS * 32, 0, st>>>(mp);
            CK(cudaGetLastError());
            tm.kernel_launches++;
            if (want_means) {
                MeanParams<T> qp;
                for (int q = 0; q < 10; q++) qp.out[q] = outs[q];
                qp.nb = nb; qp.W = mean_W; qp.part_sum = d_psum.as<double>(); qp.part_cnt = d_pcnt.as<int>(); qp.block0 = mean_block0;
                const int blocks = (nb + MEAN_CHUNK - 1) / MEAN_CHUNK;
                metric_partial_kernel<T><<<dim3(blocks, 10), MEAN_CHUNK, 0, st>>>(qp);
                CK(cudaGetLastError());
                tm.kernel_launches++;
                mean_block0 += blocks;
            }
        }
        pt.stop(tm.metrics_ms);

        // results of the batch -> caller's arrays at the shard's row offset (no gather collective:
        // every shard writes its own disjoint rows)
        if (!on_dev) {
            CK(cudaEventRecord(rsx.ready[r_set], st));
            CK(cudaStreamWaitEvent(rsx.s, rsx.ready[r_set], 0));
            for (int q = 0; q < 10; q++) {
                if (!a.out[q] || (ex && ex->skip_row_copy)) continue;
                const size_t stride = q < 8 ? rs : 1;
                const size_t bytes = (size_t)nb * stride * sizeof(T);
                CK(cudaMemcpyAsync(a.out[q] + (size_t)(ub + b0) * stride, d_out[r_set][q].p, bytes, cudaMemcpyDeviceToHost, rsx.s));
                tm.d2h_bytes += (int64_t)bytes;
            }
            if (ex && ex->status) { CK(cudaMemcpyAsync(ex->status + ub + b0, d_stat_out[r_set].p, (size_t)nb * sizeof(int), cudaMemcpyDeviceToHost, rsx.s)); tm.d2h_bytes += (int64_t)nb * 4; }
            if (ex && ex->topk_items) { CK(cudaMemcpyAsync(ex->topk_items + (size_t)(ub + b0) * K, d_tki[r_set].p, (size_t)nb * K * sizeof(int), cudaMemcpyDeviceToHost, rsx.s)); tm.d2h_bytes += (int64_t)nb * K * 4; }
            if (ex && ex->topk_scores) { CK(cudaMemcpyAsync(reinterpret_cast<T*>(ex->topk_scores) + (size_t)(ub + b0) * K, d_tks[r_set].p, (size_t)nb * K * sizeof(T), cudaMemcpyDeviceToHost, rsx.s)); tm.d2h_bytes += (int64_t)nb * K * (int64_t)sizeof(T); }
            CK(cudaEventRecord(rsx.copied[r_set], rsx.s));
            rsx.copied_valid[r_set] = true;
        }
    }
    if (!on_dev) {                                        // what is still exposed of the result copies
        const auto t_d = std::chrono::steady_clock::now();
        CK(cudaStreamSynchronize(rsx.s));
        tm.d2h_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_d).count();
    }
    if (!on_dev && ex && ex->pos_rank && !g_interrupt.load()) {
        pt.start();
        CK(cudaMemcpyAsync(ex->pos_rank + lo_hi[2], pos_rank_d, nnz_te * sizeof(long long), cudaMemcpyDeviceToHost, st));
        tm.d2h_bytes += (int64_t)(nnz_te * sizeof(long long));
        pt.stop(tm.d2h_ms);
    }
    if (want_means && !g_interrupt.load()) {
        double* means_d = nullptr; long long* counts_d = nullptr;
        const size_t cells = (size_t)10 * mean_W;
        if (on_dev) { means_d = ex->metric_means; counts_d = reinterpret_cast<long long*>(ex->metric_counts); }
        else {
            if (ex->metric_means) { CK(d_means.alloc(cells * sizeof(double))); means_d = d_means.as<double>(); }
            if (ex->metric_counts) { CK(d_mcounts.alloc(cells * sizeof(long long))); counts_d = d_mcounts.as<long long>(); }
        }
        pt.start();
        metric_final_kernel<<<(int)((cells + 127) / 128), 128, 0, st>>>(d_psum.as<double>(), d_pcnt.as<int>(), mean_block0, mean_W, means_d, counts_d);
        CK(cudaGetLastError());
        tm.kernel_launches++;
        pt.stop(tm.metrics_ms);
        if (!on_dev) {
            if (ex->metric_means) { CK(cudaMemcpyAsync(ex->metric_means, means_d, cells * sizeof(double), cudaMemcpyDeviceToHost, st)); tm.d2h_bytes += (int64_t)(cells * sizeof(double)); }
            if (ex->metric_counts) { CK(cudaMemcpyAsync(ex->metric_counts, counts_d, cells * sizeof(long long), cudaMemcpyDeviceToHost, st)); tm.d2h_bytes += (int64_t)(cells * sizeof(long long)); }
        }
    }
    if (use_tensor && want_filter_stats) {
        float ratio = 0.f;
        CK(cudaMemcpyAsync(&ratio, d_maxbn.as<unsigned>() + 1, sizeof(float), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        tm.filter_err_ratio_max = (double)ratio;
    }
    CK(cudaStreamSynchronize(st));

    tm.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    tm.devices_used = 1;
    if (ex && ex->timing) *ex->timing = tm;

    if (g_interrupt.load()) {   // hpp:167-173: restore the handler, re-raise, report
        latch.restore();
        if (!nested) std::raise(SIGINT);
        set_err("interrupted", "procedure was interrupted");
        return RMB200_ERR_INTERRUPTED;
    }
    return RMB200_OK;
}

inline const rmb200_extra_t* ex_or_default(const rmb200_extra_t* ex)
{
    static const rmb200_extra_t dflt = []() {
        rmb200_extra_t e;
        std::memset(&e, 0, sizeof(e));
        e.struct_size = (int32_t)sizeof(e);
        e.device = -1;
        return e;
    }();
    return ex ? ex : &dflt;
}

// Which GPUs the call runs on: rmb200_extra_t::devices, else (no explicit `device`) env RMB200_DEVICES = "all" | "0,1,2,...".
// An empty list = the single-device rules of run_call().
int resolve_devices(const rmb200_extra_t* ex, std::vector<int>& devs)
{
    devs.clear();
    if (ex && ex->struct_size != (int32_t)sizeof(rmb200_extra_t)) return RMB200_OK;      // (run_call reports it)
    if (ex && ex->n_devices > 0) {
        if (!ex->devices) { set_err("bad argument", "n_devices > 0 but devices is NULL"); return RMB200_ERR_BAD_ARG; }
        devs.assign(ex->devices, ex->devices + ex->n_devices);
    } else if (!(ex && ex->device >= 0) && !(ex && ex->inputs_on_device)) {
        const char* env = std::getenv("RMB200_DEVICES");
        if (!env || !env[0]) return RMB200_OK;
        if (!std::strcmp(env, "all")) {
            int ndev = 0;
            if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); ndev = 0; }
            for (int d = 0; d < ndev && d < MAX_DEVICES; d++) devs.push_back(d);
        } else {
            for (const char* c = env; *c;) {
                char* end = nullptr;
                const long v = std::strtol(c, &end, 10);
                if (end == c) { set_err("bad argument", "RMB200_DEVICES is neither \"all\" nor a comma-separated list of ordinals"); return RMB200_ERR_BAD_ARG; }
                devs.push_back((int)v);
                c = (*end == ',') ? end + 1 : end;
                if (*end && *end != ',') { set_err("bad argument", "RMB200_DEVICES is neither \"all\" nor a comma-separated list of ordinals"); return RMB200_ERR_BAD_ARG; }
            }
        }
    }
    if (devs.empty()) return RMB200_OK;
    if ((int)devs.size() > MAX_DEVICES) { set_err("bad argument", "more than 16 devices"); return RMB200_ERR_BAD_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        set_err("no CUDA device", "librecometrics_b200 has no CPU fallback");
        return RMB200_ERR_NO_DEVICE;
    }
    for (size_t i = 0; i < devs.size(); i++) {
        if (devs[i] < 0 || devs[i] >= ndev) { set_err("bad argument", "device ordinal out of range"); return RMB200_ERR_BAD_ARG; }
        for (size_t j = 0; j < i; j++)
            if (devs[j] == devs[i]) { set_err("bad argument", "a device is listed twice"); return RMB200_ERR_BAD_ARG; }
    }
    if (devs.size() > 1 && ex && ex->inputs_on_device) { set_err("unsupported", "a multi-GPU call takes host pointers (device-resident inputs live on one GPU)"); return RMB200_ERR_UNSUPPORTED; }
    return RMB200_OK;
}

// /root/reference/src/recometrics.hpp:428-437: the reference's one call spreads the users over every core
// (`#pragma omp parallel for schedule(dynamic) num_threads(nthreads)`); this spreads them over every listed GPU.
template <typename T>
int run_multi(const CallArgs<T>& a0, const std::vector<int>& devs_in)
{
    const auto t_begin = std::chrono::steady_clock::now();
    const rmb200_extra_t* ex = ex_or_default(a0.ex);
    if (a0.m <= 0) { set_err("bad argument", "m, n, k and k_metrics must be positive"); return RMB200_ERR_BAD_ARG; }
    int ub = 0, ue = a0.m;
    if (ex->user_end < 0) return RMB200_OK;
    if (ex->user_begin != 0 || ex->user_end != 0) { ub = ex->user_begin; ue = ex->user_end; }
    if (ub < 0 || ue > a0.m || ub > ue) { set_err("bad argument", "user range outside [0, m]"); return RMB200_ERR_BAD_ARG; }
    // contiguous blocks in units of one CTA's 128 users; devices left without a unit sit the call out
    const int units = (ue - ub + rmb::BM - 1) / rmb::BM;
    std::vector<int> devs, lo, hi;
    {
        const int G0 = (int)devs_in.size();
        for (int g = 0; g < G0; g++) {
            const int u0 = (int)((long long)units * g / G0), u1 = (int)((long long)units * (g + 1) / G0);
            if (u1 <= u0) continue;
            devs.push_back(devs_in[g]);
            lo.push_back(ub + u0 * rmb::BM);
            hi.push_back(std::min(ue, ub + u1 * rmb::BM));
        }
    }
    const int G = (int)devs.size();
    if (G == 0) return RMB200_OK;

    std::unique_lock<std::mutex> lock(g_call_mutex);
    SignalLatch latch;
    MultiCtx mc;
    mc.G = G;
    const bool want_means = ex->metric_means || ex->metric_counts;
    const int W = a0.cumulative ? a0.K : 1;
    const size_t cells = (size_t)10 * (size_t)(W > 0 ? W : 1);
    std::vector<rmb200_extra_t> exs((size_t)G, *ex);
    std::vector<rmb200_timing_t> tms((size_t)G);
    std::vector<std::vector<double>> means((size_t)G);
    std::vector<std::vector<long long>> counts((size_t)G);
    std::vector<int> rcs((size_t)G, RMB200_OK);
    std::vector<std::string> errs((size_t)G);
    std::vector<std::thread> workers;
    for (int g = 0; g < G; g++) {
        std::memset(&tms[g], 0, sizeof(rmb200_timing_t));
        rmb200_extra_t& e = exs[g];
        e.device = devs[g]; e.devices = nullptr; e.n_devices = 0;
        e.user_begin = lo[g]; e.user_end = hi[g];
        e.timing = &tms[g];
        if (want_means) {
            means[g].assign(cells, 0.0); counts[g].assign(cells, 0);
            e.metric_means = means[g].data(); e.metric_counts = reinterpret_cast<int64_t*>(counts[g].data());
        }
        workers.emplace_back([&, g]() {
            g_err.clear();
            g_upload_share = G;
            int rc;
            try {
                // peers' memory directly over NVLink where the hardware allows it (else the copies are staged by the driver)
                if (cudaSetDevice(devs[g]) == cudaSuccess)
                    for (int h = 0; h < G; h++) {
                        int can = 0;
                        if (h != g && cudaDeviceCanAccessPeer(&can, devs[g], devs[h]) == cudaSuccess && can) cudaDeviceEnablePeerAccess(devs[h], 0);
                        cudaGetLastError();
                    }
                CallArgs<T> a = a0;
                a.ex = &exs[g]; a.mc = &mc; a.mc_rank = g;
                rc = run_call<T>(a);
            } catch (const std::bad_alloc&) {
                set_err("host allocation failed"); rc = RMB200_ERR_OOM;
            } catch (...) {
                set_err("unexpected C++ exception"); rc = RMB200_ERR_CUDA;
            }
            rcs[g] = rc; errs[g] = g_err;
            mc.depart(rc != RMB200_OK);
        });
    }
    for (auto& w : workers) w.join();

    int rc = RMB200_OK;
    for (int g = 0; g < G && rc == RMB200_OK; g++)
        if (rcs[g] != RMB200_OK && rcs[g] != RMB200_ERR_INTERRUPTED) { rc = rcs[g]; g_err = errs[g]; }
    if (rc == RMB200_OK && g_interrupt.load()) {
        latch.restore();
        std::raise(SIGINT);
        set_err("interrupted", "procedure was interrupted");
        return RMB200_ERR_INTERRUPTED;
    }
    if (rc != RMB200_OK) return rc;

    if (want_means) {       // means of the call = the devices' means weighted by their user counts
        for (size_t c = 0; c < cells; c++) {
            double s = 0.0; long long cnt = 0;
            for (int g = 0; g < G; g++) if (counts[g][c] > 0) { s += means[g][c] * (double)counts[g][c]; cnt += counts[g][c]; }
            if (ex->metric_means) ex->metric_means[c] = cnt ? s / (double)cnt : std::numeric_limits<double>::quiet_NaN();
            if (ex->metric_counts) ex->metric_counts[c] = cnt;
        }
    }
    if (ex->timing) {
        rmb200_timing_t t = tms[0];
        for (int g = 1; g < G; g++) {
            const rmb200_timing_t& o = tms[g];
            t.h2d_ms = std::max(t.h2d_ms, o.h2d_ms); t.prep_ms = std::max(t.prep_ms, o.prep_ms);
            t.score_select_ms = std::max(t.score_select_ms, o.score_select_ms); t.metrics_ms = std::max(t.metrics_ms, o.metrics_ms);
            t.d2h_ms = std::max(t.d2h_ms, o.d2h_ms); t.dominant_kernel_ms = std::max(t.dominant_kernel_ms, o.dominant_kernel_ms);
            t.kernel_launches += o.kernel_launches; t.h2d_bytes += o.h2d_bytes; t.d2h_bytes += o.d2h_bytes;
            t.filter_fallback_batches += o.filter_fallback_batches; t.filter_retry_rows += o.filter_retry_rows;
            t.filter_fallback_users += o.filter_fallback_users; t.noise_handback_users += o.noise_handback_users;
            t.filter_err_ratio_max = std::max(t.filter_err_ratio_max, o.filter_err_ratio_max);
        }
        t.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
        t.devices_used = G;
        *ex->timing = t;
    }
    return RMB200_OK;
}

template <typename T>
int entry(const T* A, size_t lda, const T* B, size_t ldb, int32_t m, int32_t n, int32_t k,
          const int32_t* trp, const int32_t* tri, const int32_t* tep, const int32_t* tei, const T* tev,
          int32_t K, int cumulative, int noise,
          T* p, T* tp, T* r, T* ap, T* tap, T* ndcg, T* hit, T* rr, T* roc, T* pr,
          int ccs, int32_t mip, int32_t mpt, uint64_t seed, const T* bias, const rmb200_extra_t* ex)
{
    g_err.clear();
    CallArgs<T> a;
    a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.m = m; a.n = n; a.k = k;
    a.trp = trp; a.tri = tri; a.tep = tep; a.tei = tei; a.tev = tev;
    a.K = K; a.cumulative = cumulative ? 1 : 0; a.noise = noise ? 1 : 0;
    T* outs[10] = {p, tp, r, ap, tap, ndcg, hit, rr, roc, pr};
    for (int q = 0; q < 10; q++) a.out[q] = outs[q];
    a.consider_cold_start = ccs ? 1 : 0; a.min_items_pool = mip; a.min_pos_test = mpt; a.seed = seed;
    a.bias = bias; a.ex = ex;
    try {
        std::vector<int> devs;
        int rc = resolve_devices(ex, devs);
        if (rc) return rc;
        if (devs.size() > 1) return run_multi<T>(a, devs);
        if (devs.size() == 1) {                 // a list of one: that device
            rmb200_extra_t one = *ex_or_default(ex);
            one.device = devs[0]; one.devices = nullptr; one.n_devices = 0;
            a.ex = &one;
            return run_call<T>(a);
        }
        return run_call<T>(a);
    } catch (const std::bad_alloc&) {
        set_err("host allocation failed");
        return RMB200_ERR_OOM;
    } catch (...) {
        set_err("unexpected C++ exception");
        return RMB200_ERR_CUDA;
    }
}

}  // namespace

extern "C" {

int rmb200_calc_metrics_f32(
    const float* A, size_t lda, const float* B, size_t ldb, int32_t m, int32_t n, int32_t k,
    const int32_t* Xtrain_csr_p, const int32_t* Xtrain_csr_i,
    const int32_t* Xtest_csr_p, const int32_t* Xtest_csr_i, const float* Xtest_csr,
    int32_t k_metrics, int cumulative, int break_ties_with_noise,
    float* p_at_k, float* tp_at_k, float* r_at_k, float* ap_at_k, float* tap_at_k,
    float* ndcg_at_k, float* hit_at_k, float* rr_at_k, float* roc_auc, float* pr_auc,
    int consider_cold_start, int32_t min_items_pool, int32_t min_pos_test, int32_t nthreads, uint64_t seed)
{
    (void)nthreads;
    return entry<float>(A, lda, B, ldb, m, n, k, Xtrain_csr_p, Xtrain_csr_i, Xtest_csr_p, Xtest_csr_i, Xtest_csr,
                        k_metrics, cumulative, break_ties_with_noise, p_at_k, tp_at_k, r_at_k, ap_at_k, tap_at_k,
                        ndcg_at_k, hit_at_k, rr_at_k, roc_auc, pr_auc, consider_cold_start, min_items_pool,
                        min_pos_test, seed, nullptr, nullptr);
}

int rmb200_calc_metrics_f64(
    const double* A, size_t lda, const double* B, size_t ldb, int32_t m, int32_t n, int32_t k,
    const int32_t* Xtrain_csr_p, const int32_t* Xtrain_csr_i,
    const int32_t* Xtest_csr_p, const int32_t* Xtest_csr_i, const double* Xtest_csr,
    int32_t k_metrics, int cumulative, int break_ties_with_noise,
    double* p_at_k, double* tp_at_k, double* r_at_k, double* ap_at_k, double* tap_at_k,
    double* ndcg_at_k, double* hit_at_k, double* rr_at_k, double* roc_auc, double* pr_auc,
    int consider_cold_start, int32_t min_items_pool, int32_t min_pos_test, int32_t nthreads, uint64_t seed)
{
    (void)nthreads;
    return entry<double>(A, lda, B, ldb, m, n, k, Xtrain_csr_p, Xtrain_csr_i, Xtest_csr_p, Xtest_csr_i, Xtest_csr,
                         k_metrics, cumulative, break_ties_with_noise, p_at_k, tp_at_k, r_at_k, ap_at_k, tap_at_k,
                         ndcg_at_k, hit_at_k, rr_at_k, roc_auc, pr_auc, consider_cold_start, min_items_pool,
                         min_pos_test, seed, nullptr, nullptr);
}

int rmb200_calc_metrics_ex_f32(
    const float* A, size_t lda, const float* B, size_t ldb, int32_t m, int32_t n, int32_t k,
    const int32_t* Xtrain_csr_p, const int32_t* Xtrain_csr_i,
    const int32_t* Xtest_csr_p, const int32_t* Xtest_csr_i, const float* Xtest_csr,
    int32_t k_metrics, int cumulative, int break_ties_with_noise,
    float* p_at_k, float* tp_at_k, float* r_at_k, float* ap_at_k, float* tap_at_k,
    float* ndcg_at_k, float* hit_at_k, float* rr_at_k, float* roc_auc, float* pr_auc,
    int consider_cold_start, int32_t min_items_pool, int32_t min_pos_test, int32_t nthreads, uint64_t seed,
    const float* item_biases, const rmb200_extra_t* extra)
{
    (void)nthreads;
    return entry<float>(A, lda, B, ldb, m, n, k, Xtrain_csr_p, Xtrain_csr_i, Xtest_csr_p, Xtest_csr_i, Xtest_csr,
                        k_metrics, cumulative, break_ties_with_noise, p_at_k, tp_at_k, r_at_k, ap_at_k, tap_at_k,
                        ndcg_at_k, hit_at_k, rr_at_k, roc_auc, pr_auc, consider_cold_start, min_items_pool,
                        min_pos_test, seed, item_biases, extra);
}

int rmb200_calc_metrics_ex_f64(
    const double* A, size_t lda, const double* B, size_t ldb, int32_t m, int32_t n, int32_t k,
    const int32_t* Xtrain_csr_p, const int32_t* Xtrain_csr_i,
    const int32_t* Xtest_csr_p, const int32_t* Xtest_csr_i, const double* Xtest_csr,
    int32_t k_metrics, int cumulative, int break_ties_with_noise,
    double* p_at_k, double* tp_at_k, double* r_at_k, double* ap_at_k, double* tap_at_k,
    double* ndcg_at_k, double* hit_at_k, double* rr_at_k, double* roc_auc, double* pr_auc,
    int consider_cold_start, int32_t min_items_pool, int32_t min_pos_test, int32_t nthreads, uint64_t seed,
    const double* item_biases, const rmb200_extra_t* extra)
{
    (void)nthreads;
    return entry<double>(A, lda, B, ldb, m, n, k, Xtrain_csr_p, Xtrain_csr_i, Xtest_csr_p, Xtest_csr_i, Xtest_csr,
                         k_metrics, cumulative, break_ties_with_noise, p_at_k, tp_at_k, r_at_k, ap_at_k, tap_at_k,
                         ndcg_at_k, hit_at_k, rr_at_k, roc_auc, pr_auc, consider_cold_start, min_items_pool,
                         min_pos_test, seed, item_biases, extra);
}

int rmb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int rmb200_version(void) { return RMB200_VERSION; }

int rmb200_sizeof_extra(void) { return (int)sizeof(rmb200_extra_t); }

int rmb200_sizeof_timing(void) { return (int)sizeof(rmb200_timing_t); }

const char* rmb200_last_error(void) { return g_err.c_str(); }

void rmb200_request_interrupt(void) { g_interrupt.store(1); }

void rmb200_release_workspace(void)
{
    std::lock_guard<std::mutex> call(g_call_mutex);      // not while a call is using them
    upload_pool_release_all();
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    pool_trim_locked(-1);
}

double rmb200_measure_fma_peak(int device, int dtype_bytes, double* elapsed_ms)
{
    g_err.clear();
    int ndev = rmb200_device_count();
    if (ndev <= 0) { set_err("no CUDA device"); return -1.0; }
    if (device < 0) cudaGetDevice(&device);
    if (device >= ndev || cudaSetDevice(device) != cudaSuccess) { set_err("bad device"); return -1.0; }
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    const int threads = 256, blocks = nsm * 8;
    const int iters = dtype_bytes == 4 ? 40000 : 20000;
    DevBuf out;
    if (out.alloc((size_t)blocks * threads * 8) != cudaSuccess) { set_err("alloc"); return -1.0; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        if (dtype_bytes == 4) rmb::fma2_peak_kernel<<<blocks, threads>>>(out.as<float>(), iters, 1.0f);   // 4 x 16 FFMA2 = 8 x 16 FMA per iteration
        else rmb::fma_peak_kernel<double><<<blocks, threads>>>(out.as<double>(), iters, 1.0);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { set_err("fma peak kernel failed"); cudaGetLastError(); return -1.0; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (elapsed_ms) *elapsed_ms = best;
    const double flops = (double)blocks * threads * (double)iters * 8.0 * 16.0 * 2.0;
    return flops / (best * 1e-3) / 1e12;
}

}  // extern "C"
