// split.cu -- the train/test splitters behind rmb200_split_* (include/recometrics_b200.h).
//
// Stands where /root/reference/src/recometrics.hpp:1015-1505 stands: split_data_selected_users (every row of X split
// into a training and a held-out part), split_data_separate_users (a random sample of eligible users split that way, the
// other users returned untouched) and split_data_joined_users (the same with the other users appended below the training
// rows).
//
// What decides the result is ONE std::mt19937 stream consumed row after row by std::shuffle (hpp:1055-1060), plus one
// shuffle of the user ids (hpp:1223-1226): a row's draws start where the previous row's ended, and how many a row takes
// depends on the rejections of libstdc++'s bounded-integer method.  That replay is sequential by the definition of the
// output; plan_rows() below does it on the host, on index arrays that never leave L1/L2, and emits ONE BYTE per entry
// ("held out" or not).  It runs while a second host thread moves X to the GPU.
//
// All work on the matrix itself is on the GPU, phrased per ENTRY rather than per row so that a catalogue's power-law row
// lengths do not matter:
//   * the reference leaves each half of a split row ordered by item id (std::sort, hpp:1064-1066, :1075-1077).  With rows
//     that arrive sorted (what the reference's Python and R fronts guarantee) this is a STABLE PARTITION of the row; and
//     since the held-out entries of rows 0..r-1 fill exactly test_p[r] slots, the stable partition of all rows at once is
//     one global exclusive scan of the bytes: held-out entry j goes to slot scan[j] of the test arrays, any other entry to
//     slot j - scan[j] of the training arrays.  No per-row loop, no row lookup, coalesced reads.
//   * rows that arrive unsorted are first ordered by one segmented radix sort of (item id, position) pairs; rows that the
//     reference copies verbatim (nothing or everything held out, hpp:1043-1054) keep their input order.
//   * the rows of the users a split leaves alone (hpp:1303-1321) and the sub-matrix of the sampled users (hpp:1271-1284)
//     are gathers through a row list; the sub-matrix is never materialised -- the scatter reads X through the list.
#include "../../include/recometrics_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <exception>
#include <new>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_radix_sort.cuh>
#include <thrust/iterator/transform_iterator.h>

namespace rmb {
void set_last_error(const char* what, const char* detail);   // api.cu
}

namespace {

using clk = std::chrono::steady_clock;
double ms_since(clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); }

// ------------------------------------------------------------------------------------------------------------------
// Host: the replay of the reference's random stream
// ------------------------------------------------------------------------------------------------------------------
struct SplitPlan {
    bool whole = true;                 // every row of X is split, in place (split_data_selected_users)
    std::vector<int32_t> sel_rows;     // rows of X that are split, ascending (empty when `whole`)
    std::vector<int32_t> rem_rows;     // the other rows, ascending
    std::vector<int32_t> sel_p;        // [ns+1] the split rows as a matrix of their own (when `whole`: X's own pointer)
    std::vector<int32_t> test_p;       // [ns+1]
    std::vector<int32_t> train_p;      // [ns+1] (+ the remainder's rows when joined)
    std::vector<int32_t> rem_p;        // [nr+1]
    std::vector<uint8_t> held;         // [nnz of the split rows] 1 = held out
    int32_t ns = 0, nr = 0;
};

// /root/reference/src/recometrics.hpp:1015-1106 without the data movement: the held-out count of every row (:1037-1040)
// and, for rows with something on both sides, the positions std::shuffle puts first (:1057-1060).
void plan_rows(const int32_t* Xp, const int32_t* rows, int32_t ns, double test_fraction, uint64_t seed, SplitPlan& P)
{
    P.ns = ns;
    P.sel_p.resize((size_t)ns + 1);
    P.test_p.resize((size_t)ns + 1);
    P.sel_p[0] = 0;
    P.test_p[0] = 0;
    int32_t longest = 0;
    for (int32_t r = 0; r < ns; r++) {
        const int32_t u = rows ? rows[r] : r;
        const int32_t cnt = Xp[u + 1] - Xp[u];
        P.sel_p[r + 1] = P.sel_p[r] + cnt;
        P.test_p[r + 1] = P.test_p[r] + (int32_t)std::round(cnt * test_fraction);
        longest = std::max(longest, cnt);
    }
    P.train_p.resize((size_t)ns + 1);
    for (int32_t r = 0; r <= ns; r++) P.train_p[r] = P.sel_p[r] - P.test_p[r];

    P.held.assign((size_t)P.sel_p[ns], 0);
    std::mt19937 rng(seed);
    std::vector<int32_t> order((size_t)longest);
    for (int32_t r = 0; r < ns; r++) {
        const int32_t cnt = P.sel_p[r + 1] - P.sel_p[r];
        const int32_t out = P.test_p[r + 1] - P.test_p[r];
        if (!cnt || !out) continue;
        uint8_t* mark = P.held.data() + P.sel_p[r];
        if (out == cnt) { std::memset(mark, 1, (size_t)cnt); continue; }
        std::iota(order.begin(), order.begin() + cnt, (int32_t)0);
        std::shuffle(order.begin(), order.begin() + cnt, rng);
        for (int32_t j = 0; j < out; j++) mark[order[j]] = 1;
    }
}

// /root/reference/src/recometrics.hpp:1223-1269: which users are split.  Returns the reference's error text or nullptr.
const char* pick_users(const int32_t* Xp, int32_t m, int32_t n, int32_t n_users_test, double test_fraction,
                       bool consider_cold_start, int32_t min_items_pool, int32_t min_pos_test, uint64_t seed, SplitPlan& P)
{
    if (n_users_test > m) return "Target number of test users is larger than available users.\n";
    if (min_items_pool >= n) return "Selected minimum number of items is larger than total number of items.\n";
    std::vector<int32_t> ids((size_t)m);
    std::iota(ids.begin(), ids.end(), (int32_t)0);
    std::mt19937 rng(seed);
    std::shuffle(ids.begin(), ids.end(), rng);

    auto eligible = [&](int32_t u) {
        const int32_t cnt = Xp[u + 1] - Xp[u];
        if (!cnt) return false;
        const int32_t out = (int32_t)std::round(cnt * test_fraction);
        if (out < min_pos_test) return false;
        if (n - (cnt - out) < min_items_pool) return false;
        if (!consider_cold_start && out == cnt) return false;
        return cnt + 1 < n;
    };
    int32_t taken = 0, end = m;
    do {   // (a do-while in the reference as well: the first candidate is looked at even for n_users_test = 0)
        if (eligible(ids[taken])) taken++;
        else std::swap(ids[taken], ids[--end]);
    } while (taken < n_users_test && taken < end);
    if (!taken) return "No users satisfy criteria for test inclusion.\n";

    std::sort(ids.begin(), ids.begin() + taken);
    std::sort(ids.begin() + taken, ids.end());
    P.whole = false;
    P.sel_rows.assign(ids.begin(), ids.begin() + taken);
    P.rem_rows.assign(ids.begin() + taken, ids.end());
    P.nr = m - taken;
    P.rem_p.resize((size_t)P.nr + 1);
    P.rem_p[0] = 0;
    for (int32_t r = 0; r < P.nr; r++) P.rem_p[r + 1] = P.rem_p[r] + (Xp[P.rem_rows[r] + 1] - Xp[P.rem_rows[r]]);
    return nullptr;
}

// ------------------------------------------------------------------------------------------------------------------
// Device
// ------------------------------------------------------------------------------------------------------------------
constexpr int TPB = 256;

// last row r with ptr[r] <= j (rows may be empty)
__device__ __forceinline__ int row_of(const int* __restrict__ ptr, int rows, int j)
{
    int lo = 0, hi = rows;          // invariant: ptr[lo] <= j < ptr[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(ptr + mid) <= j) lo = mid; else hi = mid;
    }
    return lo;
}

struct SplitView {
    const int* Xp; const int* Xi;
    const int* sel_rows;            // nullptr: the split rows are X's rows
    const int* sel_p; const int* test_p;
    int ns; int total;              // entries of the split rows
};

// entry of X behind position j of the split rows' own matrix
__device__ __forceinline__ int source_entry(const SplitView& V, int r, int j)
{
    return V.sel_rows ? __ldg(V.Xp + __ldg(V.sel_rows + r)) + (j - __ldg(V.sel_p + r)) : j;
}

// rows the reference reorders: something held out AND something kept (hpp:1043-1054 copies the others verbatim)
__device__ __forceinline__ bool row_is_split(const SplitView& V, int r)
{
    const int cnt = __ldg(V.sel_p + r + 1) - __ldg(V.sel_p + r);
    const int out = __ldg(V.test_p + r + 1) - __ldg(V.test_p + r);
    return out > 0 && out < cnt;
}

// 1 into *unsorted when a split row has a descending pair of item ids
__global__ void __launch_bounds__(TPB) check_sorted_kernel(SplitView V, int* __restrict__ unsorted)
{
    for (long long jj = (long long)blockIdx.x * TPB + threadIdx.x; jj < V.total; jj += (long long)gridDim.x * TPB) {
        const int j = (int)jj;
        const int r = row_of(V.sel_p, V.ns, j);
        if (j == __ldg(V.sel_p + r) || !row_is_split(V, r)) continue;
        const int e = source_entry(V, r, j);
        if (__ldg(V.Xi + e) < __ldg(V.Xi + e - 1)) *unsorted = 1;
    }
}

// sort input of the unsorted case: key = item id, value = position in the split rows' matrix
__global__ void __launch_bounds__(TPB) sort_input_kernel(SplitView V, int* __restrict__ keys, int* __restrict__ vals)
{
    for (long long jj = (long long)blockIdx.x * TPB + threadIdx.x; jj < V.total; jj += (long long)gridDim.x * TPB) {
        const int j = (int)jj;
        const int r = row_of(V.sel_p, V.ns, j);
        keys[j] = __ldg(V.Xi + source_entry(V, r, j));
        vals[j] = j;
    }
}

// position j of the ORDERED split rows -> its position before ordering
__device__ __forceinline__ int unordered_pos(const SplitView& V, const int* __restrict__ perm, int j, int* row)
{
    if (!perm && !V.sel_rows) { *row = -1; return j; }
    const int r = row_of(V.sel_p, V.ns, j);
    *row = r;
    return (perm && row_is_split(V, r)) ? __ldg(perm + j) : j;
}

// the held-out bytes in the order the entries will be written
__global__ void __launch_bounds__(TPB) ordered_flags_kernel(SplitView V, const int* __restrict__ perm,
                                                            const uint8_t* __restrict__ held, uint8_t* __restrict__ flags)
{
    for (long long jj = (long long)blockIdx.x * TPB + threadIdx.x; jj < V.total; jj += (long long)gridDim.x * TPB) {
        const int j = (int)jj;
        int r;
        flags[j] = held[unordered_pos(V, perm, j, &r)];
    }
}

// the stable partition of every row at once: scan[j] = held-out entries before j
template <typename T>
__global__ void __launch_bounds__(TPB) partition_kernel(SplitView V, const T* __restrict__ Xv, const int* __restrict__ perm,
                                                        const uint8_t* __restrict__ flags, const int* __restrict__ scan,
                                                        int* __restrict__ train_i, T* __restrict__ train_v,
                                                        int* __restrict__ test_i, T* __restrict__ test_v)
{
    for (long long jj = (long long)blockIdx.x * TPB + threadIdx.x; jj < V.total; jj += (long long)gridDim.x * TPB) {
        const int j = (int)jj;
        int r;
        const int s = unordered_pos(V, perm, j, &r);
        const int e = (r >= 0) ? source_entry(V, r, s) : s;
        const int item = __ldg(V.Xi + e);
        const T val = __ldg(Xv + e);
        const int before = scan[j];
        if (flags[j]) { test_i[before] = item; test_v[before] = val; }
        else { train_i[j - before] = item; train_v[j - before] = val; }
    }
}

// rows of X listed in `rows`, copied one below the other (hpp:1303-1321)
template <typename T>
__global__ void __launch_bounds__(TPB) gather_rows_kernel(const int* __restrict__ Xp, const int* __restrict__ Xi,
                                                          const T* __restrict__ Xv, const int* __restrict__ rows,
                                                          const int* __restrict__ out_p, int n_rows, int total,
                                                          int* __restrict__ out_i, T* __restrict__ out_v)
{
    for (long long jj = (long long)blockIdx.x * TPB + threadIdx.x; jj < total; jj += (long long)gridDim.x * TPB) {
        const int j = (int)jj;
        const int r = row_of(out_p, n_rows, j);
        const int e = __ldg(Xp + __ldg(rows + r)) + (j - __ldg(out_p + r));
        out_i[j] = __ldg(Xi + e);
        out_v[j] = __ldg(Xv + e);
    }
}

struct ByteToInt {
    __host__ __device__ __forceinline__ int operator()(const uint8_t& b) const { return (int)b; }
};

struct DevMem {
    void* p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    ~DevMem() { if (p) cudaFree(p); }
    template <typename U> U* as() const { return (U*)p; }
};

struct Streams {
    cudaStream_t up = nullptr, st = nullptr;
    ~Streams() { if (up) cudaStreamDestroy(up); if (st) cudaStreamDestroy(st); }
};
struct Events {
    cudaEvent_t a = nullptr, b = nullptr;
    ~Events() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};
struct Joiner {
    std::thread& t;
    ~Joiner() { if (t.joinable()) t.join(); }
};

// what rmb200_split_t::owner points to
struct SplitOwner { std::vector<void*> blocks; };

void* host_block(SplitOwner* own, size_t bytes)
{
    void* p = std::malloc(bytes ? bytes : 1);
    if (p) own->blocks.push_back(p);
    return p;
}

int fail(int code, const char* what, const char* detail = nullptr)
{
    rmb::set_last_error(what, detail);
    return code;
}

#define SPLIT_CUDA(call)                                                                                       \
    do {                                                                                                       \
        const cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                               \
            cudaGetLastError();                                                                                \
            rmb200_split_free(out);                                                                            \
            return fail(e_ == cudaErrorMemoryAllocation ? RMB200_ERR_OOM : RMB200_ERR_CUDA, #call, cudaGetErrorString(e_)); \
        }                                                                                                      \
    } while (0)

enum SplitKind { SPLIT_WHOLE = 0, SPLIT_SEPARATE = 1, SPLIT_JOINED = 2 };

template <typename T>
int run_split_impl(SplitKind kind, const int32_t* Xp, const int32_t* Xi, const T* Xv, int32_t m, int32_t n,
              int32_t n_users_test, double test_fraction, bool consider_cold_start, int32_t min_items_pool,
              int32_t min_pos_test, uint64_t seed, int32_t device, rmb200_split_t* out)
{
    rmb::set_last_error("", nullptr);
    if (!out) return fail(RMB200_ERR_BAD_ARG, "rmb200_split: out is NULL");
    std::memset(out, 0, sizeof(*out));
    out->value_bytes = (int32_t)sizeof(T);
    const auto t_call = clk::now();

    // the reference's own argument checks come first, in its order (hpp:1033-1036, :1225-1228)
    if (kind == SPLIT_WHOLE) {
        if (!m) return RMB200_OK;          // hpp:1033: nothing is written, every output stays empty
        if (m < 0 || n < 0) return fail(RMB200_ERR_RUNTIME, "Passed negative dimensions.\n");
    } else {
        if (n_users_test > m) return fail(RMB200_ERR_RUNTIME, "Target number of test users is larger than available users.\n");
        if (min_items_pool >= n) return fail(RMB200_ERR_RUNTIME, "Selected minimum number of items is larger than total number of items.\n");
        if (m <= 0) return fail(RMB200_ERR_BAD_ARG, "rmb200_split: X has no rows");
    }
    if (!Xp) return fail(RMB200_ERR_BAD_ARG, "rmb200_split: X_csr_p is NULL");
    const int64_t nnz = Xp[m];
    if (Xp[0] != 0) return fail(RMB200_ERR_BAD_ARG, "rmb200_split: X_csr_p[0] must be 0");
    if (nnz < 0 || (nnz > 0 && (!Xi || !Xv))) return fail(RMB200_ERR_BAD_ARG, "rmb200_split: X_csr_i / X_csr missing");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(RMB200_ERR_NO_DEVICE, "no usable CUDA device (the splitters have no CPU path)");
    }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= ndev) return fail(RMB200_ERR_BAD_ARG, "rmb200_split: no such device");
    out->device = device;
    SPLIT_CUDA(cudaSetDevice(device));

    // X on its way to the GPU on a second thread while this one replays the random stream
    DevMem dXp, dXi, dXv;
    SPLIT_CUDA(dXp.alloc(sizeof(int32_t) * ((size_t)m + 1)));
    SPLIT_CUDA(dXi.alloc(sizeof(int32_t) * (size_t)nnz));
    SPLIT_CUDA(dXv.alloc(sizeof(T) * (size_t)nnz));
    Streams streams;
    SPLIT_CUDA(cudaStreamCreateWithFlags(&streams.up, cudaStreamNonBlocking));
    SPLIT_CUDA(cudaStreamCreateWithFlags(&streams.st, cudaStreamNonBlocking));
    const cudaStream_t st_up = streams.up, st = streams.st;
    cudaError_t up_err = cudaSuccess;
    double up_ms = 0.0;
    std::thread uploader([&]() {
        const auto t0 = clk::now();
        cudaSetDevice(device);
        cudaError_t e = cudaMemcpyAsync(dXp.p, Xp, sizeof(int32_t) * ((size_t)m + 1), cudaMemcpyHostToDevice, st_up);
        if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(dXi.p, Xi, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, st_up);
        if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(dXv.p, Xv, sizeof(T) * (size_t)nnz, cudaMemcpyHostToDevice, st_up);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st_up);
        up_err = e;
        up_ms = ms_since(t0);
    });
    Joiner joiner{uploader};

    const auto t_plan = clk::now();
    SplitPlan P;
    const char* refusal = nullptr;
    if (kind != SPLIT_WHOLE)
        refusal = pick_users(Xp, m, n, n_users_test, test_fraction, consider_cold_start, min_items_pool, min_pos_test, seed, P);
    if (!refusal) plan_rows(Xp, P.whole ? nullptr : P.sel_rows.data(), P.whole ? m : (int32_t)P.sel_rows.size(), test_fraction, seed, P);
    out->plan_ms = ms_since(t_plan);
    uploader.join();
    if (refusal) return fail(RMB200_ERR_RUNTIME, refusal);
    if (up_err != cudaSuccess) { cudaGetLastError(); return fail(RMB200_ERR_CUDA, "upload of X", cudaGetErrorString(up_err)); }
    out->h2d_ms = up_ms;
    out->h2d_bytes = (int64_t)(sizeof(int32_t) * ((size_t)m + 1) + (sizeof(int32_t) + sizeof(T)) * (size_t)nnz);

    const int32_t ns = P.ns, nr = P.nr;
    const int64_t sel_nnz = P.sel_p[ns], test_nnz = P.test_p[ns], rem_nnz = nr ? P.rem_p[nr] : 0;
    const int64_t train_nnz = sel_nnz - test_nnz + (kind == SPLIT_JOINED ? rem_nnz : 0);

    // ---- the result (host side): pointer arrays come straight from the plan ----
    SplitOwner* own = new (std::nothrow) SplitOwner();
    if (!own) return fail(RMB200_ERR_OOM, "host allocation");
    out->owner = own;
    auto csr_alloc = [&](rmb200_csr_t& M, int32_t rows, int64_t count) {
        M.rows = rows; M.cols = n; M.nnz = count;
        M.indptr = (int32_t*)host_block(own, sizeof(int32_t) * ((size_t)rows + 1));
        M.indices = (int32_t*)host_block(own, sizeof(int32_t) * (size_t)count);
        M.values = host_block(own, sizeof(T) * (size_t)count);
        return M.indptr && M.indices && M.values;
    };
    bool ok = csr_alloc(out->test, ns, test_nnz) && csr_alloc(out->train, ns + (kind == SPLIT_JOINED ? nr : 0), train_nnz);
    if (ok && kind == SPLIT_SEPARATE) ok = csr_alloc(out->rem, nr, rem_nnz);
    if (ok && kind != SPLIT_WHOLE) {
        out->users_test = (int32_t*)host_block(own, sizeof(int32_t) * (size_t)ns);
        ok = out->users_test != nullptr;
    }
    if (!ok) { rmb200_split_free(out); return fail(RMB200_ERR_OOM, "host allocation of the split"); }
    std::memcpy(out->test.indptr, P.test_p.data(), sizeof(int32_t) * ((size_t)ns + 1));
    std::memcpy(out->train.indptr, P.train_p.data(), sizeof(int32_t) * ((size_t)ns + 1));
    if (kind == SPLIT_JOINED)       // concat_csr_matrices, hpp:1324-1359: the remainder's pointer shifted by the training entries above it
        for (int32_t r = 1; r <= nr; r++) out->train.indptr[ns + r] = P.train_p[ns] + P.rem_p[r];
    if (kind == SPLIT_SEPARATE) std::memcpy(out->rem.indptr, P.rem_p.data(), sizeof(int32_t) * ((size_t)nr + 1));
    if (kind != SPLIT_WHOLE) {
        std::memcpy(out->users_test, P.sel_rows.data(), sizeof(int32_t) * (size_t)ns);
        out->n_users_test = ns;
    }

    // ---- device: plan arrays up, kernels, results down ----
    const auto t_h2d2 = clk::now();
    DevMem d_selrows, d_selp, d_testp, d_remrows, d_remp, d_held, d_flags, d_scan, d_tri, d_trv, d_tei, d_tev, d_unsorted;
    if (!P.whole) {
        SPLIT_CUDA(d_selrows.alloc(sizeof(int32_t) * (size_t)ns));
        SPLIT_CUDA(cudaMemcpyAsync(d_selrows.p, P.sel_rows.data(), sizeof(int32_t) * (size_t)ns, cudaMemcpyHostToDevice, st));
        SPLIT_CUDA(d_selp.alloc(sizeof(int32_t) * ((size_t)ns + 1)));
        SPLIT_CUDA(cudaMemcpyAsync(d_selp.p, P.sel_p.data(), sizeof(int32_t) * ((size_t)ns + 1), cudaMemcpyHostToDevice, st));
        if (nr) {
            SPLIT_CUDA(d_remrows.alloc(sizeof(int32_t) * (size_t)nr));
            SPLIT_CUDA(cudaMemcpyAsync(d_remrows.p, P.rem_rows.data(), sizeof(int32_t) * (size_t)nr, cudaMemcpyHostToDevice, st));
            SPLIT_CUDA(d_remp.alloc(sizeof(int32_t) * ((size_t)nr + 1)));
            SPLIT_CUDA(cudaMemcpyAsync(d_remp.p, P.rem_p.data(), sizeof(int32_t) * ((size_t)nr + 1), cudaMemcpyHostToDevice, st));
        }
    }
    SPLIT_CUDA(d_testp.alloc(sizeof(int32_t) * ((size_t)ns + 1)));
    SPLIT_CUDA(cudaMemcpyAsync(d_testp.p, P.test_p.data(), sizeof(int32_t) * ((size_t)ns + 1), cudaMemcpyHostToDevice, st));
    SPLIT_CUDA(d_held.alloc((size_t)sel_nnz));
    if (sel_nnz) SPLIT_CUDA(cudaMemcpyAsync(d_held.p, P.held.data(), (size_t)sel_nnz, cudaMemcpyHostToDevice, st));
    SPLIT_CUDA(d_flags.alloc((size_t)sel_nnz));
    SPLIT_CUDA(d_scan.alloc(sizeof(int32_t) * (size_t)sel_nnz));
    SPLIT_CUDA(d_tri.alloc(sizeof(int32_t) * (size_t)train_nnz));
    SPLIT_CUDA(d_trv.alloc(sizeof(T) * (size_t)train_nnz));
    SPLIT_CUDA(d_tei.alloc(sizeof(int32_t) * (size_t)test_nnz));
    SPLIT_CUDA(d_tev.alloc(sizeof(T) * (size_t)test_nnz));
    SPLIT_CUDA(d_unsorted.alloc(sizeof(int)));
    SPLIT_CUDA(cudaMemsetAsync(d_unsorted.p, 0, sizeof(int), st));
    SPLIT_CUDA(cudaStreamSynchronize(st));
    out->h2d_ms += ms_since(t_h2d2);
    out->h2d_bytes += (int64_t)sel_nnz + (int64_t)sizeof(int32_t) * (3 * (int64_t)ns + 2 * (int64_t)nr + 4);

    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    auto grid_for = [&](int64_t work) { return (int)std::max<int64_t>(1, std::min<int64_t>((work + TPB - 1) / TPB, (int64_t)nsm * 16)); };

    Events events;
    SPLIT_CUDA(cudaEventCreate(&events.a));
    SPLIT_CUDA(cudaEventCreate(&events.b));
    const cudaEvent_t ev0 = events.a, ev1 = events.b;
    SPLIT_CUDA(cudaEventRecord(ev0, st));
    SplitView V;
    V.Xp = dXp.as<int>(); V.Xi = dXi.as<int>();
    V.sel_rows = P.whole ? nullptr : d_selrows.as<int>();
    V.sel_p = P.whole ? dXp.as<int>() : d_selp.as<int>();
    V.test_p = d_testp.as<int>();
    V.ns = ns; V.total = (int)sel_nnz;

    DevMem d_keys0, d_keys1, d_vals0, d_perm, d_temp;
    const int* perm = nullptr;
    if (sel_nnz) {
        check_sorted_kernel<<<grid_for(sel_nnz), TPB, 0, st>>>(V, d_unsorted.as<int>());
        out->kernel_launches++;
        int unsorted = 0;
        SPLIT_CUDA(cudaMemcpyAsync(&unsorted, d_unsorted.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        SPLIT_CUDA(cudaStreamSynchronize(st));
        if (unsorted) {
            out->rows_sorted_on_device = 1;
            SPLIT_CUDA(d_keys0.alloc(sizeof(int) * (size_t)sel_nnz));
            SPLIT_CUDA(d_keys1.alloc(sizeof(int) * (size_t)sel_nnz));
            SPLIT_CUDA(d_vals0.alloc(sizeof(int) * (size_t)sel_nnz));
            SPLIT_CUDA(d_perm.alloc(sizeof(int) * (size_t)sel_nnz));
            sort_input_kernel<<<grid_for(sel_nnz), TPB, 0, st>>>(V, d_keys0.as<int>(), d_vals0.as<int>());
            out->kernel_launches++;
            size_t temp_bytes = 0;
            SPLIT_CUDA(cub::DeviceSegmentedRadixSort::SortPairs(nullptr, temp_bytes, (const int*)d_keys0.as<int>(), d_keys1.as<int>(),
                                                                (const int*)d_vals0.as<int>(), d_perm.as<int>(), (int)sel_nnz, ns,
                                                                V.sel_p, V.sel_p + 1, 0, 32, st));
            SPLIT_CUDA(d_temp.alloc(temp_bytes));
            SPLIT_CUDA(cub::DeviceSegmentedRadixSort::SortPairs(d_temp.p, temp_bytes, (const int*)d_keys0.as<int>(), d_keys1.as<int>(),
                                                                (const int*)d_vals0.as<int>(), d_perm.as<int>(), (int)sel_nnz, ns,
                                                                V.sel_p, V.sel_p + 1, 0, 32, st));
            out->kernel_launches += 2;     // (cub's segmented sort: at least its partition + sort kernels)
            perm = d_perm.as<int>();
        }
        ordered_flags_kernel<<<grid_for(sel_nnz), TPB, 0, st>>>(V, perm, d_held.as<uint8_t>(), d_flags.as<uint8_t>());
        out->kernel_launches++;
        DevMem d_scan_temp;
        size_t scan_bytes = 0;
        auto as_int = thrust::make_transform_iterator((const uint8_t*)d_flags.as<uint8_t>(), ByteToInt());
        SPLIT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, as_int, d_scan.as<int>(), (int)sel_nnz, st));
        SPLIT_CUDA(d_scan_temp.alloc(scan_bytes));
        SPLIT_CUDA(cub::DeviceScan::ExclusiveSum(d_scan_temp.p, scan_bytes, as_int, d_scan.as<int>(), (int)sel_nnz, st));
        out->kernel_launches += 2;
        partition_kernel<T><<<grid_for(sel_nnz), TPB, 0, st>>>(V, dXv.as<T>(), perm, d_flags.as<uint8_t>(), d_scan.as<int>(),
                                                                d_tri.as<int>(), d_trv.as<T>(), d_tei.as<int>(), d_tev.as<T>());
        out->kernel_launches++;
        SPLIT_CUDA(cudaStreamSynchronize(st));      // (d_scan_temp goes out of scope here)
    }
    DevMem d_rei, d_rev;
    if (rem_nnz) {
        int* dst_i; T* dst_v;
        if (kind == SPLIT_JOINED) {
            dst_i = d_tri.as<int>() + (sel_nnz - test_nnz);
            dst_v = d_trv.as<T>() + (sel_nnz - test_nnz);
        } else {
            SPLIT_CUDA(d_rei.alloc(sizeof(int32_t) * (size_t)rem_nnz));
            SPLIT_CUDA(d_rev.alloc(sizeof(T) * (size_t)rem_nnz));
            dst_i = d_rei.as<int>(); dst_v = d_rev.as<T>();
        }
        gather_rows_kernel<T><<<grid_for(rem_nnz), TPB, 0, st>>>(dXp.as<int>(), dXi.as<int>(), dXv.as<T>(), d_remrows.as<int>(),
                                                                  d_remp.as<int>(), nr, (int)rem_nnz, dst_i, dst_v);
        out->kernel_launches++;
    }
    SPLIT_CUDA(cudaGetLastError());
    SPLIT_CUDA(cudaEventRecord(ev1, st));
    SPLIT_CUDA(cudaEventSynchronize(ev1));
    float kms = 0.f;
    cudaEventElapsedTime(&kms, ev0, ev1);
    out->kernel_ms = kms;

    const auto t_d2h = clk::now();
    if (train_nnz) {
        SPLIT_CUDA(cudaMemcpyAsync(out->train.indices, d_tri.p, sizeof(int32_t) * (size_t)train_nnz, cudaMemcpyDeviceToHost, st));
        SPLIT_CUDA(cudaMemcpyAsync(out->train.values, d_trv.p, sizeof(T) * (size_t)train_nnz, cudaMemcpyDeviceToHost, st));
    }
    if (test_nnz) {
        SPLIT_CUDA(cudaMemcpyAsync(out->test.indices, d_tei.p, sizeof(int32_t) * (size_t)test_nnz, cudaMemcpyDeviceToHost, st));
        SPLIT_CUDA(cudaMemcpyAsync(out->test.values, d_tev.p, sizeof(T) * (size_t)test_nnz, cudaMemcpyDeviceToHost, st));
    }
    if (kind == SPLIT_SEPARATE && rem_nnz) {
        SPLIT_CUDA(cudaMemcpyAsync(out->rem.indices, d_rei.p, sizeof(int32_t) * (size_t)rem_nnz, cudaMemcpyDeviceToHost, st));
        SPLIT_CUDA(cudaMemcpyAsync(out->rem.values, d_rev.p, sizeof(T) * (size_t)rem_nnz, cudaMemcpyDeviceToHost, st));
    }
    SPLIT_CUDA(cudaStreamSynchronize(st));
    out->d2h_ms = ms_since(t_d2h);
    out->d2h_bytes = (int64_t)(sizeof(int32_t) + sizeof(T)) * (train_nnz + test_nnz + (kind == SPLIT_SEPARATE ? rem_nnz : 0));
    out->total_ms = ms_since(t_call);
    return RMB200_OK;
}

// nothing may unwind through the C boundary (std::vector / std::thread can throw)
template <typename T, typename... Args>
int run_split(rmb200_split_t* out, Args... args)
{
    try {
        return run_split_impl<T>(args..., out);
    } catch (const std::bad_alloc&) {
        if (out) rmb200_split_free(out);
        return fail(RMB200_ERR_OOM, "host allocation failed during the split");
    } catch (const std::exception& e) {
        if (out) rmb200_split_free(out);
        return fail(RMB200_ERR_CUDA, "rmb200_split", e.what());
    }
}

}  // namespace

extern "C" {

int rmb200_split_selected_users_f32(const int32_t* Xp, const int32_t* Xi, const float* Xv, int32_t m, int32_t n,
                                    double test_fraction, uint64_t seed, int32_t device, rmb200_split_t* out)
{
    return run_split<float>(out, SPLIT_WHOLE, Xp, Xi, Xv, m, n, 0, test_fraction, false, 0, 0, seed, device);
}

int rmb200_split_selected_users_f64(const int32_t* Xp, const int32_t* Xi, const double* Xv, int32_t m, int32_t n,
                                    double test_fraction, uint64_t seed, int32_t device, rmb200_split_t* out)
{
    return run_split<double>(out, SPLIT_WHOLE, Xp, Xi, Xv, m, n, 0, test_fraction, false, 0, 0, seed, device);
}

int rmb200_split_separate_users_f32(const int32_t* Xp, const int32_t* Xi, const float* Xv, int32_t m, int32_t n,
                                    int32_t n_users_test, double test_fraction, int consider_cold_start, int32_t min_items_pool,
                                    int32_t min_pos_test, uint64_t seed, int32_t device, rmb200_split_t* out)
{
    return run_split<float>(out, SPLIT_SEPARATE, Xp, Xi, Xv, m, n, n_users_test, test_fraction, consider_cold_start != 0,
                            min_items_pool, min_pos_test, seed, device);
}

int rmb200_split_separate_users_f64(const int32_t* Xp, const int32_t* Xi, const double* Xv, int32_t m, int32_t n,
                                    int32_t n_users_test, double test_fraction, int consider_cold_start, int32_t min_items_pool,
                                    int32_t min_pos_test, uint64_t seed, int32_t device, rmb200_split_t* out)
{
    return run_split<double>(out, SPLIT_SEPARATE, Xp, Xi, Xv, m, n, n_users_test, test_fraction, consider_cold_start != 0,
                             min_items_pool, min_pos_test, seed, device);
}

int rmb200_split_joined_users_f32(const int32_t* Xp, const int32_t* Xi, const float* Xv, int32_t m, int32_t n,
                                  int32_t n_users_test, double test_fraction, int consider_cold_start, int32_t min_items_pool,
                                  int32_t min_pos_test, uint64_t seed, int32_t device, rmb200_split_t* out)
{
    return run_split<float>(out, SPLIT_JOINED, Xp, Xi, Xv, m, n, n_users_test, test_fraction, consider_cold_start != 0,
                            min_items_pool, min_pos_test, seed, device);
}

int rmb200_split_joined_users_f64(const int32_t* Xp, const int32_t* Xi, const double* Xv, int32_t m, int32_t n,
                                  int32_t n_users_test, double test_fraction, int consider_cold_start, int32_t min_items_pool,
                                  int32_t min_pos_test, uint64_t seed, int32_t device, rmb200_split_t* out)
{
    return run_split<double>(out, SPLIT_JOINED, Xp, Xi, Xv, m, n, n_users_test, test_fraction, consider_cold_start != 0,
                             min_items_pool, min_pos_test, seed, device);
}

int rmb200_split_plan(const int32_t* Xp, int32_t m, int32_t n, int32_t sample_users, int32_t n_users_test, double test_fraction,
                      int consider_cold_start, int32_t min_items_pool, int32_t min_pos_test, uint64_t seed,
                      int32_t* users_test, int32_t* n_users_out, uint8_t* held, int64_t* n_entries_out)
{
    rmb::set_last_error("", nullptr);
    if (!Xp || m <= 0 || !held || !n_entries_out) return fail(RMB200_ERR_BAD_ARG, "rmb200_split_plan: bad arguments");
    try {
        SplitPlan P;
        if (sample_users) {
            if (!users_test || !n_users_out) return fail(RMB200_ERR_BAD_ARG, "rmb200_split_plan: users_test missing");
            if (const char* refusal = pick_users(Xp, m, n, n_users_test, test_fraction, consider_cold_start != 0, min_items_pool,
                                                 min_pos_test, seed, P))
                return fail(RMB200_ERR_RUNTIME, refusal);
        }
        plan_rows(Xp, P.whole ? nullptr : P.sel_rows.data(), P.whole ? m : (int32_t)P.sel_rows.size(), test_fraction, seed, P);
        if (sample_users) {
            std::memcpy(users_test, P.sel_rows.data(), sizeof(int32_t) * P.sel_rows.size());
            *n_users_out = (int32_t)P.sel_rows.size();
        } else if (n_users_out) *n_users_out = m;
        std::memcpy(held, P.held.data(), P.held.size());
        *n_entries_out = (int64_t)P.held.size();
    } catch (const std::bad_alloc&) {
        return fail(RMB200_ERR_OOM, "host allocation failed during the split plan");
    }
    return RMB200_OK;
}

void rmb200_split_free(rmb200_split_t* split)
{
    if (!split) return;
    if (SplitOwner* own = (SplitOwner*)split->owner) {
        for (void* p : own->blocks) std::free(p);
        delete own;
    }
    const int32_t vb = split->value_bytes;
    std::memset(split, 0, sizeof(*split));
    split->value_bytes = vb;
}

int rmb200_sizeof_split(void) { return (int)sizeof(rmb200_split_t); }

}  // extern "C"
