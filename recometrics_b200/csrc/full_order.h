// full_order.h -- interface of the full-order path (full_order.cu): every candidate of a user scored, (optionally) given
// the reference's tie-breaking noise, and sorted -- what /root/reference/src/recometrics.hpp:499-563 does literally
// (dot1 over the candidate list, noise :514-535, std::sort :552-554).  The selection kernels of the other two paths keep
// a bounded number of candidates per user (k_metrics <= 384) and count ranks without noise; this path has no such bound:
// it serves k_metrics up to n (the reference's only limit, hpp:391) and the users for whom the noise decides a rank.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace rmb {

template <typename T>
struct FullOrderArgs {
    const T* At;            // [user tiles][p_pad][128] user factors of the batch (pack_tiles_kernel)
    const T* Bt;            // [item tiles][p_pad][BN] item factors
    const T* bias;          // [n_pad] or nullptr
    int p_pad, n, K, C;     // C: row pitch of cand_score / cand_item (>= min(K, n))
    int user0;              // CSR / status row of the batch's first user
    int nb;                 // users of the batch
    const int* umap;        // optional [nb]: row r of At is batch-local user umap[r] (hand-backs); else the identity
    const int *trp, *tri, *tep, *tei;
    const int* ustatus;
    int* uflags;
    T* cand_score; int* cand_item; int* cand_count;      // ranked top-K of every user (batch-local rows)
    // rank outputs (all nullptr when no ROC/PR-AUC / held-out ranks are wanted)
    unsigned long long* umin;   // [m] orderable(smallest candidate score)
    unsigned* auc_cnt;          // [nnz_test] slot j of a row = its (npos - j)-th best held-out item: candidates ranked before it
    int* pos_perm;              // [nnz_test] ... and its entry offset inside the row
    int noise;                  // break_ties_with_noise
    unsigned long long seed_user0;   // seed + global index of the batch's first user
    void* scratch; size_t scratch_bytes;
    int chunk_users;            // users sorted at a time (full_order_plan)
};

// users per chunk and scratch bytes for a catalogue of n items (elem_bytes = sizeof(T)); max_users = users of a batch
void full_order_plan(int n, int elem_bytes, int max_users, int* chunk_users, size_t* scratch_bytes);

template <typename T>
cudaError_t full_order_run(const FullOrderArgs<T>& a, cudaStream_t st, long long* launches);

}  // namespace rmb
