"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile
from /root/reference/src).  Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Each file holds the inputs of one case and the reference's outputs for it (break_ties_with_noise
off unless the case name says otherwise), so the fixtures travel to machines without the reference.
"""
import os
import sys

import numpy as np
from scipy.sparse import csr_array

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from tools import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ALL10 = synth.ALL10


def _save(name, A, B, Xtr, Xte, k, metrics, cumulative, dtype, ref, **params):
    out = dict(A=A.astype(dtype), B=B.astype(dtype),
               tr_indptr=Xtr.indptr.astype(np.int32), tr_indices=Xtr.indices.astype(np.int32),
               te_indptr=Xte.indptr.astype(np.int32), te_indices=Xte.indices.astype(np.int32),
               te_data=Xte.data.astype(dtype), n=np.int64(Xte.shape[1]),
               k=np.int64(k), cumulative=np.int64(cumulative), metrics=np.array(metrics),
               params=np.array([f"{a}={b}" for a, b in sorted(params.items())]))
    for q, v in ref.items():
        out["ref_" + q] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {q: float(np.nanmean(v)) if np.isfinite(np.nanmean(v)) else None for q, v in ref.items()})


def synth_case(name, cfg_id, m, n, p, k, metrics, cumulative, dtype, **kw):
    d = synth.make(cfg_id, m=m, n=n, p=p, k=k)
    A, B = synth.fold_biases(d["A"].astype(dtype), d["B"].astype(dtype),
                             None if d["item_biases"] is None else d["item_biases"].astype(dtype))
    ref = oracle.ref_calc(A, B, d["X_train"], d["X_test"], k, metrics=metrics, cumulative=cumulative,
                          nthreads=1, dtype=dtype, **kw)
    _save(name, A, B, d["X_train"], d["X_test"], k, metrics, cumulative, dtype, ref, **kw)


def edge_case(name, dtype, metrics, cumulative, k=5, **kw):
    """Hand-built users hitting every branch of the eligibility / NaN rules."""
    rng = np.random.default_rng(7)
    n, p = 12, 4
    rows_tr, rows_te, vals = [], [], []

    def user(tr, te, v=None):
        rows_tr.append(sorted(tr)); rows_te.append(sorted(te))
        vals.append(list(v) if v is not None else [float(1 + (i % 5)) for i in range(len(te))])

    user([0, 1], [2, 3, 7])                       # ordinary
    user([], [4, 5])                              # cold start
    user([1, 2, 3], [])                           # no held-out items            -> NaN (hpp:440)
    user(list(range(0, 8)), [8, 9, 10, 11])       # train+test == n (only_ndcg)  (hpp:479-482)
    user(list(range(0, 7)), [8, 9])               # cand == k (k_leq_n)          (hpp:483)
    user(list(range(0, 9)), [9, 10])              # cand < k  -> min_items_pool  (hpp:445)
    user([0], [1, 2, 3, 4, 5, 6, 7, 8], [5, 4, 3, 2, 1, 1, 2, 3])   # npos > k
    user([3], [0, 6, 9], [1.0, -2.0, 3.0])        # a negative held-out value    (hpp:906-913)
    user([3], [0, 6, 9], [-1.0, -2.0, -3.0])      # all negative                 -> NDCG NaN (hpp:880)
    user([3], [0, 6, 9], [0.0, 0.0, 0.0])         # all zero                     -> NDCG NaN
    user([5], [0])                                # single held-out item
    m = len(rows_tr)
    A = rng.standard_normal((m, p))
    B = rng.standard_normal((n, p))

    def csr(rows, data=None):
        indptr = np.cumsum([0] + [len(r) for r in rows]).astype(np.int32)
        idx = np.array([j for r in rows for j in r], dtype=np.int32)
        dat = np.ones(len(idx)) if data is None else np.array([x for r in data for x in r], dtype=np.float64)
        return csr_array((dat, idx, indptr), shape=(m, n))

    Xtr, Xte = csr(rows_tr), csr(rows_te, vals)
    ref = oracle.ref_calc(A.astype(dtype), B.astype(dtype), Xtr, Xte, k, metrics=metrics, cumulative=cumulative,
                          nthreads=1, dtype=dtype, **kw)
    _save(name, A, B, Xtr, Xte, k, metrics, cumulative, dtype, ref, **kw)


def kat_case(name, scores, te_items, te_vals, k, dtype=np.float64):
    """Single-user known-answer cases restated from the reference's tests/testthat/test-ndcg.R."""
    n = len(scores)
    A = np.ones((1, 1))
    B = np.array(scores, dtype=np.float64).reshape(n, 1)
    Xtr = csr_array((1, n), dtype=np.float64)
    Xte = csr_array((np.array(te_vals, dtype=np.float64), np.array(te_items, dtype=np.int32),
                     np.array([0, len(te_items)], dtype=np.int32)), shape=(1, n))
    ref = oracle.ref_calc(A.astype(dtype), B.astype(dtype), Xtr, Xte, k, metrics=("ndcg",), nthreads=1, dtype=dtype)
    _save(name, A, B, Xtr, Xte, k, ("ndcg",), False, dtype, ref)


def tie_case(name, dtype, scale, metrics, cumulative=False, k=6, all_equal_users=2, **kw):
    """Tie-heavy users: every item factor row is one of 5 distinct rows, so each user has only 5 distinct scores and the
    order inside a group of tied items is decided by the tie-breaking noise alone (src/recometrics.hpp:531-534) -- the
    case that pins the per-user mt19937 stream and the uniform_real_distribution arithmetic.  `scale` shrinks the scores
    (float32 noise of 1e-12 only survives the addition for |score| < 2^-15).  The last users have all-equal scores
    (NaN row by the noise branch's validity rule, :527)."""
    rng = np.random.default_rng(11)
    m, n, p = 40, 60, 3
    proto = rng.standard_normal((5, p))
    B = proto[rng.integers(0, 5, n)] * scale
    A = rng.standard_normal((m, p))
    A[m - all_equal_users:] = 0.0
    rows_tr, rows_te = [], []
    for u in range(m):
        items = rng.permutation(n)
        ntr, nte = int(rng.integers(0, 8)), int(rng.integers(1, 9))
        rows_tr.append(sorted(items[:ntr].tolist()))
        rows_te.append(sorted(items[ntr:ntr + nte].tolist()))

    def csr(rows, vals=None):
        indptr = np.cumsum([0] + [len(r) for r in rows]).astype(np.int32)
        idx = np.array([j for r in rows for j in r], dtype=np.int32)
        dat = np.ones(len(idx)) if vals is None else vals
        return csr_array((dat, idx, indptr), shape=(m, n))

    Xtr = csr(rows_tr)
    Xte = csr(rows_te, rng.integers(1, 6, sum(len(r) for r in rows_te)).astype(np.float64))
    ref = oracle.ref_calc(A.astype(dtype), B.astype(dtype), Xtr, Xte, k, metrics=metrics, cumulative=cumulative,
                          nthreads=1, dtype=dtype, **kw)
    _save(name, A, B, Xtr, Xte, k, metrics, cumulative, dtype, ref, **kw)


if __name__ == "__main__":
    oracle.build()
    assert oracle.have_ref(), "oracle/_ref is missing: run `make -C oracle` where /root/reference exists"
    synth_case("g_f32_all", 1, 96, 300, 12, 7, ALL10, False, np.float32)
    synth_case("g_f32_all_cum", 1, 96, 300, 12, 7, ALL10, True, np.float32)
    synth_case("g_f64_all", 3, 80, 260, 9, 20, ALL10, False, np.float64)
    synth_case("g_f64_cum_apndcg_cold", 5, 120, 400, 16, 50, ("ap", "ndcg"), True, np.float64, min_pos_test=2)
    synth_case("g_f32_bias_prapndcg", 2, 64, 220, 8, 10, ("p", "r", "ap", "ndcg"), False, np.float32)
    synth_case("g_f32_nocold", 1, 64, 220, 8, 5, ("p", "tp", "hit", "rr", "roc"), False, np.float32,
               consider_cold_start=False, min_items_pool=30)
    edge_case("g_edge_f64_all", np.float64, ALL10, False)
    edge_case("g_edge_f64_all_cum", np.float64, ALL10, True)
    edge_case("g_edge_f32_noauc", np.float32, ("p", "tp", "r", "ap", "tap", "ndcg", "hit", "rr"), False)
    edge_case("g_edge_f64_pr_only_top", np.float64, ("p", "r", "hit"), False)
    edge_case("g_edge_f64_k7", np.float64, ALL10, True, k=7)
    # test-ndcg.R:107-124 "Fewer items than k": NDCG@3 == NDCG@5 == 0.5525005
    kat_case("g_kat_fewer_k3", [0, 0, 3, 4, 0, 0, 0, 0, 2, 1], [2, 3, 7], [1, 2, 3], 3)
    kat_case("g_kat_fewer_k5", [0, 0, 3, 4, 0, 0, 0, 0, 2, 1], [2, 3, 7], [1, 2, 3], 5)
    # test-ndcg.R:37-71 "Some negative values"
    kat_case("g_kat_neg_a", [0, 0, 3, 4, 0, 6, 0, 8, 9, 10], [2, 3, 5, 7, 9], [1, 2, -3, 4, 5], 5)
    kat_case("g_kat_neg_b", [0, 0, 3, 4, 0, 600, 0, 8, 9, 10], [2, 3, 5, 7, 9], [1, 2, -300, 4, 5], 5)
    kat_case("g_kat_neg_c", [0, 0, 3, 4, 0, -6, 0, 8, 9, 10], [2, 3, 5, 7, 9], [1, 2, -300, 4, 5], 5)
    kat_case("g_kat_allzero_scores", [0] * 10, [2, 3, 5, 7, 9], [1, 2, -3, 4, 5], 5)
    # break_ties_with_noise=True (the reference's default): ordinary inputs, and ties that only the noise orders
    synth_case("g_noise_f64_all", 1, 96, 300, 12, 7, ALL10, False, np.float64, break_ties_with_noise=True, seed=7)
    synth_case("g_noise_f32_all_cum", 1, 96, 300, 12, 7, ALL10, True, np.float32, break_ties_with_noise=True, seed=1)
    tie_case("g_noise_ties_f64", np.float64, 1.0, ALL10, break_ties_with_noise=True, seed=3)
    tie_case("g_noise_ties_f64_cum", np.float64, 1.0, ("p", "ap", "ndcg", "rr"), cumulative=True, break_ties_with_noise=True, seed=12345678901)
    tie_case("g_noise_ties_f32_tiny", np.float32, 1e-7, ALL10, break_ties_with_noise=True, seed=3)
    edge_case("g_noise_edge_f64_all", np.float64, ALL10, False, break_ties_with_noise=True, seed=5)
