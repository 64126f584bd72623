"""GPU: exact score ties in the rank counts (ROC-AUC, PR-AUC, held-out ranks).

The reference leaves the order of equal scores to libstdc++'s introsort; this library ranks equal
scores by ascending item id EVERYWHERE: in the top-K selection and -- since round 2 -- in the rank counting pass, so that a
held-out item is not placed ahead of every equal-scored negative (ROC/PR-AUC were inflated on ties, and AP@K
with K = all candidates disagreed with PR_AUC).  Small-integer factors make every score exact in any summation order, so
the expected ranks follow from a plain numpy sort by (score descending, item id ascending)."""
import numpy as np
import pytest
from scipy.sparse import csr_array

pytestmark = pytest.mark.gpu


def _case(m, n, p, dtype, seed, zero_items=0.0):
    rng = np.random.default_rng(seed)
    A = rng.integers(-2, 3, size=(m, p)).astype(dtype)
    B = rng.integers(-2, 3, size=(n, p)).astype(dtype)
    if zero_items:
        B[rng.random(n) < zero_items] = 0          # cold items: every user scores them exactly 0
    tr_rows, te_rows, te_vals = [], [], []
    for u in range(m):
        items = rng.choice(n, size=int(rng.integers(6, 40)), replace=False)
        items.sort()
        is_te = rng.random(items.shape[0]) < 0.4
        if not is_te.any():
            is_te[0] = True
        tr_rows.append(items[~is_te]); te_rows.append(items[is_te])
        te_vals.append(rng.integers(1, 6, int(is_te.sum())).astype(dtype))

    def csr(rows, vals=None):
        indptr = np.zeros(m + 1, dtype=np.int32)
        indptr[1:] = np.cumsum([len(r) for r in rows])
        idx = np.concatenate(rows).astype(np.int32)
        data = np.ones(idx.shape[0], dtype=dtype) if vals is None else np.concatenate(vals)
        return csr_array((data, idx, indptr), shape=(m, n))
    return A, B, csr(tr_rows), csr(te_rows, te_vals)


def _expected(A, B, Xtr, Xte):
    """ranks of the held-out items (1-based, among the candidates), ROC and PR from them; None for degenerate users."""
    S = A.astype(np.float64) @ B.astype(np.float64).T
    m, n = S.shape
    out = []
    for u in range(m):
        tr = Xtr.indices[Xtr.indptr[u]:Xtr.indptr[u + 1]]
        te = Xte.indices[Xte.indptr[u]:Xte.indptr[u + 1]]
        cand = np.setdiff1d(np.arange(n), tr)
        s = S[u, cand]
        if s.max() == s.min():
            out.append(None)
            continue
        order = cand[np.lexsort((cand, -s))]               # score descending, item id ascending
        pos_of = np.empty(n, dtype=np.int64)
        pos_of[order] = np.arange(1, order.shape[0] + 1)
        ranks = pos_of[te]
        rs = np.sort(ranks)
        npos, nneg = te.shape[0], cand.shape[0] - te.shape[0]
        roc = 1.0 - (rs.sum() - npos * (npos + 1) / 2) / (npos * nneg)
        pr = np.mean(np.arange(1, npos + 1) / rs)
        out.append((ranks, roc, pr))
    return out


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,p,zero_items", [(300, 3, 0.0), (1000, 5, 0.3), (257, 2, 0.0)])
def test_tied_scores_rank_by_item_id_in_the_rank_counts(rb, dtype, n, p, zero_items):
    A, B, Xtr, Xte = _case(700, n, p, dtype, seed=n + p, zero_items=zero_items)
    res = rb.calc_reco_metrics_ex(Xtr, Xte, A, B, k=10, roc_auc=True, pr_auc=True, average_precision=True, precision=False,
                                  ndcg=False, break_ties_with_noise=False, return_ranks=True, return_status=True)
    exp = _expected(A, B, Xtr, Xte)
    checked = tied = 0
    for u, e in enumerate(exp):
        if e is None or res.status[u] != 0:
            continue
        ranks, roc, pr = e
        got = res.pos_rank[Xte.indptr[u]:Xte.indptr[u + 1]]
        assert np.array_equal(got, ranks), (u, got, ranks)
        assert abs(res.metrics["ROC_AUC"][u] - roc) <= 1e-6, u
        assert abs(res.metrics["PR_AUC"][u] - pr) <= 1e-6, u
        checked += 1
        tied += int(len(set(ranks.tolist())) == len(ranks))
    assert checked > 600


def test_pr_auc_equals_ap_at_all_candidates_on_ties(rb):
    """AP@K over (almost) the whole candidate list walks the same order the rank counts assume."""
    n = 200
    A, B, Xtr, Xte = _case(300, n, 3, np.float64, seed=5)
    ntr_max = int(np.diff(Xtr.indptr).max())
    K = 150
    res = rb.calc_reco_metrics_ex(Xtr, Xte, A, B, k=K, roc_auc=True, pr_auc=True, average_precision=True, precision=False,
                                  ndcg=False, break_ties_with_noise=False, return_ranks=True, return_status=True)
    n_cmp = 0
    for u in range(300):
        ranks = res.pos_rank[Xte.indptr[u]:Xte.indptr[u + 1]]
        if res.status[u] != 0 or np.isnan(res.metrics["PR_AUC"][u]) or ranks.max() > K:
            continue
        assert abs(res.metrics["PR_AUC"][u] - res.metrics["AP@K"][u]) <= 1e-9, u
        n_cmp += 1
    assert n_cmp > 30 and ntr_max < n
