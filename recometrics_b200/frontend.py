"""Host-side mirror of the reference's Python front-end for the evaluation path.

``calc_reco_metrics`` keeps the call surface of ``recometrics.calc_reco_metrics``
(/root/reference/recometrics/__init__.py:44-628): same argument names and defaults, same
validation errors, same float32/float64 rule (:469), same output packaging (:590-628) -- so a user
of the reference can switch the import and nothing else.  What changes is below the surface:

* the call goes to ``librecometrics_b200.so`` (sm_100a CUDA) through the C-ABI instead of the
  Cython/OpenMP extension;
* ``item_biases`` travel as their own vector and are added inside the scoring kernel, instead of
  being folded into copies of A and B (reference :548-551);
* ``nthreads`` is accepted and ignored (the parallelism is the GPU's);
* extras the reference cannot return (top-K ids, ranks of held-out items, per-call timing) are
  available through ``calc_reco_metrics_ex``.

There is no CPU fallback: without the built library or without a CUDA device the call raises.
"""
import ctypes
import re
from dataclasses import dataclass
from warnings import warn

import numpy as np
from scipy.sparse import csr_array, issparse

from . import _capi

__all__ = ["calc_reco_metrics", "calc_reco_metrics_ex", "EvalResult"]

# reference dict keys, in the order the reference inserts them (__init__.py:591-610)
_KEYS = (("p", "P@K"), ("tp", "TP@K"), ("r", "R@K"), ("ap", "AP@K"), ("tap", "TAP@K"),
         ("ndcg", "NDCG@K"), ("hit", "Hit@K"), ("rr", "RR@K"), ("roc", "ROC_AUC"), ("pr", "PR_AUC"))


@dataclass
class EvalResult:
    metrics: dict            # reference-style dict: "P@K" -> array, ..., "K" -> k
    timing: dict             # rmb200_timing_t of the call
    status: np.ndarray = None
    topk_items: np.ndarray = None
    topk_scores: np.ndarray = None
    pos_rank: np.ndarray = None
    means: dict = None       # return_means: reference-style keys -> mean over the evaluated users (NaN rows left out)
    counts: dict = None      # ... -> how many users entered each mean


def _rows_c_layout(X):
    """Row-major view + leading dimension in elements (reference _as_row_major, :11-16)."""
    if X.flags["C_CONTIGUOUS"]:
        return X, X.shape[1]
    if X.strides[1] != X.dtype.itemsize or X.strides[0] % X.dtype.itemsize or X.strides[0] < X.shape[1] * X.dtype.itemsize:
        return np.ascontiguousarray(X), X.shape[1]
    return X, X.strides[0] // X.dtype.itemsize


def _canonical_csr(X):
    """CSR with sorted indices and int32 index arrays (reference :553-558)."""
    if not issparse(X):
        X = csr_array(X)
    elif X.format != "csr":
        X = X.tocsr()
    if not X.has_sorted_indices:
        X = X.copy()
        X.sort_indices()
    indptr = X.indptr if X.indptr.dtype == np.int32 else X.indptr.astype(np.int32)
    indices = X.indices if X.indices.dtype == np.int32 else X.indices.astype(np.int32)
    return X, np.ascontiguousarray(indptr), np.ascontiguousarray(indices)


def _validate_and_prepare(X_train, X_test, A, B, k, item_biases, flags, break_ties_with_noise,
                          min_pos_test, min_items_pool, consider_cold_start, cumulative, nthreads, seed):
    if (A is None) != (B is None):
        raise ValueError("'A' and 'B' must either be passed together or passed as 'None' together.")
    if item_biases is not None and hasattr(item_biases, "to_numpy"):
        item_biases = item_biases.to_numpy()
    if A is None:
        # non-personalised model: score = item bias (reference :429-436)
        if item_biases is None:
            raise ValueError("Must pass item biases if not passing factors.")
        A = np.ones((X_test.shape[0], 1), dtype=np.float64)
        B = np.ascontiguousarray(item_biases, dtype=np.float64).reshape((-1, 1))
        item_biases = None

    if not isinstance(A, np.ndarray) or not isinstance(B, np.ndarray):
        raise TypeError("'A' and 'B' must be numpy arrays.")
    if not issparse(X_test):
        raise TypeError("'X_test' must be a sparse matrix.")
    imax = np.iinfo(np.int32).max
    if X_test.shape[0] >= imax:
        raise ValueError("Number of test user is larger than maximum supported.")
    if X_test.shape[1] >= imax:
        raise ValueError("Number of items is larger than maximum supported.")
    if not X_test.data.shape[0]:
        raise ValueError("'X_test' is empty.")
    if A.ndim != 2:
        raise ValueError("'A' must be a 2-dimensional array.")
    if B.ndim != 2:
        raise ValueError("'B' must be a 2-dimensional array.")
    if A.shape[1] != B.shape[1]:
        raise ValueError("'A' and 'B' must have the same number of columns.")
    if 0 in (A.shape[0], A.shape[1], B.shape[1], X_test.shape[0], X_test.shape[1]):
        raise ValueError("Input matrices cannot be empty.")
    if A.shape[0] < X_test.shape[0]:
        raise ValueError("Number of users in 'A' and 'X_test' does not match.")
    if B.shape[0] < X_test.shape[1]:
        raise ValueError("Number of items in 'B' and 'X_test' does not match.")
    if A.shape[0] > X_test.shape[0]:
        warn("'A' has more users than 'X_test'.")
        A = A[:X_test.shape[0], :]
    if B.shape[0] > X_test.shape[1]:
        warn("'B' has more items than 'X_test'.")
        B = B[:X_test.shape[1], :]

    # float32 only if both factor matrices are float32 (reference :469)
    dtype = np.float32 if (A.dtype == np.float32 and B.dtype == np.float32) else np.float64

    if X_train is None:
        X_train = csr_array(X_test.shape, dtype=dtype)
        consider_cold_start = True
    if not issparse(X_train):
        raise TypeError("'X_train' must be a sparse matrix.")
    if X_train.shape[1] != X_test.shape[1]:
        raise ValueError("'X_train' and 'X_test' should have the same number of columns.")
    if X_train.shape[0] < X_test.shape[0]:
        raise ValueError("'X_train' and 'X_test' should have the same number of rows.")
    if X_train.shape[0] > X_test.shape[0]:
        warn("'X_train' mas more rows than 'X_test'.")

    # the reference's guard ignores recall / trunc_* / pr_auc (quirk Q7, :499-500) -- kept as is
    if not (flags["p"] or flags["ap"] or flags["ndcg"] or flags["hit"] or flags["rr"] or flags["roc"]):
        raise ValueError("Must pass at least one metric to calculate.")

    if isinstance(seed, np.random.RandomState):
        seed = int(seed.randint(imax))
    elif isinstance(seed, np.random.Generator):
        seed = int(seed.integers(imax))
    nthreads = 1 if nthreads is None else int(nthreads)
    seed, k, min_pos_test, min_items_pool = int(seed), int(k), int(min_pos_test), int(min_items_pool)
    if seed < 1 or k < 1 or min_pos_test < 1 or min_items_pool < 1:
        raise ValueError("'seed', 'k', 'min_pos_test' and 'min_items_pool' must be >= 1.")
    if k > X_test.shape[1]:
        raise ValueError("'k' should be smaller than the number of items.")

    if item_biases is not None:
        if not isinstance(item_biases, np.ndarray):
            raise TypeError("'item_biases' must be a numpy array.")
        if item_biases.ndim > 2:
            raise ValueError("'item_biases' should be a 1-d array.")
        item_biases = item_biases.reshape(-1)
        if not item_biases.shape[0]:
            raise ValueError("'item_biases' is empty.")
        if item_biases.shape[0] < X_test.shape[1]:
            raise ValueError("Number of items in 'item_biases' must match with 'X_test'.")
        if item_biases.shape[0] > X_test.shape[1]:
            warn("'item_biases' has more items than 'X_test'.")
            item_biases = item_biases[:X_test.shape[1]]
        item_biases = np.ascontiguousarray(item_biases, dtype=dtype)

    X_train, trp, tri = _canonical_csr(X_train)
    X_test, tep, tei = _canonical_csr(X_test)
    tev = np.ascontiguousarray(X_test.data, dtype=dtype)
    A = A if A.dtype == dtype else A.astype(dtype)
    B = B if B.dtype == dtype else B.astype(dtype)
    A, lda = _rows_c_layout(A)
    B, ldb = _rows_c_layout(B)
    return dict(A=A, lda=lda, B=B, ldb=ldb, dtype=dtype, m=X_test.shape[0], n=X_test.shape[1], p=A.shape[1],
                trp=trp, tri=tri, tep=tep, tei=tei, tev=tev, k=k, item_biases=item_biases,
                min_pos_test=min_pos_test, min_items_pool=min_items_pool,
                consider_cold_start=bool(consider_cold_start), nthreads=nthreads, seed=seed)


def calc_reco_metrics_ex(
        X_train, X_test, A, B, k=5, item_biases=None,
        precision=True, trunc_precision=False, recall=False, average_precision=True,
        trunc_average_precision=False, ndcg=True, hit=False, rr=False, roc_auc=False, pr_auc=False,
        all_metrics=False, break_ties_with_noise=True, min_pos_test=1, min_items_pool=2,
        consider_cold_start=True, cumulative=False, nthreads=-1, seed=1,
        device=-1, user_range=None, strict_min_pos_test=False,
        return_topk=False, return_ranks=False, return_status=False, scoring_path="auto",
        return_means=False, means_only=False, filter_stats=False, nan_bits=None, devices=None):
    """Same evaluation as :func:`calc_reco_metrics`, returning an :class:`EvalResult` with the
    reference-style dict plus timing and the optional extras (top-K ids/scores, held-out ranks,
    per-user status).  ``user_range=(begin, end)`` evaluates only those rows (the sharding unit);
    rows outside it are left as NaN in the returned arrays.  ``scoring_path``: "auto" | "fma" (every score on
    the FP32/FP64 FMA pipe) | "tensor" (fp16 tensor-core candidate filter + exact FMA re-scoring of the
    survivors: identical top-K and scores, top-K metrics only).  ``return_means``: also reduce every requested
    metric to its mean over the evaluated users on the device (``numpy.nanmean`` of the per-user output; ``(k,)`` vectors
    when ``cumulative``) -- ``result.means`` / ``result.counts``; with ``means_only`` the per-user rows are not copied
    back at all (``result.metrics`` then only holds ``"K"``).  ``devices``: list of CUDA ordinals to spread the users of
    this one call over (one host thread per GPU inside the native call).  ``nan_bits``: bit pattern written wherever a
    metric is undefined (R's ``NA_REAL``) instead of a plain NaN.  ``filter_stats``: fill ``timing["filter_err_ratio_max"]``."""
    flags = dict(p=precision, tp=trunc_precision, r=recall, ap=average_precision, tap=trunc_average_precision,
                 ndcg=ndcg, hit=hit, rr=rr, roc=roc_auc, pr=pr_auc)
    flags = {q: bool(v) or bool(all_metrics) for q, v in flags.items()}
    cumulative = bool(cumulative)
    prep = _validate_and_prepare(X_train, X_test, A, B, k, item_biases, flags, break_ties_with_noise,
                                 min_pos_test, min_items_pool, consider_cold_start, cumulative, nthreads, seed)
    dtype, m, K = prep["dtype"], prep["m"], prep["k"]

    outs = {}
    for q in _capi.METRIC_ORDER:
        if flags[q]:
            size = m * K if (cumulative and q in _capi.TOPK_METRICS) else m
            outs[q] = np.empty(size, dtype=dtype)
            if user_range is not None:
                outs[q].fill(np.nan)
    timing = _capi.Timing()
    status = np.zeros(m, dtype=np.int32) if return_status else None
    topk_items = np.full(m * K, -1, dtype=np.int32) if return_topk else None
    topk_scores = np.full(m * K, np.nan, dtype=dtype) if return_topk else None
    pos_rank = np.zeros(max(int(prep["tep"][-1]), 1), dtype=np.int64) if return_ranks else None
    ub, ue = (0, 0) if user_range is None else (int(user_range[0]), int(user_range[1]))
    if user_range is not None and ub == ue:
        ub, ue = 0, -1          # an EMPTY block (a rank with no users); 0,0 would mean "all users" in the C-ABI
    return_means = bool(return_means) or bool(means_only)
    W = K if cumulative else 1
    means = np.full(10 * W, np.nan, dtype=np.float64) if return_means else None
    counts = np.zeros(10 * W, dtype=np.int64) if return_means else None
    extra = _capi.make_extra(device=device, user_begin=ub, user_end=ue, strict_min_pos_test=strict_min_pos_test,
                             topk_items=topk_items, topk_scores=topk_scores, pos_rank=pos_rank, status=status,
                             timing=timing, scoring_path=scoring_path, metric_means=means, metric_counts=counts,
                             skip_row_copy=bool(means_only), filter_stats=filter_stats, nan_bits=nan_bits, devices=devices)
    rc = _capi.calc_metrics(dtype, prep["A"], prep["lda"], prep["B"], prep["ldb"], m, prep["n"], prep["p"],
                            prep["trp"], prep["tri"], prep["tep"], prep["tei"], prep["tev"], K, cumulative,
                            bool(break_ties_with_noise), outs, prep["consider_cold_start"], prep["min_items_pool"],
                            prep["min_pos_test"], prep["nthreads"], prep["seed"], prep["item_biases"], extra)
    _capi.raise_for_status(rc)

    metrics, mean_d, count_d = {}, None, None
    for q, key in _KEYS:
        if q in outs and not means_only:
            metrics[key] = outs[q].reshape(m, K) if (cumulative and q in _capi.TOPK_METRICS) else outs[q]
    metrics["K"] = K
    if return_means:
        mean_d, count_d = {}, {}
        for i, q in enumerate(_capi.METRIC_ORDER):
            if q not in outs:
                continue
            key = dict(_KEYS)[q]
            wide = cumulative and q in _capi.TOPK_METRICS
            mean_d[key] = means[i * W:(i + 1) * W].copy() if wide else float(means[i * W])
            count_d[key] = counts[i * W:(i + 1) * W].copy() if wide else int(counts[i * W])
    return EvalResult(metrics=metrics, timing=timing.as_dict(), status=status, means=mean_d, counts=count_d,
                      topk_items=None if topk_items is None else topk_items.reshape(m, K),
                      topk_scores=None if topk_scores is None else topk_scores.reshape(m, K),
                      pos_rank=None if pos_rank is None else pos_rank[: int(prep["tep"][-1])])


def calc_reco_metrics(
        X_train, X_test, A, B, k=5, item_biases=None, as_df=True,
        precision=True, trunc_precision=False, recall=False, average_precision=True,
        trunc_average_precision=False, ndcg=True, hit=False, rr=False, roc_auc=False, pr_auc=False,
        all_metrics=False, rename_k=True, break_ties_with_noise=True, min_pos_test=1, min_items_pool=2,
        consider_cold_start=True, cumulative=False, nthreads=-1, seed=1):
    """Drop-in for ``recometrics.calc_reco_metrics`` (reference __init__.py:44-628), computed on a B200.

    Returns what the reference returns: a ``pandas.DataFrame`` (``as_df=True``; columns ``P@5`` ... or
    ``P@1..P@k`` blocks when ``cumulative``) or a dict keyed ``"P@K"``, ..., ``"ROC_AUC"``, ``"PR_AUC"``,
    ``"K"`` (``as_df=False``).  Rows of users that cannot be evaluated are NaN (reference rules,
    src/recometrics.hpp:193-209)."""
    res = calc_reco_metrics_ex(
        X_train, X_test, A, B, k=k, item_biases=item_biases,
        precision=precision, trunc_precision=trunc_precision, recall=recall, average_precision=average_precision,
        trunc_average_precision=trunc_average_precision, ndcg=ndcg, hit=hit, rr=rr, roc_auc=roc_auc, pr_auc=pr_auc,
        all_metrics=all_metrics, break_ties_with_noise=break_ties_with_noise, min_pos_test=min_pos_test,
        min_items_pool=min_items_pool, consider_cold_start=consider_cold_start, cumulative=cumulative,
        nthreads=nthreads, seed=seed)
    out = res.metrics
    if not as_df:
        return out
    import pandas as pd
    K = out.pop("K")
    if not cumulative:
        df = pd.DataFrame(out)
        if rename_k:
            df.columns = [re.sub(r"@K$", "@" + str(K), c) for c in df.columns]
        return df
    blocks = []
    for name, v in out.items():
        if name.endswith("@K"):
            cols = [name[:-1] + str(i + 1) for i in range(v.shape[1])]
            blocks.append(pd.DataFrame(v, columns=cols))
        else:
            blocks.append(pd.DataFrame({name: v}))
    return pd.concat(blocks, axis=1)
