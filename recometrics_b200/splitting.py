"""``split_reco_train_test`` -- drop-in for ``recometrics.split_reco_train_test``
(/root/reference/recometrics/__init__.py:630-851): same arguments, same checks in the same order with the same exceptions,
same return tuples, and -- built against the same libstdc++ -- the same split, entry for entry.

The call goes to ``rmb200_split_{selected,separate,joined}_users_f32/_f64`` (include/recometrics_b200.h): the reference's
sequential ``mt19937`` stream is replayed on the host, the matrix work (row ordering, partition, gathers) runs on the GPU.
There is no CPU fallback.
"""
from warnings import warn

import numpy as np
from scipy.sparse import csr_array, issparse

from . import _capi


def _as_csr(X):
    """/root/reference/recometrics/__init__.py:34-41 (indices are sorted IN PLACE, as there)."""
    if issparse(X):
        if X.format != "csr":
            X = X.tocsr()
        X.sort_indices()
        return X
    return csr_array(X)


def _int32_indices(X):
    """/root/reference/recometrics/__init__.py:27-32."""
    if (X.indptr.dtype != np.int32) or (X.indices.dtype != np.int32):
        X = X.copy()
        X.indptr = X.indptr.astype(np.int32)
        X.indices = X.indices.astype(np.int32)
    return X


def _to_scipy(parts):
    indptr, indices, data, shape = parts
    return csr_array((data, indices, indptr), shape=shape)


def split_reco_train_test(X, split_type="separated", users_test_fraction=0.1, max_test_users=10000, items_test_fraction=0.3,
                          min_items_pool=2, min_pos_test=1, consider_cold_start=False, seed=1, device=-1, return_timing=False):
    """See ``recometrics.split_reco_train_test``.  Returns, as the reference does (recometrics/wrapper.pyx:597-820):

    * ``split_type="all"``: ``(X_train, X_test)``
    * ``split_type="separated"``: ``(X_rem, X_train, X_test, users_test)``  (the reference's actual order, wrapper.pyx:669)
    * ``split_type="joined"``: ``(X_train, X_test, users_test)``

    Extensions: ``device`` (CUDA ordinal, -1 = current) and ``return_timing`` (appends the call's timing dict to the tuple).
    """
    # The reference's argument handling (recometrics/__init__.py:764-808), check for check and in its order.  It uses bare
    # `assert`s: the same AssertionError is raised here explicitly, so that `python -O` does not switch the checks off.
    def require(condition):
        if not condition:
            raise AssertionError()

    n_rows, n_cols = X.shape[0], X.shape[1]
    if not max_test_users:                       # None or 0: "as many as there are"
        max_test_users = n_rows
    require(max_test_users > 0)
    for lower_bounded in (seed, min_pos_test, min_items_pool):
        require(lower_bounded >= 0)
    max_test_users, seed = int(max_test_users), int(seed)
    min_pos_test, min_items_pool = int(min_pos_test), int(min_items_pool)
    if users_test_fraction is not None:
        require(0 < users_test_fraction < 1)
        users_test_fraction = float(users_test_fraction)
    require(0 < items_test_fraction < 1)
    items_test_fraction = float(items_test_fraction)
    require(split_type in ("all", "separated", "joined"))
    consider_cold_start = bool(consider_cold_start)

    for name, value in (("min_pos_test", min_pos_test), ("min_items_pool", min_items_pool)):
        if value >= n_cols:
            raise ValueError("'%s' must be smaller than the number of columns in 'X'." % name)

    n_users_take = 0
    if split_type != "all":
        if n_rows < 2:
            raise ValueError("'X' has less than 2 rows.")
        if users_test_fraction is None:
            if max_test_users > n_rows:
                warn("'max_test_users' is larger than number of users. Will take all.")
            n_users_take = min(max_test_users, n_rows)
        else:
            wanted = n_rows * users_test_fraction
            if wanted < 1:
                warn("Desired fraction of test users implies <1, will select 1 user.")
                wanted = 1
            n_users_take = min(round(wanted), max_test_users)

    X = _as_csr(X)
    if (not X.shape[0]) or (not X.shape[1]):
        raise ValueError("'X' cannot be empty.")
    if X.dtype not in (np.float32, np.float64):
        X = X.astype(np.float64)
    if not X.data.shape[0]:
        raise ValueError("'X' contains no non-zero entries.")
    X = _int32_indices(X)

    res = _capi.split(split_type, np.ascontiguousarray(X.indptr), np.ascontiguousarray(X.indices), np.ascontiguousarray(X.data),
                      int(X.shape[0]), int(X.shape[1]), n_users_test=int(n_users_take), test_fraction=items_test_fraction,
                      consider_cold_start=consider_cold_start, min_items_pool=min_items_pool, min_pos_test=min_pos_test,
                      seed=seed, device=device)
    if split_type == "all":
        out = (_to_scipy(res["train"]), _to_scipy(res["test"]))
    elif split_type == "separated":
        out = (_to_scipy(res["rem"]), _to_scipy(res["train"]), _to_scipy(res["test"]), res["users_test"])
    else:
        out = (_to_scipy(res["train"]), _to_scipy(res["test"]), res["users_test"])
    return out + (res["timing"],) if return_timing else out
