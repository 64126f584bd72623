"""Synthetic, reference-shaped inputs for the parity tests and bench.py.

Deterministic and BLOCK-REPRODUCIBLE: users are generated in blocks of ``BLOCK`` rows, every block
from its own ``default_rng([cfg_id, block])`` stream, so the first ``m_sub`` users of a configuration
are bit-identical whether one generates ``m_sub`` or all ``m`` users (the CPU baseline is timed on
such a prefix).  Item factors come from ``default_rng([cfg_id, 10**6])``.

This is measurement/test tooling, not part of the product path.
"""
from dataclasses import dataclass, field

import numpy as np
from scipy.sparse import csr_array

BLOCK = 16384


@dataclass
class Config:
    cfg_id: int
    name: str
    m: int
    n: int
    p: int
    k: int
    dtype: type
    metrics: tuple
    mean_nnz: float
    cumulative: bool = False
    item_biases: bool = False
    cold_frac: float = 0.0
    min_pos_test: int = 1
    extra: dict = field(default_factory=dict)


ALL10 = ("p", "tp", "r", "ap", "tap", "ndcg", "hit", "rr", "roc", "pr")

# BASELINE.json "configs" 
CONFIGS = {
    1: Config(1, "cfg1 ML-1M-shaped 6040x3706 p=32 f32 K=10 all_metrics", 6040, 3706, 32, 10, np.float32, ALL10, 165),
    2: Config(2, "cfg2 ML-20M-shaped 138493x26744 p=64+bias f32 K=10 P/R/AP/NDCG", 138493, 26744, 64, 10, np.float32,
              ("p", "r", "ap", "ndcg"), 144, item_biases=True),
    3: Config(3, "cfg3 LFM-360K-shaped 359347x160112 p=128 f32 K=20 P/R/AP/NDCG+ROC+PR", 359347, 160112, 128, 20,
              np.float32, ("p", "r", "ap", "ndcg", "roc", "pr"), 48),
    4: Config(4, "cfg4 1Mx1M p=128 f32 K=100 P/R/AP/NDCG", 1000000, 1000000, 128, 100, np.float32,
              ("p", "r", "ap", "ndcg"), 100),
    5: Config(5, "cfg5 f64 500Kx300K p=64 cumulative K=50 AP/NDCG min_pos_test=2 2% cold", 500000, 300000, 64, 50,
              np.float64, ("ap", "ndcg"), 60, cumulative=True, cold_frac=0.02, min_pos_test=2),
}


def _user_block(cfg, block, rows, n):
    """CSR pieces (train/test) for users [block*BLOCK, block*BLOCK+rows)."""
    # one stream per purpose, so that a short (last) block is an exact prefix of the full block
    rng, rng_it, rng_split, rng_val = (np.random.default_rng([cfg.cfg_id, block, q]) for q in range(4))
    full = BLOCK
    c = np.clip(np.rint(rng.lognormal(np.log(cfg.mean_nnz) - 0.5, 1.0, full)), 10, max(10, min(n // 4, 5000)))
    c = np.minimum(c, max(1, n // 2)).astype(np.int64)
    cold = rng.random(full) < cfg.cold_frac
    c = c[:rows]
    cold = cold[:rows]
    u = np.repeat(np.arange(rows, dtype=np.int64), c)
    it = np.minimum(np.floor(n * rng_it.random(u.shape[0]) ** 2), n - 1).astype(np.int64)
    key = np.sort(u * n + it, kind="stable")
    key = key[np.concatenate(([True], key[1:] != key[:-1]))]   # sorted, duplicate-free (user, item) pairs
    u = key // n
    it = (key - u * n).astype(np.int32)
    is_test = rng_split.random(key.shape[0]) < 0.3
    is_test |= cold[u]
    vals = rng_val.integers(1, 6, key.shape[0])
    tr_cnt = np.bincount(u[~is_test], minlength=rows)
    te_cnt = np.bincount(u[is_test], minlength=rows)
    return tr_cnt, it[~is_test], te_cnt, it[is_test], vals[is_test]


def make(cfg_id, m=None, n=None, p=None, k=None, seed_shift=0, shared_items=False):
    """Generate (a prefix of) a BASELINE configuration.  m/n/p/k override the sizes (tests use small ones).
    seed_shift: another set of users (rank r of a multi-GPU run); shared_items: ... with the SAME item factors."""
    base = CONFIGS[cfg_id]
    cfg = Config(**{**base.__dict__})
    if m is not None:
        cfg.m = int(m)
    if n is not None:
        cfg.n = int(n)
    if p is not None:
        cfg.p = int(p)
    if k is not None:
        cfg.k = int(k)
    cfg.cfg_id = base.cfg_id + 1000 * seed_shift
    T = cfg.dtype
    mm, nn = cfg.m, cfg.n

    items_id = base.cfg_id if shared_items else cfg.cfg_id
    rngB = np.random.default_rng([items_id, 10 ** 6])
    B = rngB.standard_normal((nn, cfg.p), dtype=np.float32 if T == np.float32 else np.float64).astype(T, copy=False)
    bias = None
    if cfg.item_biases:
        bias = (0.5 * np.random.default_rng([items_id, 10 ** 6 + 100]).standard_normal(nn)).astype(T)

    A = np.empty((mm, cfg.p), dtype=T)
    tr_cnts, tr_idx, te_cnts, te_idx, te_val = [], [], [], [], []
    nblocks = (mm + BLOCK - 1) // BLOCK
    for b in range(nblocks):
        rows = min(BLOCK, mm - b * BLOCK)
        rngA = np.random.default_rng([cfg.cfg_id, 2 * 10 ** 6 + b])
        blockA = rngA.standard_normal((BLOCK, cfg.p), dtype=np.float32 if T == np.float32 else np.float64)
        A[b * BLOCK: b * BLOCK + rows] = blockA[:rows]
        a, bb, c, d, e = _user_block(cfg, b, rows, nn)
        tr_cnts.append(a); tr_idx.append(bb); te_cnts.append(c); te_idx.append(d); te_val.append(e)

    def _csr(cnts, idx, val):
        indptr = np.zeros(mm + 1, dtype=np.int64)
        np.cumsum(np.concatenate(cnts), out=indptr[1:])
        assert indptr[-1] < 2 ** 31
        indices = np.concatenate(idx).astype(np.int32, copy=False)
        data = np.ones(indices.shape[0], dtype=T) if val is None else np.concatenate(val).astype(T)
        return csr_array((data, indices, indptr.astype(np.int32)), shape=(mm, nn))

    X_train = _csr(tr_cnts, tr_idx, None)
    X_test = _csr(te_cnts, te_idx, te_val)
    cfg.cfg_id = base.cfg_id
    return dict(cfg=cfg, A=A, B=B, item_biases=bias, X_train=X_train, X_test=X_test)


def fold_biases(A, B, bias):
    """What the reference front-end does with item_biases (recometrics/__init__.py:548-551)."""
    if bias is None:
        return A, B
    return (np.c_[A, np.ones((A.shape[0], 1), dtype=A.dtype)], np.c_[B, bias.reshape(-1, 1).astype(B.dtype)])


def algorithmic_flops(cfg, X_train, X_test, k=None, min_items_pool=2, consider_cold_start=True, has_ndcg=True):
    """F = 2 * p' * sum over eligible users of (n - ntrain_u); early-NaN users count 0."""
    n = X_train.shape[1]
    K = cfg.k if k is None else k
    ntrain = np.diff(X_train.indptr).astype(np.int64)
    npos = np.diff(X_test.indptr).astype(np.int64)
    mip = max(min_items_pool, K, 2)
    elig = (npos > 0) & ~(((ntrain + npos) >= n) & (not has_ndcg)) & ((n - ntrain) >= mip)
    if not consider_cold_start:
        elig &= ntrain > 0
    pp = cfg.p + (1 if cfg.item_biases else 0)
    return float(2 * pp * np.sum((n - ntrain)[elig]))
