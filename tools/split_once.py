"""One timed train/test split on the GPU box: the product (rmb200_split_*) beside the unmodified reference (oracle/_ref) on
the same matrix, results compared entry for entry.  Prints one JSON line.
    python tools/split_once.py [--users 138493] [--items 26744] [--mean-len 144] [--kind all|separated|joined] [--reps 3]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recometrics_b200 import _capi      # noqa: E402


def power_law_matrix(m, n, mean_len, seed=1, dtype=np.float32):
    rs = np.random.RandomState(seed)
    lens = np.minimum(n - 2, np.maximum(1, (rs.pareto(1.6, m) * mean_len * 0.6 + 1).astype(np.int64)))
    indptr = np.zeros(m + 1, np.int32)
    indptr[1:] = np.cumsum(lens)
    nnz = int(indptr[-1])
    row_of = np.repeat(np.arange(m), lens)
    pos = np.arange(nnz) - indptr[row_of]
    step = np.maximum(1, (n - 1) // np.maximum(lens[row_of], 1))
    indices = (pos * step + rs.randint(0, 1 << 30, size=m)[row_of] % step).astype(np.int32)
    data = ((np.arange(nnz) % 10) * 0.5 + 0.5).astype(dtype)
    return indptr, indices, data


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=138493)      # MovieLens-20M-shaped (BASELINE configs[1])
    ap.add_argument("--items", type=int, default=26744)
    ap.add_argument("--mean-len", type=float, default=144.0)
    ap.add_argument("--kind", default="all")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-reference", action="store_true")
    a = ap.parse_args()
    p, i, v = power_law_matrix(a.users, a.items, a.mean_len)
    kw = dict(test_fraction=0.3, seed=1)
    if a.kind != "all":
        kw.update(n_users_test=min(10000, a.users // 10), consider_cold_start=False, min_items_pool=2, min_pos_test=1)
    best, res = None, None
    for _ in range(a.reps + 1):                               # first call also pays CUDA context creation
        t0 = time.perf_counter()
        res = _capi.split(a.kind, p, i, v, a.users, a.items, **kw)
        wall = (time.perf_counter() - t0) * 1e3
        if best is None or (_ > 0 and wall < best[0]):
            best = (wall, res["timing"])
    out = {"workload": "split %s: %d users x %d items, %d entries, f32" % (a.kind, a.users, a.items, i.size),
           "product_wall_ms": round(best[0], 2), "product": {k: (round(x, 3) if isinstance(x, float) else x) for k, x in best[1].items()}}
    nnz_split = int(res["train"][1].size + res["test"][1].size)
    # partition kernel: reads item + value + flag + scan, writes item + value; flags / scan kernels: 1 + 1 + 1 + 4 bytes
    algo_bytes = nnz_split * (4 + 4 + 1 + 4 + 4 + 4) + nnz_split * (1 + 1 + 1 + 4) + (i.size - nnz_split if a.kind != "all" else 0) * 16
    out["kernels_GBps_algorithmic"] = round(algo_bytes / (best[1]["kernel_ms"] * 1e-3) / 1e9, 1) if best[1]["kernel_ms"] > 0 else None
    if not a.no_reference:
        import oracle
        if oracle.have_ref():
            ref_ms = None
            for _ in range(max(1, a.reps)):
                t0 = time.perf_counter()
                r = oracle.ref_split(p, i, v, a.users, a.items, split_type=a.kind, **kw)
                ms = (time.perf_counter() - t0) * 1e3
                ref_ms = ms if ref_ms is None else min(ref_ms, ms)
            same = all(np.array_equal(x, y) for key in ("train", "test", "rem") if r[key] is not None
                       for x, y in zip(r[key], res[key][:3]))
            if r["users_test"] is not None:
                same = same and np.array_equal(r["users_test"], res["users_test"])
            out.update(reference_ms=round(ref_ms, 2), identical_to_reference=bool(same), speedup=round(ref_ms / best[0], 2),
                       reference="oracle/_ref/librecometrics_ref.so, one host thread (the reference's splitters are sequential)")
    print(json.dumps(out))
