"""Seeded inputs for the splitter tests (tests/test_split_*.py, tests/golden/make_golden_split.py, tools/split_once.py).
numpy's legacy RandomState is used on purpose: its streams are frozen, so the committed golden outputs stay valid."""
import numpy as np


def make_csr(m, n, seed, mean_len=12.0, heavy_rows=(), empty_frac=0.05, dtype=np.float64, unsorted=False):
    """Random CSR rows with power-law-ish lengths; `heavy_rows` = explicit lengths of the first rows (long rows: the shuffle's
    second regime starts at 65,536 entries); a share of rows is empty.  Item ids are unique inside a row."""
    rs = np.random.RandomState(seed)
    lens = np.minimum(n - 1, np.maximum(1, (rs.pareto(1.5, m) * mean_len * 0.5 + 1).astype(np.int64)))
    lens[rs.rand(m) < empty_frac] = 0
    for r, ln in enumerate(heavy_rows):
        lens[r] = min(ln, n)
    indptr = np.zeros(m + 1, dtype=np.int32)
    indptr[1:] = np.cumsum(lens)
    indices = np.empty(int(indptr[-1]), dtype=np.int32)
    for r in range(m):
        ln = int(lens[r])
        if not ln:
            continue
        if ln * 3 > n:
            row = rs.permutation(n)[:ln]
        else:
            row = np.unique(rs.randint(0, n, size=ln * 2 + 8))
            while row.size < ln:
                row = np.unique(np.concatenate([row, rs.randint(0, n, size=ln)]))
            row = rs.permutation(row)[:ln]
        indices[indptr[r]:indptr[r + 1]] = row if unsorted else np.sort(row)
    data = (rs.randint(1, 50, size=indices.size) / 4.0).astype(dtype)
    return indptr, indices, data


# name -> (matrix kwargs, split kwargs)
CASES = {
    "split_all_f64": (dict(m=300, n=500, seed=11), dict(split_type="all", test_fraction=0.3, seed=1)),
    "split_all_f32_frac09": (dict(m=400, n=300, seed=12, mean_len=5.0, dtype=np.float32), dict(split_type="all", test_fraction=0.9, seed=7)),
    "split_all_f32_half": (dict(m=250, n=900, seed=13, dtype=np.float32), dict(split_type="all", test_fraction=0.5, seed=123456789012)),
    "split_all_unsorted_f64": (dict(m=200, n=400, seed=14, unsorted=True), dict(split_type="all", test_fraction=0.4, seed=3)),
    "split_all_long_rows_f32": (dict(m=6, n=70001, seed=15, heavy_rows=(66000, 65535, 65536, 40000), dtype=np.float32, empty_frac=0.0),
                                dict(split_type="all", test_fraction=0.2, seed=5)),
    "split_separated_f64": (dict(m=500, n=350, seed=16), dict(split_type="separated", n_users_test=60, test_fraction=0.3,
                                                             consider_cold_start=False, min_items_pool=2, min_pos_test=1, seed=1)),
    "split_separated_f32_strict": (dict(m=400, n=120, seed=17, mean_len=20.0, dtype=np.float32),
                                   dict(split_type="separated", n_users_test=400, test_fraction=0.5, consider_cold_start=False,
                                        min_items_pool=60, min_pos_test=3, seed=99)),
    "split_separated_cold_f64": (dict(m=300, n=200, seed=18, mean_len=2.0), dict(split_type="separated", n_users_test=100, test_fraction=0.7,
                                                                                consider_cold_start=True, min_items_pool=0, min_pos_test=0, seed=4)),
    "split_joined_f64": (dict(m=450, n=260, seed=19), dict(split_type="joined", n_users_test=45, test_fraction=0.3,
                                                          consider_cold_start=False, min_items_pool=2, min_pos_test=1, seed=2)),
    "split_joined_unsorted_f32": (dict(m=220, n=500, seed=20, unsorted=True, dtype=np.float32),
                                  dict(split_type="joined", n_users_test=0, test_fraction=0.25, consider_cold_start=True,
                                       min_items_pool=1, min_pos_test=2, seed=8)),
}

# inputs the reference answers with std::runtime_error (message checked)
REFUSALS = {
    "too_many_users": (dict(m=50, n=80, seed=21), dict(split_type="separated", n_users_test=51, test_fraction=0.3),
                       "Target number of test users is larger than available users."),
    "pool_too_large": (dict(m=50, n=80, seed=21), dict(split_type="joined", n_users_test=5, test_fraction=0.3, min_items_pool=80),
                       "Selected minimum number of items is larger than total number of items."),
    "nobody_eligible": (dict(m=50, n=80, seed=21, mean_len=2.0), dict(split_type="separated", n_users_test=5, test_fraction=0.3, min_pos_test=70),
                        "No users satisfy criteria for test inclusion."),
}


def flatten(res):
    """A split result ({"train": (p, i, v[, shape]), ...}) as a flat {name: array} dict (fixtures, comparisons)."""
    out = {}
    for key in ("train", "test", "rem"):
        if res.get(key) is not None:
            for nm, a in zip("piv", res[key][:3]):
                out[key + "_" + nm] = np.asarray(a)
    if res.get("users_test") is not None:
        out["users_test"] = np.asarray(res["users_test"])
    return out
