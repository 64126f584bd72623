"""Load a tests/golden/g_*.npz evaluation case (inputs + the unmodified reference's outputs)."""
import glob
import os

import numpy as np
from scipy.sparse import csr_array

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "g_*.npz")))


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    A, B = z["A"], z["B"]
    m, n = A.shape[0], int(z["n"])
    Xtr = csr_array((np.ones(z["tr_indices"].shape[0], dtype=A.dtype), z["tr_indices"], z["tr_indptr"]), shape=(m, n))
    Xte = csr_array((z["te_data"], z["te_indices"], z["te_indptr"]), shape=(m, n))
    params = {}
    for kv in z["params"]:
        a, b = str(kv).split("=")
        params[a] = (b == "True") if b in ("True", "False") else int(b)
    ref = {q[4:]: z[q] for q in z.files if q.startswith("ref_")}
    return dict(A=A, B=B, X_train=Xtr, X_test=Xte, k=int(z["k"]), cumulative=bool(int(z["cumulative"])),
                metrics=tuple(str(q) for q in z["metrics"]), params=params, ref=ref, dtype=A.dtype.type,
                item_biases=None)
