"""GPU: randomised differential test of the three scoring paths (tools/fuzz_paths.py): random shapes, metric sets, flags, exact
ties, zero-factor users, noise on and off -- FMA tiles, tensor-core filter (+ hand-backs) and full-order path must agree bit for
bit.  A 25-second budget here (a few hundred cases); profiles/r02fin_fuzz_paths.txt holds a 200-second run (6161 cases)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [11, 12])
def test_random_cases_agree_across_paths(rb, seed):
    run = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_paths.py"), "25", str(seed)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "all paths identical" in run.stdout, run.stdout[-2000:] + run.stderr[-2000:]


def test_random_small_cases_against_the_oracle(rb, oracle_mod):
    """Random shapes, metric sets and flags against the oracle (the C restatement pinned to the compiled reference), with the
    north-star's comparison rule (tests/parity_utils.py)."""
    import numpy as np
    import parity_utils as pu
    from tools import synth
    rng = np.random.default_rng(2024)
    names = ("p", "tp", "r", "ap", "tap", "ndcg", "hit", "rr", "roc", "pr")
    for case in range(40):
        m, n, p = int(rng.integers(1, 400)), int(rng.integers(30, 3000)), int(rng.integers(1, 70))
        K = int(min(n - 1, rng.choice([1, 2, 5, 10, 50, 300, 500])))
        dtype = np.float32 if rng.random() < 0.5 else np.float64
        d = synth.make(int(rng.integers(1, 5)), m=m, n=n, p=p)
        metrics = tuple(q for q in names if rng.random() < 0.5)
        if not any(q in metrics for q in ("p", "ap", "ndcg")):
            metrics = ("p",) + metrics
        if "pr" in metrics and "roc" not in metrics:
            metrics = metrics + ("roc",)                     # (quirk Q3: PR-AUC alone walks a partially sorted list in the reference)
        cumulative = bool(rng.random() < 0.3)
        kw = dict(min_pos_test=int(rng.integers(1, 3)), consider_cold_start=bool(rng.random() < 0.8))
        res = pu.run_product(rb, d, metrics, K, cumulative=cumulative, dtype=dtype, **kw)
        orc = pu.run_oracle(oracle_mod, d, metrics, K, cumulative=cumulative, dtype=dtype, **kw)
        dd = dict(d, A=d["A"].astype(dtype), B=d["B"].astype(dtype),
                  item_biases=None if d["item_biases"] is None else d["item_biases"].astype(dtype))
        pu.compare(res, orc, dd, metrics, K, cumulative=cumulative, max_amb_frac=1.0,
                   label="random case %d: m=%d n=%d p=%d K=%d %s %s" % (case, m, n, p, K, np.dtype(dtype).name, metrics))
