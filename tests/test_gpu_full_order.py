"""GPU: the full-order path (recometrics_b200/csrc/full_order.cu) -- every candidate scored, given its tie-breaking noise
and sorted in HBM, as /root/reference/src/recometrics.hpp:499-563 does per user.  It serves what the two selection
kernels cannot:
  * k_metrics beyond their candidate buffers (384), up to n -- the reference's only bound (hpp:391);
  * break_ties_with_noise with ROC/PR-AUC: ranks of the NOISY scores (the selection paths count ranks without noise).
Checked: bit-for-bit equality with the FMA path where both apply, parity with the oracle for large k_metrics, and the
reference's golden noise cases with nothing set aside (the noise alone decides every rank there)."""
import numpy as np
import pytest

import parity_utils as pu
from golden_io import case_names, load_case
from tools import synth

pytestmark = pytest.mark.gpu


def _run(rb, d, k, path, **kw):
    return rb.calc_reco_metrics_ex(d["X_train"], d["X_test"], d["A"], d["B"], k=k, item_biases=d["item_biases"],
                                   break_ties_with_noise=False, return_topk=True, return_status=True, scoring_path=path, **kw)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("cfg,m,n,p,k,cum", [(1, 900, 1700, 24, 10, False), (2, 700, 5000, 33, 50, True), (3, 500, 9001, 64, 20, False)])
def test_full_order_equals_fma_path_bit_for_bit(rb, dtype, cfg, m, n, p, k, cum):
    d = synth.make(cfg, m=m, n=n, p=p)
    d["A"], d["B"] = d["A"].astype(dtype), d["B"].astype(dtype)
    if d["item_biases"] is not None:
        d["item_biases"] = d["item_biases"].astype(dtype)
    kw = dict(all_metrics=True, cumulative=cum, return_ranks=True)
    a, b = _run(rb, d, k, "fma", **kw), _run(rb, d, k, "full", **kw)
    assert a.timing["scoring_path"] == 1 and b.timing["scoring_path"] == 3
    assert np.array_equal(a.status, b.status)
    assert np.array_equal(a.topk_items, b.topk_items)
    assert np.array_equal(a.topk_scores, b.topk_scores, equal_nan=True)
    assert np.array_equal(a.pos_rank, b.pos_rank)
    for key, v in a.metrics.items():
        if key != "K":
            assert np.array_equal(v, b.metrics[key], equal_nan=True), key


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("k", [385, 1000, 2500])
def test_k_metrics_beyond_the_selection_kernels_against_the_oracle(rb, oracle_mod, dtype, k):
    """k_metrics 385 .. n: the automatic choice is the full-order path; K = n puts every user under the cand <= K rules."""
    d = synth.make(1, m=600, n=2500, p=16)
    metrics = ("p", "tp", "r", "ap", "tap", "ndcg", "hit", "rr", "roc", "pr")
    res = pu.run_product(rb, d, metrics, k, dtype=dtype)
    assert res.timing["scoring_path"] == 3
    orc = pu.run_oracle(oracle_mod, d, metrics, k, dtype=dtype)
    dd = dict(d, A=d["A"].astype(dtype), B=d["B"].astype(dtype))
    # (the 1e-6 near-tie clause of the parity bar touches more users the deeper the ranked list goes: no cap on their share here)
    pu.compare(res, orc, dd, metrics, k, max_amb_frac=1.0, label="full order k=%d %s" % (k, np.dtype(dtype).name))


def test_large_k_cumulative_rows(rb, oracle_mod):
    d = synth.make(1, m=300, n=1200, p=8)
    metrics = ("p", "r", "ap", "ndcg")
    res = pu.run_product(rb, d, metrics, 600, cumulative=True)
    assert res.timing["scoring_path"] == 3 and res.metrics["P@K"].shape == (300, 600)
    orc = pu.run_oracle(oracle_mod, d, metrics, 600, cumulative=True)
    pu.compare(res, orc, d, metrics, 600, cumulative=True, max_amb_frac=1.0, label="full order k=600 cumulative")


def test_forced_selection_paths_refuse_large_k(rb):
    d = synth.make(1, m=100, n=900, p=8)
    for path in ("fma", "tensor"):
        with pytest.raises(NotImplementedError):
            _run(rb, d, 500, path, precision=True)


def _noise_ties_left(oracle_mod, c):
    """Users whose NOISY scores still hold an exact tie that matters (float32 scores of ~1e-7 sit on a grid of ~7e-15, the
    noise takes ~280 values on it): the order of those is libstdc++'s (quirk Q8).  From the oracle's restatement of the
    noise, which tests/test_oracle_golden.py pins bit-for-bit to the compiled reference.  bit0: inside the top K, bit1: anywhere."""
    kw = {k: v for k, v in c["params"].items() if k in ("seed", "min_pos_test", "min_items_pool", "consider_cold_start")}
    o = oracle_mod.oracle_calc(c["A"], c["B"], c["X_train"], c["X_test"], c["k"], metrics=c["metrics"], cumulative=c["cumulative"],
                               nthreads=4, fix_quirks=False, extras=True, break_ties_with_noise=True, dtype=c["dtype"], **kw)
    return (o["tie_flags"] & 1) != 0, (o["tie_flags"] & 2) != 0


@pytest.mark.parametrize("path", ["full", "auto"])
@pytest.mark.parametrize("name", [c for c in case_names() if c.startswith("g_noise")])
def test_golden_noise_cases_with_nothing_set_aside(rb, oracle_mod, name, path):
    """The reference's outputs with break_ties_with_noise=True on inputs full of exact ties: the per-user mt19937 stream
    decides the whole order.  Every metric -- ROC/PR-AUC included -- must be the reference's: on the full-order path, and on
    the automatic one (tensor-core filter whose exact stage adds the noise; with ROC/PR-AUC the FMA tiles count the ranks and
    hand the users for whom the noise decides a rank to the full-order path)."""
    c = load_case(name)
    assert c["params"].get("break_ties_with_noise")
    res = pu.run_product(rb, c, c["metrics"], c["k"], cumulative=c["cumulative"], extras=True, scoring_path=path, **c["params"])
    assert res.timing["scoring_path"] == (3 if path == "full" else 2)
    top_tie, any_tie = _noise_ties_left(oracle_mod, c)
    for q in c["metrics"]:
        if q in ("hit", "rr") and not any(x in c["metrics"] for x in ("p", "tp", "r", "ap", "tap", "ndcg")):
            continue   # quirk Q2
        if q == "pr" and "roc" not in c["metrics"]:
            continue   # quirk Q3
        g, o = res.metrics[pu.KEY[q]], c["ref"][q]
        ok = pu.nan_equal_close(g, o, pu.METRIC_TOL)
        if ok.ndim == 2:
            ok = ok.all(axis=1)
        aside = any_tie if q in ("roc", "pr") else top_tie
        assert (ok | aside).all(), f"{name}: {q}: {np.asarray(g)[~ok & ~aside][:4]} vs {np.asarray(o)[~ok & ~aside][:4]}"
    assert any_tie.mean() < 0.95, "test data too degenerate"
    if path == "auto" and ("roc" in c["metrics"] or "pr" in c["metrics"]) and "ties" in name:
        assert res.timing["noise_handback_users"] > 0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_noise_default_call_all_metrics_on_the_full_order_path(rb, oracle_mod, dtype):
    """All ten metrics of a default (noise on) call against the oracle's restatement of the noise, ranks included."""
    d = synth.make(1, m=500, n=1300, p=16)
    A, B = d["A"].astype(dtype), d["B"].astype(dtype)
    res = rb.calc_reco_metrics_ex(d["X_train"], d["X_test"], A, B, k=10, all_metrics=True, seed=77, return_status=True,
                                  return_ranks=True, scoring_path="full")
    o = oracle_mod.oracle_calc(A, B, d["X_train"], d["X_test"], 10, metrics=synth.ALL10, nthreads=4, dtype=dtype,
                               break_ties_with_noise=True, seed=77, extras=True)
    S64 = pu.scores_f64(A, B)
    topk_amb, rank_amb, _ = pu.ambiguity(S64, d["X_train"], d["X_test"], 10)
    assert np.array_equal(res.status[~topk_amb], o["status"][~topk_amb])
    for q in synth.ALL10:
        ok = pu.nan_equal_close(res.metrics[pu.KEY[q]], o[q], pu.METRIC_TOL)
        amb = rank_amb if q in ("roc", "pr") else topk_amb
        assert (ok | amb).all(), q
