"""Run the hot path a few times on one synthetic workload (used under ncu / for quick timings)."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth
import recometrics_b200 as rb

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--users", type=int, default=18944)
ap.add_argument("--items", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--factors", type=int, default=0)
ap.add_argument("--k", type=int, default=0)
ap.add_argument("--f64", action="store_true")
ap.add_argument("--noise", action="store_true", help="break_ties_with_noise=True (the reference default)")
ap.add_argument("--no-auc", action="store_true", help="drop ROC/PR-AUC from the configuration's metrics")
ap.add_argument("--devices", default="", help="comma-separated CUDA ordinals: spread this ONE call over them")
a = ap.parse_args()
cfg = synth.CONFIGS[a.config]
d = synth.make(a.config, m=a.users, n=a.items or cfg.n, p=a.factors or None)
flags = {q: (q in cfg.metrics) and not (a.no_auc and q in ("roc", "pr")) for q in synth.ALL10}
kw = dict(precision=flags["p"], trunc_precision=flags["tp"], recall=flags["r"], average_precision=flags["ap"],
          trunc_average_precision=flags["tap"], ndcg=flags["ndcg"], hit=flags["hit"], rr=flags["rr"],
          roc_auc=flags["roc"], pr_auc=flags["pr"])
A, B = d["A"], d["B"]
if a.f64:
    A, B = A.astype(np.float64), B.astype(np.float64)
F = synth.algorithmic_flops(cfg, d["X_train"], d["X_test"])
for i in range(a.reps):
    r = rb.calc_reco_metrics_ex(d["X_train"], d["X_test"], A, B, k=a.k or cfg.k, item_biases=d["item_biases"],
                                cumulative=cfg.cumulative, break_ties_with_noise=a.noise, min_pos_test=cfg.min_pos_test,
                                devices=[int(x) for x in a.devices.split(",")] if a.devices else None, **kw)
    t = r.timing
    print(json.dumps({"rep": i, "dom_ms": round(t["dominant_kernel_ms"], 3), "users": a.users, "kernel_ms": t["score_select_ms"], "tflops": F / t["score_select_ms"] / 1e9,
                      "prep_ms": t["prep_ms"], "metrics_ms": t["metrics_ms"], "h2d_ms": t["h2d_ms"], "d2h_ms": t["d2h_ms"],
                      "total_ms": t["total_ms"], "devices_used": t["devices_used"], "users_per_s_e2e": a.users / t["total_ms"] * 1e3, "users_per_s_kernel": a.users / t["score_select_ms"] * 1e3,
                      "dom_ms": t["dominant_kernel_ms"], "retry_rows": t.get("filter_retry_rows"), "fallback": t.get("filter_fallback_batches")}))
