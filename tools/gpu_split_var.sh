for k in 1 2 3; do timeout 100 python tools/split_once.py --no-reference --reps 5 | tail -1 | cut -c60-330; done
RMB200_SPLIT_DBG=1 timeout 100 python - <<'PY'
import sys, time
sys.path.insert(0, '.')
from recometrics_b200 import _capi
from tools.split_once import power_law_matrix
p, i, v = power_law_matrix(138493, 26744, 144.0)
for rep in range(8):
    t0 = time.perf_counter()
    r = _capi.split("all", p, i, v, 138493, 26744, test_fraction=0.3, seed=1)
    w = (time.perf_counter() - t0) * 1e3
    t = r["timing"]
    print("rep %d wall %.1f total %.1f plan %.1f h2d %.1f d2h %.1f" % (rep, w, t["total_ms"], t["plan_ms"], t["h2d_ms"], t["d2h_ms"]))
    del r
PY
