"""CPU, world_size 2, gloo: the N>1 plumbing -- block partition of the users and the gather of
per-user metric rows to rank 0 (recometrics_b200/dist.py).  Each rank's rows come from the oracle
here (the GPU evaluator cannot run on this machine); the GPU version of this test is in
tests/test_gpu_parity.py::test_user_range_shards_equal_full_call."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import oracle
    from recometrics_b200.dist import gather_metric_rows, shard_bounds
    from tools import synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = synth.make(1, m=301, n=400, p=8, k=5)
        m = 301
        b, e = shard_bounds(m, world, rank)
        sub = dict(A=d["A"][b:e], X_train=d["X_train"][b:e], X_test=d["X_test"][b:e])
        o = oracle.oracle_calc(sub["A"], d["B"], sub["X_train"], sub["X_test"], 5, metrics=("p", "ndcg", "roc"), nthreads=1)
        oc = oracle.oracle_calc(sub["A"], d["B"], sub["X_train"], sub["X_test"], 5, metrics=("ap",), cumulative=True, nthreads=1)
        local = {"P@K": o["p"], "NDCG@K": o["ndcg"], "ROC_AUC": o["roc"], "AP@K": oc["ap"]}
        full = gather_metric_rows(local, m, dst=0)
        if rank == 0:
            ref = oracle.oracle_calc(d["A"], d["B"], d["X_train"], d["X_test"], 5, metrics=("p", "ndcg", "roc"), nthreads=1)
            refc = oracle.oracle_calc(d["A"], d["B"], d["X_train"], d["X_test"], 5, metrics=("ap",), cumulative=True, nthreads=1)
            ok = True
            for key, want in (("P@K", ref["p"]), ("NDCG@K", ref["ndcg"]), ("ROC_AUC", ref["roc"]), ("AP@K", refc["ap"])):
                got = full[key]
                ok &= got.shape == want.shape and bool(np.all((got == want) | (np.isnan(got) & np.isnan(want))))
            q.put(("ok" if ok else "mismatch"))
        else:
            assert full is None
            q.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gather_rows_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert res == ["ok", "ok"], res
