"""Generates tests/golden/split_*.npz: the outputs of the UNMODIFIED reference's splitters (oracle/_ref, compiled by
oracle/Makefile from /root/reference/src) on the seeded inputs of tools/split_cases.py.  Run in the build container:
    python tests/golden/make_golden_split.py
Arrays above 64K elements are stored as their SHA-256 (the inputs are regenerated from their seeds, never stored)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle                      # noqa: E402
from tools import split_cases      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def pack(flat):
    out = {}
    for k, a in flat.items():
        if a.size > 65536:
            out["sha256_" + k] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
            out["len_" + k] = np.array([a.size], dtype=np.int64)
        else:
            out[k] = a
    return out


if __name__ == "__main__":
    oracle.build()
    assert oracle.have_ref(), "needs oracle/_ref (the compiled reference)"
    for name, (mk, kw) in split_cases.CASES.items():
        p, i, v = split_cases.make_csr(**mk)
        res = oracle.ref_split(p, i, v, mk["m"], mk["n"], **kw)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **pack(split_cases.flatten(res)))
        print(name, {k: a.shape for k, a in split_cases.flatten(res).items()})
