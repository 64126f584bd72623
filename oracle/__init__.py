"""TEST INFRASTRUCTURE -- not part of the product.

ctypes doors onto
  * ``oracle/librmoracle.so``             -- the C restatement (recometrics_oracle.c)
  * ``oracle/_ref/librecometrics_ref.so`` -- the unmodified reference compiled by oracle/Makefile

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  ``recometrics_b200`` never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "librmoracle.so")
REF_SO = os.path.join(_HERE, "_ref", "librecometrics_ref.so")

METRICS = ("p", "tp", "r", "ap", "tap", "ndcg", "hit", "rr", "roc", "pr")
ALL_METRICS = METRICS
_TOPK = METRICS[:8]

_vp = ctypes.c_void_p
_i32 = ctypes.c_int32
_int = ctypes.c_int
_sz = ctypes.c_size_t
_u64 = ctypes.c_uint64


def build(force=False):
    """Compile the restatement (always possible) and, when /root/reference exists, oracle/_ref."""
    srcs = [os.path.join(_HERE, f) for f in ("recometrics_oracle.c", "recometrics_oracle_impl.h", "split_oracle.c", "ref_shim.cpp")]
    stale = os.path.exists(ORACLE_SO) and any(os.path.getmtime(f) > os.path.getmtime(ORACLE_SO) for f in srcs[:3])
    ref_stale = os.path.isdir("/root/reference/src") and (not os.path.exists(REF_SO) or os.path.getmtime(srcs[3]) > os.path.getmtime(REF_SO))
    if force or stale or ref_stale or not os.path.exists(ORACLE_SO):
        subprocess.run(["make", "-B", "-C", _HERE], check=True, capture_output=True)


def have_ref():
    return os.path.exists(REF_SO)


REF_V4_SO = os.path.join(_HERE, "_ref", "librecometrics_ref_v4.so")
_timing_so = None


def _cpu_has_avx512():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("flags"):
                f = set(ln.split(":", 1)[1].split())
                return {"avx512f", "avx512bw", "avx512vl", "avx512dq"} <= f
    except OSError:
        pass
    return False


def use_best_ref_build_for_timing():
    """bench.py only: time the reference build closest to `-march=native` on THIS host (the x86-64-v4 build when the CPU has
    AVX-512, else the portable x86-64-v3 one).  Parity tests never call this: they stay on the v3 build the goldens came from."""
    global _timing_so
    _timing_so = REF_V4_SO if (os.path.exists(REF_V4_SO) and _cpu_has_avx512()) else REF_SO
    return ref_build_note()


def ref_build_note():
    so = _timing_so or REF_SO
    return "oracle/_ref/%s (g++ -O3 -fopenmp %s; the reference's setup.py flags with a portable -march)" % (
        os.path.basename(so), "-march=x86-64-v4 -mprefer-vector-width=256" if so == REF_V4_SO else "-march=x86-64-v3")


_libs = {}


def _lib(path):
    if path not in _libs:
        _libs[path] = ctypes.CDLL(path)
    return _libs[path]


def _ptr(a):
    return _vp(a.ctypes.data) if (a is not None and a.size) else _vp(None)


def _prep(A, B, Xtr, Xte, dtype):
    A = np.ascontiguousarray(A, dtype=dtype)
    B = np.ascontiguousarray(B, dtype=dtype)
    trp = np.ascontiguousarray(Xtr.indptr, dtype=np.int32)
    tri = np.ascontiguousarray(Xtr.indices, dtype=np.int32)
    tep = np.ascontiguousarray(Xte.indptr, dtype=np.int32)
    tei = np.ascontiguousarray(Xte.indices, dtype=np.int32)
    tev = np.ascontiguousarray(Xte.data, dtype=dtype)
    return A, B, trp, tri, tep, tei, tev


def _alloc_outs(metrics, m, K, cumulative, dtype):
    outs = {}
    for name in METRICS:
        if name in metrics:
            size = m * K if (cumulative and name in _TOPK) else m
            # poison so that "left untouched" is visible
            outs[name] = np.full(size, -12345.0, dtype=dtype)
        else:
            outs[name] = None
    return outs


def _shape_outs(outs, m, K, cumulative):
    res = {}
    for name, a in outs.items():
        if a is None:
            continue
        res[name] = a.reshape(m, K) if (cumulative and name in _TOPK) else a
    return res


def ref_calc(A, B, Xtr, Xte, k, metrics=("p", "ap", "ndcg"), cumulative=False,
             consider_cold_start=True, min_items_pool=2, min_pos_test=1, nthreads=1,
             break_ties_with_noise=False, seed=1, dtype=np.float32):
    """Run the UNMODIFIED reference (calc_metrics_float/_double, recometrics_signatures.hpp:46-98)."""
    lib = _lib(_timing_so or REF_SO)
    A, B, trp, tri, tep, tei, tev = _prep(A, B, Xtr, Xte, dtype)
    m, n, p = A.shape[0], B.shape[0], A.shape[1]
    outs = _alloc_outs(metrics, m, k, cumulative, dtype)
    fn = lib.rmref_calc_metrics_f32 if dtype == np.float32 else lib.rmref_calc_metrics_f64
    fn.restype = _int
    rc = fn(_ptr(A), _sz(A.shape[1]), _ptr(B), _sz(B.shape[1]), _i32(m), _i32(n), _i32(p),
            _ptr(trp), _ptr(tri), _ptr(tep), _ptr(tei), _ptr(tev),
            _i32(k), _int(int(cumulative)), _int(int(break_ties_with_noise)),
            *[_ptr(outs[q]) for q in METRICS],
            _int(int(consider_cold_start)), _i32(min_items_pool), _i32(min_pos_test),
            _i32(nthreads), _u64(seed))
    if rc != 0:
        raise RuntimeError("reference threw")
    return _shape_outs(outs, m, k, cumulative)


def oracle_calc(A, B, Xtr, Xte, k, metrics=("p", "ap", "ndcg"), cumulative=False,
                consider_cold_start=True, min_items_pool=2, min_pos_test=1, nthreads=1,
                fix_quirks=True, extras=False, break_ties_with_noise=False, seed=1, dtype=np.float32):
    """Run the C restatement.  With extras=True also returns status / top-K ids+scores / ranks."""
    lib = _lib(ORACLE_SO)
    A, B, trp, tri, tep, tei, tev = _prep(A, B, Xtr, Xte, dtype)
    m, n, p = A.shape[0], B.shape[0], A.shape[1]
    outs = _alloc_outs(metrics, m, k, cumulative, dtype)
    status = np.zeros(m, dtype=np.int32)
    topk_items = np.zeros(m * k, dtype=np.int32) if extras else None
    topk_scores = np.zeros(m * k, dtype=dtype) if extras else None
    pos_rank = np.zeros(max(int(tep[-1]), 1), dtype=np.int64) if extras else None
    tie_flags = np.zeros(m, dtype=np.int32) if extras else None
    fn = lib.rmo_calc_metrics_f32 if dtype == np.float32 else lib.rmo_calc_metrics_f64
    fn.restype = _int
    rc = fn(_ptr(A), _sz(A.shape[1]), _ptr(B), _sz(B.shape[1]), _i32(m), _i32(n), _i32(p),
            _ptr(trp), _ptr(tri), _ptr(tep), _ptr(tei), _ptr(tev),
            _i32(k), _int(int(cumulative)),
            *[_ptr(outs[q]) for q in METRICS],
            _int(int(consider_cold_start)), _i32(min_items_pool), _i32(min_pos_test),
            _i32(nthreads), _int(int(fix_quirks)), _int(int(break_ties_with_noise)), _u64(seed),
            _ptr(status), _ptr(topk_items), _ptr(topk_scores), _ptr(pos_rank), _ptr(tie_flags))
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    res = _shape_outs(outs, m, k, cumulative)
    if extras:
        res["status"] = status
        res["topk_items"] = topk_items.reshape(m, k)
        res["topk_scores"] = topk_scores.reshape(m, k)
        res["pos_rank"] = pos_rank[: int(tep[-1])]
        res["tie_flags"] = tie_flags
    return res


# ---------------------------------------------------------------------------------------------------------------------
# Train/test splitters (split_oracle.c; reference: src/recometrics.hpp:1015-1505).  Both doors take the CSR arrays of X
# and return plain numpy arrays: {"users_test", "train": (p, i, v), "test": (p, i, v), "rem": (p, i, v) or None}.
# ---------------------------------------------------------------------------------------------------------------------
SPLIT_ERRORS = {1: "Target number of test users is larger than available users.\n",
                2: "Selected minimum number of items is larger than total number of items.\n",
                3: "No users satisfy criteria for test inclusion.\n"}


def _split_bufs(m, nnz, dtype):
    mk = lambda size, dt: np.full(max(size, 1), -7, dtype=dt)   # noqa: E731  (poison: "left untouched" is visible)
    return dict(ut=mk(m, np.int32), rep=mk(m + 1, np.int32), rei=mk(nnz, np.int32), rev=mk(nnz, dtype),
                trp=mk(m + 1, np.int32), tri=mk(nnz, np.int32), trv=mk(nnz, dtype),
                tep=mk(m + 1, np.int32), tei=mk(nnz, np.int32), tev=mk(nnz, dtype))


def _split_args(Xp, Xi, Xv):
    Xv = np.ascontiguousarray(Xv)
    assert Xv.dtype in (np.float32, np.float64)
    return np.ascontiguousarray(Xp, dtype=np.int32), np.ascontiguousarray(Xi, dtype=np.int32), Xv


def _csr3(p, i, v, rows, nnz=None):
    p = p[: rows + 1].copy()
    nnz = int(p[-1]) if nnz is None else nnz
    return p, i[:nnz].copy(), v[:nnz].copy()


def oracle_split(Xp, Xi, Xv, m, n, split_type="all", n_users_test=0, test_fraction=0.3, consider_cold_start=False,
                 min_items_pool=2, min_pos_test=1, seed=1):
    """The C restatement.  The float32 entry points of the reference take the fraction as a float
    (src/recometrics_instantiated.cpp:192, :267, :337): the same rounding is applied here."""
    lib = _lib(ORACLE_SO)
    Xp, Xi, Xv = _split_args(Xp, Xi, Xv)
    frac = float(np.float32(test_fraction)) if Xv.dtype == np.float32 else float(test_fraction)
    b = _split_bufs(m, Xi.size, Xv.dtype)
    vsz = Xv.dtype.itemsize
    if split_type == "all":
        lib.rmo_split_selected_users.restype = _int
        rc = lib.rmo_split_selected_users(_ptr(Xp), _ptr(Xi), _ptr(Xv), _int(vsz), _i32(m), _i32(n), ctypes.c_double(frac),
                                          _u64(seed), _ptr(b["trp"]), _ptr(b["tri"]), _ptr(b["trv"]),
                                          _ptr(b["tep"]), _ptr(b["tei"]), _ptr(b["tev"]))
        if rc:
            raise RuntimeError("Passed negative dimensions.\n")
        return {"users_test": None, "train": _csr3(b["trp"], b["tri"], b["trv"], m), "test": _csr3(b["tep"], b["tei"], b["tev"], m),
                "rem": None}
    joined = split_type == "joined"
    sizes = np.zeros(2, dtype=np.int64)
    lib.rmo_split_users.restype = _int
    rc = lib.rmo_split_users(_ptr(Xp), _ptr(Xi), _ptr(Xv), _int(vsz), _i32(m), _i32(n), _i32(n_users_test), ctypes.c_double(frac),
                             _int(int(consider_cold_start)), _i32(min_items_pool), _i32(min_pos_test), _u64(seed), _int(int(joined)),
                             _ptr(b["ut"]), _ptr(b["rep"]), _ptr(b["rei"]), _ptr(b["rev"]),
                             _ptr(b["trp"]), _ptr(b["tri"]), _ptr(b["trv"]), _ptr(b["tep"]), _ptr(b["tei"]), _ptr(b["tev"]), _ptr(sizes))
    if rc:
        raise RuntimeError(SPLIT_ERRORS.get(rc, "split oracle failed (%d)" % rc))
    taken, others = int(sizes[0]), int(sizes[1])
    return {"users_test": b["ut"][:taken].copy(),
            "train": _csr3(b["trp"], b["tri"], b["trv"], taken + (others if joined else 0)),
            "test": _csr3(b["tep"], b["tei"], b["tev"], taken),
            "rem": None if joined else _csr3(b["rep"], b["rei"], b["rev"], others)}


def ref_split(Xp, Xi, Xv, m, n, split_type="all", n_users_test=0, test_fraction=0.3, consider_cold_start=False,
              min_items_pool=2, min_pos_test=1, seed=1):
    """The UNMODIFIED reference (split_data_*_float/_double, src/recometrics_signatures.hpp:100-221)."""
    lib = _lib(REF_SO)
    Xp, Xi, Xv = _split_args(Xp, Xi, Xv)
    sfx = "f32" if Xv.dtype == np.float32 else "f64"
    b = _split_bufs(m, Xi.size, Xv.dtype)
    sizes = np.zeros(7, dtype=np.int64)
    err = ctypes.create_string_buffer(256)
    if split_type == "all":
        fn = getattr(lib, "rmref_split_selected_users_" + sfx)
        fn.restype = _int
        rc = fn(_ptr(Xp), _ptr(Xi), _ptr(Xv), _i32(m), _i32(n), ctypes.c_double(test_fraction), _u64(seed),
                _ptr(b["trp"]), _ptr(b["tri"]), _ptr(b["trv"]), _ptr(b["tep"]), _ptr(b["tei"]), _ptr(b["tev"]), _ptr(sizes), err)
    else:
        fn = getattr(lib, "rmref_split_users_" + sfx)
        fn.restype = _int
        rc = fn(_ptr(Xp), _ptr(Xi), _ptr(Xv), _i32(m), _i32(n), _i32(n_users_test), ctypes.c_double(test_fraction),
                _int(int(consider_cold_start)), _i32(min_items_pool), _i32(min_pos_test), _u64(seed), _int(int(split_type == "joined")),
                _ptr(b["ut"]), _ptr(b["rep"]), _ptr(b["rei"]), _ptr(b["rev"]),
                _ptr(b["trp"]), _ptr(b["tri"]), _ptr(b["trv"]), _ptr(b["tep"]), _ptr(b["tei"]), _ptr(b["tev"]), _ptr(sizes), err)
    if rc:
        raise RuntimeError(err.value.decode())
    nut, nrp, nri, ntp, nti, nep, nei = (int(x) for x in sizes)
    return {"users_test": b["ut"][:nut].copy() if split_type != "all" else None,
            "train": _csr3(b["trp"], b["tri"], b["trv"], ntp - 1, nti),
            "test": _csr3(b["tep"], b["tei"], b["tev"], nep - 1, nei),
            "rem": _csr3(b["rep"], b["rei"], b["rev"], nrp - 1, nri) if split_type == "separated" else None}
