"""GPU: ONE call spread over several GPUs (rmb200_extra_t::devices / env RMB200_DEVICES) -- the counterpart of the
reference's one call using every core (/root/reference/src/recometrics.hpp:428-437).  Users go to the devices in
contiguous blocks, the item factors are uploaded in slices and exchanged GPU-to-GPU; every row must be what the
single-GPU call writes, bit for bit.  The two-device cases skip on a one-GPU box (`gpurun --gpus 2` runs them)."""
import os

import numpy as np
import pytest

from tools import synth

pytestmark = pytest.mark.gpu

ALL = dict(all_metrics=True)
TOPK4 = dict(precision=True, recall=True, average_precision=True, ndcg=True)


def _same(a, b):
    assert np.array_equal(a.status, b.status)
    assert np.array_equal(a.topk_items, b.topk_items)
    assert np.array_equal(a.topk_scores, b.topk_scores, equal_nan=True)
    for key, v in a.metrics.items():
        if key != "K":
            assert np.array_equal(v, b.metrics[key], equal_nan=True), key


def _run(rb, d, k, devices=None, **kw):
    return rb.calc_reco_metrics_ex(d["X_train"], d["X_test"], d["A"], d["B"], k=k, item_biases=d["item_biases"],
                                   break_ties_with_noise=False, return_topk=True, return_status=True,
                                   devices=devices, **kw)


def test_a_list_of_one_device_is_that_device(rb):
    d = synth.make(1, m=700, n=1500, p=32)
    a = _run(rb, d, 10, **ALL)
    b = _run(rb, d, 10, devices=[0], **ALL)
    _same(a, b)
    assert b.timing["devices_used"] == 1


def test_bad_device_lists_are_refused(rb):
    d = synth.make(1, m=100, n=300, p=8)
    for devs in ([0, 0], [0, 99], [-1]):
        with pytest.raises(ValueError):
            _run(rb, d, 5, devices=devs, **TOPK4)


def _need(rb, n):
    if rb.device_count() < n:
        pytest.skip("needs %d GPUs" % n)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_two_devices_equal_one_device_all_metrics(rb, dtype):
    """FMA tiles with rank counting (all ten metrics), users not a multiple of the 128-user unit."""
    _need(rb, 2)
    d = synth.make(3, m=3001, n=9000, p=48)
    d["A"], d["B"] = d["A"].astype(dtype), d["B"].astype(dtype)
    a = _run(rb, d, 20, return_ranks=True, return_means=True, **ALL)
    b = _run(rb, d, 20, devices=[0, 1], return_ranks=True, return_means=True, **ALL)
    _same(a, b)
    assert np.array_equal(a.pos_rank, b.pos_rank)
    assert b.timing["devices_used"] == 2
    for key, v in a.means.items():                     # means: the devices' means weighted by their counts
        assert a.counts[key] == b.counts[key], key
        assert np.allclose(v, b.means[key], rtol=1e-12, atol=0, equal_nan=True), key


def test_two_devices_equal_one_device_tensor_path_with_bias_and_cumulative(rb):
    """Tensor-core filter path (top-K metrics only), biases, cumulative rows, a catalogue with the sampled threshold guess."""
    _need(rb, 2)
    d = synth.make(2, m=2500, n=70001, p=64)
    a = _run(rb, d, 10, cumulative=True, **TOPK4)
    b = _run(rb, d, 10, devices=[1, 0], cumulative=True, **TOPK4)
    assert a.timing["scoring_path"] == 2 and b.timing["scoring_path"] == 2
    _same(a, b)


def test_two_devices_user_range_and_fewer_units_than_devices(rb):
    """A user range inside the call is split as well; 100 users are one 128-user unit -> one device does all the work."""
    _need(rb, 2)
    d = synth.make(1, m=900, n=2000, p=16)
    a = _run(rb, d, 10, user_range=(130, 777), **TOPK4)
    b = _run(rb, d, 10, devices=[0, 1], user_range=(130, 777), **TOPK4)
    _same(a, b)
    d = synth.make(1, m=100, n=2000, p=16)
    a = _run(rb, d, 10, **TOPK4)
    b = _run(rb, d, 10, devices=[0, 1], **TOPK4)
    _same(a, b)
    assert b.timing["devices_used"] == 1


def test_reference_named_entry_uses_every_gpu_through_the_environment(rb, monkeypatch):
    """The reference's bindings know nothing about devices: RMB200_DEVICES=all spreads their unchanged call."""
    _need(rb, 2)
    d = synth.make(1, m=1500, n=2500, p=32)
    a = rb.calc_reco_metrics(d["X_train"], d["X_test"], d["A"], d["B"], k=10, all_metrics=True, break_ties_with_noise=False, as_df=False)
    monkeypatch.setenv("RMB200_DEVICES", "all")
    b = rb.calc_reco_metrics(d["X_train"], d["X_test"], d["A"], d["B"], k=10, all_metrics=True, break_ties_with_noise=False, as_df=False)
    for key, v in a.items():
        if key != "K":
            assert np.array_equal(v, b[key], equal_nan=True), key


def test_pageable_item_factors_go_up_in_slices(rb, monkeypatch):
    """Force the host-thread upload pipeline on every slice (pageable numpy memory is what the reference's callers pass)."""
    _need(rb, 2)
    monkeypatch.setenv("RMB200_UPLOAD_MIN_BYTES", "1")
    d = synth.make(4, m=1000, n=30000, p=40)
    a = _run(rb, d, 50, **TOPK4)
    b = _run(rb, d, 50, devices=[0, 1], **TOPK4)
    _same(a, b)


def test_reference_cython_wrapper_spreads_its_call_over_the_gpus(rb, tmp_path, monkeypatch):
    """The reference's UNMODIFIED Cython wrapper (oracle/build_ref_cython.py) with RMB200_DEVICES=all in the environment: its one
    calc_metrics_float call runs on every GPU of the box and returns the one-GPU rows bit for bit (m = 300 is three 128-user
    units: two devices get work)."""
    _need(rb, 2)
    import test_capi_host
    monkeypatch.setenv("RMB200_DEVICES", "all")
    stdout, path = test_capi_host.run_reference_cython_wrapper(tmp_path)
    assert "HAS 1" in stdout and "COMPUTED" in stdout, stdout
    rows = np.load(path)
    monkeypatch.delenv("RMB200_DEVICES")
    d = synth.make(1, m=300, n=500, p=8)
    r = rb.calc_reco_metrics_ex(d["X_train"], d["X_test"], d["A"], d["B"], k=5, precision=True, average_precision=True, ndcg=True,
                                break_ties_with_noise=False)
    for i, key in enumerate(("P@K", "AP@K", "NDCG@K")):
        assert np.array_equal(rows[i], r.metrics[key], equal_nan=True), key
