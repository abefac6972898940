"""Host logic of the band-chunk pipeline (no GPU): grouping of variables by host buffer, band-chunk
schedule, and which targets of a group share one two-method gather launch (``xrs_gather_ij2``)."""

import numpy as np

from xcube_resampling_b200._pipeline import Target, chunk_schedule, group_by_buffer, pair_targets


def _t(name, method, dtype=np.float32):
    return Target(name, method, 0.0, np.empty((1, 2, 2), dtype=dtype))


def test_pair_targets_pairs_nearest_with_an_interpolating_method():
    a, b = _t("a", "nearest"), _t("b", "bilinear")
    assert pair_targets([a, b], np.float32) == [(b, a)]
    assert pair_targets([b, a], np.float32) == [(b, a)]
    c = _t("c", "triangular")
    assert pair_targets([a, c], np.float32) == [(c, a)]


def test_pair_targets_leaves_the_rest_alone_and_keeps_order():
    n1, n2, b1, t1, b2 = _t("n1", "nearest"), _t("n2", "nearest"), _t("b1", "bilinear"), _t("t1", "triangular"), \
        _t("b2", "bilinear")
    jobs = pair_targets([n1, b1, t1, n2, b2], np.float32)
    assert jobs == [(b1, n1), (t1, n2), (b2,)]
    assert pair_targets([n1, n2], np.float32) == [(n1,), (n2,)]
    assert pair_targets([b1, b2], np.float32) == [(b1,), (b2,)]
    assert pair_targets([], np.float32) == []


def test_pair_targets_needs_the_source_dtype_on_both_sides():
    n, b = _t("n", "nearest", np.uint8), _t("b", "bilinear", np.float32)
    assert pair_targets([n, b], np.float32) == [(n,), (b,)]
    assert pair_targets([n, b], np.uint8) == [(n,), (b,)]


def test_every_target_appears_exactly_once():
    rng = np.random.default_rng(0)
    methods = ["nearest", "bilinear", "triangular"]
    for _ in range(200):
        ts = [_t(str(i), methods[rng.integers(3)], [np.float32, np.int16][rng.integers(2)]) for i in range(rng.integers(1, 8))]
        jobs = pair_targets(ts, np.float32)
        flat = [t for job in jobs for t in job]
        assert sorted(map(id, flat)) == sorted(map(id, ts))
        for job in jobs:
            if len(job) == 2:
                assert job[0].method in ("bilinear", "triangular") and job[1].method == "nearest"
                assert job[0].out_dtype == job[1].out_dtype == np.float32


def test_group_by_buffer_and_chunk_schedule():
    data = np.zeros((5, 4, 6), np.float32)
    other = np.zeros((5, 4, 6), np.float32)
    groups = group_by_buffer([(data, _t("a", "nearest")), (other, _t("c", "nearest")), (data, _t("b", "bilinear"))])
    assert [len(g.targets) for g in groups] == [2, 1]
    assert groups[0].values is data or groups[0].values.base is data
    assert chunk_schedule(21, 4) == [(0, 1), (1, 2), (3, 4), (7, 4), (11, 4), (15, 4), (19, 2)]
    assert chunk_schedule(1, 4) == [(0, 1)]
