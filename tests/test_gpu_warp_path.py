"""The warp-per-query exact search (csrc/warp_core.cuh) — what the tile pass and the block search
hand their unfinished queries to — run over EVERY query (tuning "warp_all") and compared with the
fixtures of the unmodified reference and with the oracle: bit-exact rows (indices, distances,
counts) and mean distances, normals within 1 - |cos| <= 1e-4 where the eigengap is > 1e-3.
Clouds: the reference fixtures (incl. the exact-tie lattice and the near-duplicates), outliers,
k > n, a voxel grid that leaves points unindexed, external queries far outside the cloud."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIX = os.path.join(os.path.dirname(__file__), "golden", "ref_fixtures.npz")
PAD = 0xFFFFFFFF


def as_i64(idx):
    out = idx.astype(np.int64)
    out[idx == PAD] = -1
    return out


@pytest.fixture()
def warp_all(pcpx):
    pcpx.set_tuning("warp_all", 1)
    yield pcpx
    pcpx.set_tuning("warp_all", 0)


@pytest.mark.parametrize("name", ["sphere", "cube", "plane", "lattice", "dup"])
def test_warp_path_reference_fixtures(warp_all, name):
    pcpx, fix = warp_all, np.load(FIX)
    xyz, q = fix[name + "_xyz"], fix[name + "_queries"]
    with pcpx.Index(xyz) as ix:
        for k in (1, 8, 15):
            for qq, qn in ((None, "self"), (q, "ext")):
                idx, d2, cnt = ix.knn(qq, k)
                assert ix.timings()["deferred_queries"] == len(xyz if qq is None else qq)
                assert np.array_equal(as_i64(idx), fix["%s_%s_k%d_idx" % (name, qn, k)].astype(np.int64))
                assert np.array_equal(d2, fix["%s_%s_k%d_d2" % (name, qn, k)])
        per, mean = ix.mean_knn_distance(15)
        assert np.array_equal(per, fix[name + "_mean15"], equal_nan=True)


@pytest.mark.parametrize("k", [1, 5, 15, 16, 31, 32])
def test_warp_path_against_the_oracle(warp_all, oracle, k):
    pcpx = warp_all
    rng = np.random.default_rng(100 + k)
    slab = rng.uniform(0, 1, (30_000, 3)).astype(np.float32) * np.array([1, 1, 0.02], np.float32)
    stray = rng.uniform(-3, 4, (600, 3)).astype(np.float32)          # outliers far from the slab
    dup = slab[:300] + np.float32(2e-6)                               # inside the exclusion box
    xyz = np.concatenate([slab, stray, dup, slab[:100]])
    q = rng.uniform(-5, 6, (1_500, 3)).astype(np.float32)
    oc = oracle.cloud(xyz)
    with pcpx.Index(xyz) as ix:
        for qq in (None, q):
            idx, d2, cnt = ix.knn(qq, k)
            oi, od2, ocnt = oc.knn(qq, k)
            assert np.array_equal(as_i64(idx), oi) and np.array_equal(d2, od2) and np.array_equal(cnt, ocnt)
        per, _ = ix.mean_knn_distance(k)
        assert np.array_equal(per, oc.mean_knn_distance(k)[0], equal_nan=True)
        if k >= 5:
            nrm = ix.estimate_normals(None, k)
            onrm, gap = oc.normals(None, k)
            err = 1 - np.abs((nrm * onrm).sum(1))
            assert err[gap > 1e-3].max() <= 1e-4


def test_warp_path_edge_cases(warp_all, oracle):
    pcpx = warp_all
    rng = np.random.default_rng(3)
    # k > n: rows padded, counts = what exists
    tiny = rng.uniform(0, 1, (7, 3)).astype(np.float32)
    with pcpx.Index(tiny) as ix:
        idx, d2, cnt = ix.knn(None, 10)
        oi, od2, ocnt = oracle.cloud(tiny).knn(None, 10)
        assert np.array_equal(as_i64(idx), oi) and np.array_equal(d2, od2) and np.array_equal(cnt, ocnt)
    # a voxel grid that leaves a third of the points unindexed
    xyz = rng.uniform(-1, 1, (20_000, 3)).astype(np.float32)
    box = (np.array([-1, -1, -1], np.float32), np.array([1, 1, 0.3], np.float32))
    oc = oracle.cloud(xyz, bbox=box)
    with pcpx.Index(xyz, voxel_grid=box) as ix:
        idx, d2, cnt = ix.knn(None, 12)
        oi, od2, ocnt = oc.knn(None, 12)
        assert np.array_equal(as_i64(idx), oi) and np.array_equal(d2, od2) and np.array_equal(cnt, ocnt)


def test_handed_on_queries_take_the_warp_path(pcpx, oracle):
    """Default flow on a cloud that mixes a surface with uniform noise: the tile pass finishes the
    surface, everything it hands on (sparse tiles, outliers) goes through the warp kernel; the
    same call with the per-thread retry kernels must agree bit for bit."""
    synth = __import__("importlib").import_module("point-cloud-processing_b200.synth")
    xyz = synth.noise_mix(200_000, seed=3)
    oc = oracle.cloud(xyz)
    res = {}
    for mode in (1, 0):
        pcpx.set_tuning("warp_retry", mode)
        with pcpx.Index(xyz) as ix:
            res[mode] = ix.knn(None, 15) + (ix.mean_knn_distance(15)[0],)
            t = ix.timings()
            assert t["deferred_queries"] > 0
    pcpx.set_tuning("warp_retry", 1)
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b, equal_nan=True)
    sample = np.random.default_rng(0).choice(len(xyz), 5_000, replace=False)
    oi, od2, _ = oc.knn(None, 15)
    assert np.array_equal(as_i64(res[1][0][sample]), oi[sample])
    assert np.array_equal(res[1][1][sample], od2[sample])
