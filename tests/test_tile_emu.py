"""CPU-only check of the tile path (point-cloud-processing_b200/csrc/tile_core.cuh): tests/emu runs
the SAME staging phases and per-query search the tile kernel runs, one emulated CTA per tile, with
the placement "atomics" served in a scrambled thread order.  Every row the tile pass declares final
must be bit-equal to the exact (d2, original index) search — itself pinned to the oracle and the
reference fixtures by test_emu_parity.py — and rows it does not finish must be flagged (the
product sends those to the retry queue)."""
import importlib
import os

import numpy as np
import pytest

synth = importlib.import_module("point-cloud-processing_b200.synth")
FIX = os.path.join(os.path.dirname(__file__), "golden", "ref_fixtures.npz")
PAD = 0xFFFFFFFF


def clouds():
    rng = np.random.default_rng(5)
    plane = synth.noisy_plane(60_000, seed=7)
    sphere = synth.noisy_sphere(40_000, seed=42)
    mix = synth.noise_mix(50_000, seed=11)
    # exact ties: a jittered lattice snapped to a coarse float grid, plus exact duplicates
    g = np.stack(np.meshgrid(np.arange(40), np.arange(40), np.arange(3), indexing="ij"), -1)
    lattice = (g.reshape(-1, 3) * np.float32(0.25)).astype(np.float32)
    dup = np.concatenate([plane[:20_000], plane[:2_000] + np.float32(3e-6), plane[:500]])
    tilted = plane[:50_000] @ np.array([[0.8, 0.0, 0.6], [0.0, 1.0, 0.0], [-0.6, 0.0, 0.8]],
                                       np.float32).T
    cube = rng.uniform(-1, 1, (30_000, 3)).astype(np.float32)
    return dict(plane=plane, sphere=sphere, mix=mix, lattice=lattice, dup=dup,
                tilted=np.ascontiguousarray(tilted), cube=cube)


@pytest.fixture(scope="module")
def cloud_set():
    return clouds()


def tile_levels(ix, k):
    info = ix.info()
    lv = ix.plan(k)["level"]
    return sorted({max(2, min(info["lfine"], l)) for l in (lv, lv - 1)})


@pytest.mark.parametrize("name", ["plane", "sphere", "mix", "lattice", "dup", "tilted", "cube"])
def test_tile_rows_are_exact(emu, cloud_set, name):
    xyz = cloud_set[name]
    ix = emu.index(xyz)
    if ix.info()["lfine"] < 2:
        pytest.skip("cloud too small for a tile level")
    for k in (1, 8, 15, 30):
        ridx, rd2, rcnt, _ = ix.knn(None, k, exact_only=True)
        rnrm, rctr, rmean, _ = ix.normals(None, k, want_means=True, exact_only=True)
        for level in tile_levels(ix, k):
            for sub, alg in ((1, 2), (2, 2), (4, 2)):
                r = ix.tile(k, 0, level, sub=sub, alg=alg, max_points=4096)
                d = r["done"]
                assert np.array_equal(r["idx"][d], ridx[d]), (name, k, level, sub)
                assert np.array_equal(r["d2"][d], rd2[d]), (name, k, level, sub)
                assert np.all(r["cnt"][d] == k)
                assert np.all(rcnt[d] == k)
                m = ix.tile(k, 1, level, sub=sub, alg=alg, max_points=4096)
                assert np.array_equal(m["means"][m["done"]], rmean[m["done"]]), (name, k, level, sub)
                if k >= 3:
                    nr = ix.tile(k, 2, level, sub=sub, alg=alg, max_points=4096)
                    dn = nr["done"]
                    cos = np.abs((nr["normals"][dn] * rnrm[dn]).sum(1))
                    # same neighbour set; moments taken about the query point in both
                    assert np.all(1 - cos[np.isfinite(cos)] <= 1e-4) or name in ("lattice", "cube")
                    assert np.allclose(nr["centroids"][dn], rctr[dn], atol=1e-5 * np.abs(xyz).max())


def test_tile_finishes_most_of_a_surface(emu, cloud_set):
    for name, k in (("plane", 15), ("sphere", 15), ("tilted", 8)):
        ix = emu.index(cloud_set[name])
        for sub in (2, 4):
            r = ix.tile(k, 0, ix.plan(k)["level"], sub=sub, max_points=4096)
            assert r["stats"]["fallback_tiles"] == 0
            assert r["done"].mean() > 0.9, (name, sub, r["stats"])


def test_tile_scan_cap_and_fallback(emu, cloud_set):
    xyz = cloud_set["plane"]
    ix = emu.index(xyz)
    k = 15
    level = ix.plan(k)["level"]
    ridx, rd2, _, _ = ix.knn(None, k, exact_only=True)
    for cap in (0.75, 1.0, 1.5, 2.0):
        for sub in (2, 4):
            r = ix.tile(k, 0, level, sub=sub, scan_cap=cap, max_points=4096)
            d = r["done"]
            assert np.array_equal(r["idx"][d], ridx[d]) and np.array_equal(r["d2"][d], rd2[d]), cap
    # batched form: the size of the first round only changes the work, never the rows
    for first in (8, 16, 24):
        emu.L.emu_tile_first_cap(first)
        r = ix.tile(k, 0, level, sub=2, alg=2, max_points=4096)
        d = r["done"]
        assert d.mean() > 0.99
        assert np.array_equal(r["idx"][d], ridx[d]) and np.array_equal(r["d2"][d], rd2[d]), first
    emu.L.emu_tile_first_cap(32)
    # a region that does not fit the staging capacity: the whole tile is handed on
    r = ix.tile(k, 0, level, sub=2, max_points=128)
    assert r["stats"]["fallback_tiles"] > 0
    d = r["done"]
    assert np.array_equal(r["idx"][d], ridx[d])
    # thread count of the emulated CTA must not matter
    a = ix.tile(k, 0, level, sub=4, nthreads=160)
    b = ix.tile(k, 0, level, sub=4, nthreads=256)
    assert np.array_equal(a["done"], b["done"]) and np.array_equal(a["idx"], b["idx"])


@pytest.mark.parametrize("name", ["sphere", "plane", "lattice", "dup"])
def test_tile_against_reference_fixtures(emu, name):
    fix = np.load(FIX)
    xyz = fix[name + "_xyz"]
    ix = emu.index(xyz)
    if ix.info()["lfine"] < 2:
        pytest.skip("cloud too small for a tile level")
    for k in (1, 8, 15):
        want = fix["%s_self_k%d_idx" % (name, k)].astype(np.int64)
        wd2 = fix["%s_self_k%d_d2" % (name, k)]
        level = max(2, ix.plan(k)["level"])
        r = ix.tile(k, 0, level, sub=4, max_points=4096)
        d = r["done"] & (want[:, k - 1] >= 0)
        got = r["idx"].astype(np.int64)
        assert np.array_equal(got[d], want[d]), (name, k)
        assert np.array_equal(r["d2"][d], wd2[d]), (name, k)


def test_closed_form_eigenvector(emu):
    """eig3.cuh: smallest_eigenvector_fast (what the tile kernel's normal epilogue runs) against
    numpy.linalg.eigh in float64 on scatter matrices of planar, tilted and isotropic
    neighbourhoods and on matrices with a prescribed small eigengap: 1 - |cos| <= 1e-5 wherever
    the relative gap exceeds 1e-3 (the north star's tolerance is 1e-4)."""
    import ctypes as C

    f32p = C.POINTER(C.c_float)
    emu.L.emu_smallest_eigenvector_fast.argtypes = [f32p, C.c_size_t, f32p]
    rng = np.random.default_rng(0)

    def run(c6):
        c = np.ascontiguousarray(c6, np.float32)
        out = np.zeros((len(c), 3), np.float32)
        emu.L.emu_smallest_eigenvector_fast(c.ctypes.data_as(f32p), len(c), out.ctypes.data_as(f32p))
        return out.astype(np.float64)

    def check(c6, tol):
        c6 = np.asarray(c6, np.float32)
        A = np.zeros((len(c6), 3, 3))
        A[:, 0, 0], A[:, 0, 1], A[:, 0, 2] = c6[:, 0], c6[:, 1], c6[:, 2]
        A[:, 1, 1], A[:, 1, 2], A[:, 2, 2] = c6[:, 3], c6[:, 4], c6[:, 5]
        A[:, 1, 0], A[:, 2, 0], A[:, 2, 1] = A[:, 0, 1], A[:, 0, 2], A[:, 1, 2]
        w, v = np.linalg.eigh(A)
        got = run(c6)
        assert np.all(np.abs(np.linalg.norm(got, axis=1) - 1) < 1e-5)
        well = (w[:, 1] - w[:, 0]) / np.maximum(w[:, 2], 1e-300) > 1e-3
        err = 1 - np.abs((got * v[:, :, 0]).sum(1))
        assert err[well].max() <= tol, err[well].max()

    def scatter(P):
        V = (P - P.mean(0)).astype(np.float32)
        S = V.T @ V
        return [S[0, 0], S[0, 1], S[0, 2], S[1, 1], S[1, 2], S[2, 2]]

    cases = []
    for _ in range(4000):
        flat = np.c_[rng.uniform(-7e-3, 7e-3, (15, 2)), 1e-3 * rng.standard_normal(15)]
        q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        cases += [scatter(flat), scatter(flat @ q + 5.0), scatter(rng.standard_normal((15, 3)))]
    check(cases, 1e-5)
    for gap in (1e-2, 3e-3, 1.05e-3):
        cases = []
        for _ in range(4000):
            q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
            l0 = rng.uniform(0.0, 0.3)
            S = (q * np.array([l0, l0 + gap, 1.0])) @ q.T
            cases.append([S[0, 0], S[0, 1], S[0, 2], S[1, 1], S[1, 2], S[2, 2]])
        check(cases, 1e-5)
    # degenerate inputs give a unit vector, never NaN
    got = run([[0, 0, 0, 0, 0, 0], [1, 0, 0, 1, 0, 1], [1, 0, 0, 0, 0, 0], [2, 0, 0, 2, 0, 0]])
    assert np.all(np.isfinite(got)) and np.allclose(np.linalg.norm(got, axis=1), 1, atol=1e-5)
