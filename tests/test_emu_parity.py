"""CPU-only check of the search logic the CUDA kernels run: tests/emu builds the SAME
__host__ __device__ traversal code (grid_core / knn_core / radius_core / normals_core / eig3
.cuh) for the host and this file compares it with the oracle and the reference fixtures.
Bit-exact for indices, distances, counts and masks; 1 - |cos| <= 1e-4 for normals."""
import os

import numpy as np
import pytest

from golden import kats

FIX = os.path.join(os.path.dirname(__file__), "golden", "ref_fixtures.npz")
CLOUDS = ["sphere", "cube", "plane", "lattice", "dup"]
PAD = 0xFFFFFFFF


@pytest.fixture(scope="module")
def fix():
    return np.load(FIX)


def as_i64(idx):
    out = idx.astype(np.int64)
    out[idx == PAD] = -1
    return out


def same_knn(a, b):
    """a = emu (idx, d2, cnt, ...), b = oracle (idx, d2, cnt)"""
    return (np.array_equal(as_i64(a[0]), b[0]) and np.array_equal(a[1], b[1])
            and np.array_equal(a[2], b[2]))


def knn_list(ix, q, k):
    idx, d2, cnt, _ = ix.knn(q, k)
    return [list(r[:c]) for r, c in zip(idx, cnt)]


def test_kats(emu):
    box = np.concatenate(kats.UNIT_BOX)
    assert knn_list(emu.index(kats.OCTANT_POINTS, bbox=box), kats.OCTANT_QUERIES, 1) == \
        [[e] for e in kats.OCTANT_EXPECTED]
    assert knn_list(emu.index(kats.SELF_ONLY_POINTS, bbox=box), kats.SELF_ONLY_QUERY, 1) == [[]]
    assert knn_list(emu.index(kats.SELF_PAIR_POINTS, bbox=box), kats.SELF_PAIR_QUERY, 2) == \
        [kats.SELF_PAIR_EXPECTED]
    ix = emu.index(kats.ORDER_POINTS, bbox=box)
    assert knn_list(ix, kats.ORDER_QUERY, 4) == [kats.ORDER_EXPECTED_K4]
    assert knn_list(ix, kats.ORDER_QUERY, 3) == [kats.ORDER_EXPECTED_K3]
    ix = emu.index(kats.RANGE_POINTS, bbox=box)
    for centre, r, expected in kats.RANGE_SPHERES:
        off, idx = ix.radius_search(centre[None], r)
        assert sorted(idx) == expected
    inside, outside = kats.insertion_points()
    assert emu.index(np.concatenate([inside, outside]), bbox=box).info()["n_indexed"] == len(inside)
    for seed in (1, 2):
        cloud, q, k, planted, pbox = kats.planted_corner_case(seed)
        assert set(knn_list(emu.index(cloud, bbox=np.concatenate(pbox)), q, k)[0]) == planted
    pts, k, expected = kats.mean_distance_case()
    _, _, means, _ = emu.index(pts).normals(None, k, want_means=True)
    assert abs(float(np.float32(means.sum(dtype=np.float32) / np.float32(len(pts)))) - float(expected)) < 1e-5


@pytest.mark.parametrize("name", CLOUDS)
def test_against_reference_fixtures(emu, fix, name):
    xyz, q = fix[name + "_xyz"], fix[name + "_queries"]
    ix = emu.index(xyz)
    for k in (1, 8, 15):
        for qq, qn in ((None, "self"), (q, "ext")):
            for exact_only in (False, True):
                idx, d2, cnt, _ = ix.knn(qq, k, exact_only=exact_only)
                want = fix["%s_%s_k%d_idx" % (name, qn, k)]
                assert np.array_equal(as_i64(idx), want.astype(np.int64)), (name, k, qn)
                assert np.array_equal(d2, fix["%s_%s_k%d_d2" % (name, qn, k)])
    for frac in ("0.02", "0.1"):
        r = fix["%s_radius_%s_r" % (name, frac)]
        assert np.array_equal(ix.radius_count(None, r), fix["%s_radius_%s_count" % (name, frac)])
        assert np.array_equal(ix.radius_count(q, r), fix["%s_radius_%s_ext_count" % (name, frac)])
    _, _, means, _ = ix.normals(None, 15, want_means=True)
    assert np.array_equal(means, fix[name + "_mean15"], equal_nan=True)


@pytest.mark.parametrize("k", [1, 4, 8, 10, 15, 16, 21, 30, 32, 33, 64, 100, 256])
def test_knn_all_list_sizes(emu, oracle, k):
    rng = np.random.default_rng(k)
    xyz = rng.uniform(0, 1, (4000, 3)).astype(np.float32) * np.array([1, 1, 0.05], np.float32)
    q = rng.uniform(-0.2, 1.2, (300, 3)).astype(np.float32)
    ix, oc = emu.index(xyz), oracle.cloud(xyz)
    assert same_knn(ix.knn(None, k), oc.knn(None, k))
    assert same_knn(ix.knn(q, k), oc.knn(q, k))


def test_ties_and_duplicates_take_the_exact_path(emu, oracle):
    g = np.stack(np.meshgrid(*[np.arange(12)] * 3, indexing="ij"), -1).reshape(-1, 3)
    lattice = (g * 0.25).astype(np.float32)  # every query has many equidistant neighbours
    ix, oc = emu.index(lattice), oracle.cloud(lattice)
    for k in (6, 7, 15, 26):
        got = ix.knn(None, k)
        assert same_knn(got, oc.knn(None, k))
        assert got[3][3] > 0  # some queries fell back to the exact (d2, index) search
    nrm, ctr, _, ties = ix.normals(None, 7)
    onrm, gap = oc.normals(None, 7)
    assert ties > 0
    # on a lattice most neighbourhoods are isotropic (degenerate eigenvalues): check only the
    # well-conditioned ones, against the same neighbour SET
    well = gap > 1e-3
    if well.any():
        assert (1 - np.abs((nrm * onrm).sum(1)))[well].max() <= 1e-4


def test_near_duplicates_inside_exclusion_box(emu, oracle):
    rng = np.random.default_rng(21)
    base = rng.uniform(0, 1, (1500, 3)).astype(np.float32)
    xyz = np.concatenate([base, base + np.float32(4e-6), base[:300]], 0)
    ix, oc = emu.index(xyz), oracle.cloud(xyz)
    assert same_knn(ix.knn(None, 5), oc.knn(None, 5))
    assert same_knn(ix.knn(None, 5, eps=1e-3), oc.knn(None, 5, eps=1e-3))
    assert same_knn(ix.knn(None, 5, eps=0.0), oc.knn(None, 5, eps=0.0))
    assert np.array_equal(ix.radius_count(None, 1e-5), oc.radius_count(None, 1e-5))


def test_small_clouds_and_k_larger_than_n(emu, oracle):
    for n in (0, 1, 2, 3, 17):
        rng = np.random.default_rng(n)
        xyz = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
        q = rng.uniform(-2, 2, (9, 3)).astype(np.float32)
        ix, oc = emu.index(xyz), oracle.cloud(xyz)
        for k in (1, 5, 32):
            assert same_knn(ix.knn(q, k), oc.knn(q, k)), (n, k)
            if n:
                assert same_knn(ix.knn(None, k), oc.knn(None, k)), (n, k)
        assert np.array_equal(ix.radius_count(q, 0.7), oc.radius_count(q, 0.7))


def test_degenerate_geometry(emu, oracle):
    rng = np.random.default_rng(4)
    line = np.zeros((500, 3), np.float32)
    line[:, 0] = rng.uniform(0, 1, 500)
    same = np.tile(np.array([[0.3, -2.0, 5.0]], np.float32), (40, 1))
    far = (rng.uniform(0, 1, (800, 3)) + 1e4).astype(np.float32)  # coarse fp32 resolution
    for xyz in (line, same, far):
        ix, oc = emu.index(xyz), oracle.cloud(xyz)
        assert same_knn(ix.knn(None, 6), oc.knn(None, 6))
        r = 0.05
        assert np.array_equal(ix.radius_count(None, r), oc.radius_count(None, r, exact_prune=1))


def test_sparse_outliers_need_coarser_levels(emu, oracle):
    """5 % uniform noise around a dense plane: outliers answer from coarser levels."""
    rng = np.random.default_rng(11)
    plane = np.stack([rng.uniform(0, 1, 9500), rng.uniform(0, 1, 9500),
                      1e-3 * rng.standard_normal(9500)], 1)
    noise = rng.uniform([-0.1, -0.1, -0.6], [1.1, 1.1, 0.6], (500, 3))
    xyz = np.concatenate([plane, noise], 0).astype(np.float32)
    rng.shuffle(xyz, axis=0)
    ix, oc = emu.index(xyz), oracle.cloud(xyz)
    got = ix.knn(None, 15)
    assert same_knn(got, oc.knn(None, 15))
    assert got[3][2] > len(xyz)  # more level attempts than queries
    means, radius = oc.mean_knn_distance(15)
    keep = ix.density_keep(float(radius), 5)
    okeep, ocnt, okept = oc.density_filter(float(radius), 5)
    assert np.array_equal(keep, okeep)
    assert 0 < okept < len(xyz)


def test_radius_sweep_incl_large(emu, oracle):
    rng = np.random.default_rng(8)
    xyz = rng.uniform(-1, 1, (3000, 3)).astype(np.float32)
    q = rng.uniform(-3, 3, (200, 3)).astype(np.float32)
    ix, oc = emu.index(xyz), oracle.cloud(xyz)
    for r in (0.0, 1e-4, 0.05, 0.3, 1.0):
        assert np.array_equal(ix.radius_count(None, r), oc.radius_count(None, r))
        assert np.array_equal(ix.radius_count(q, r), oc.radius_count(q, r))
    # r > 1: the library returns the exact set (the reference's own prune would undercount)
    for r in (2.5, 10.0):
        assert np.array_equal(ix.radius_count(q, r), oracle.radius_count_bruteforce(xyz, q, r))
    radii = rng.uniform(0, 0.5, len(q)).astype(np.float32)
    assert np.array_equal(ix.radius_count(q, 0.0, radii=radii),
                          oc.radius_count(q, 0.0, radii=radii))
    off, idx = ix.radius_search(q, 0.4)
    ooff, oidx = oc.radius_search(q, 0.4)
    assert np.array_equal(off, ooff)
    seg = np.repeat(np.arange(len(q)), np.diff(off.astype(np.int64)))
    order = np.lexsort((idx, seg))
    assert np.array_equal(idx[order].astype(np.int64), oidx)


def test_voxel_grid_excludes_points(emu, oracle):
    rng = np.random.default_rng(6)
    xyz = rng.uniform(-2, 2, (5000, 3)).astype(np.float32)
    box = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    ix, oc = emu.index(xyz, bbox=box), oracle.cloud(xyz, bbox=box)
    assert ix.info()["n_indexed"] == oc.size() < len(xyz)
    assert same_knn(ix.knn(None, 8), oc.knn(None, 8))  # un-indexed points are valid queries
    assert np.array_equal(ix.radius_count(None, 0.3), oc.radius_count(None, 0.3))


def test_normals_100k_sphere_sample(emu, oracle):
    rng = np.random.default_rng(42)
    d = rng.standard_normal((30_000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    xyz = (d * (1 + 0.005 * rng.standard_normal((30_000, 1)))).astype(np.float32)
    ix, oc = emu.index(xyz), oracle.cloud(xyz)
    for k in (8, 15, 30):
        nrm, ctr, _, _ = ix.normals(None, k)
        onrm, gap = oc.normals(None, k)
        err = 1 - np.abs((nrm * onrm).sum(1))
        well = gap > 1e-3
        assert well.mean() > 0.99
        assert err[well].max() <= 1e-4
        assert np.allclose(np.linalg.norm(nrm, axis=1), 1, atol=1e-5)


def test_eigensolver_vs_numpy(emu):
    rng = np.random.default_rng(2)
    worst = 0.0
    for _ in range(500):
        basis = np.linalg.qr(rng.standard_normal((3, 3)))[0]
        lam = np.sort(rng.uniform(0, 1, 3)) * np.array([0.05, 1, 1]) * 10 ** rng.uniform(-6, 3)
        A = (basis * lam) @ basis.T
        cov6 = np.array([A[0, 0], A[0, 1], A[0, 2], A[1, 1], A[1, 2], A[2, 2]], np.float32)
        n, gap = emu.smallest_eigenvector(cov6)
        if gap < 1e-2:
            continue
        A32 = np.array([[cov6[0], cov6[1], cov6[2]], [cov6[1], cov6[3], cov6[4]],
                        [cov6[2], cov6[4], cov6[5]]], np.float64)
        w, V = np.linalg.eigh(A32)
        worst = max(worst, 1 - abs(float(V[:, 0] @ n.astype(np.float64))))
    assert worst <= 1e-5
    n, gap = emu.smallest_eigenvector(np.zeros(6, np.float32))
    assert list(n) == [0, 0, 1]
