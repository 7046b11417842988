"""GPU parity proper: the CUDA path through the C ABI against (i) the reference's known answers,
(ii) fixtures produced by the unmodified reference, (iii) the oracle on seeded clouds incl. the
edge cases the domain has (empty, tiny, k > n, duplicates, exact ties, voxel grid, outliers),
and (iv) at BASELINE.json's full sizes through size-independent properties plus a fixed-seed
sample checked against the oracle built over the full cloud.
Bit-exact for indices / distances / counts / masks; 1 - |cos| <= 1e-4 for normals."""
import os

import numpy as np
import pytest

from golden import kats

pytestmark = pytest.mark.gpu

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


def same_knn(got, want):
    return (np.array_equal(as_i64(got[0]), want[0]) and np.array_equal(got[1], want[1])
            and np.array_equal(got[2], want[2]))


def knn_list(ix, q, k):
    idx, d2, cnt = ix.knn(q, k)
    return [list(r[:c]) for r, c in zip(idx, cnt)]


# ---- (i) the reference's known answers -------------------------------------------------------
def test_kats_through_the_abi(pcpx):
    box = kats.UNIT_BOX
    with pcpx.Index(kats.OCTANT_POINTS, voxel_grid=box) as ix:
        assert knn_list(ix, kats.OCTANT_QUERIES, 1) == [[e] for e in kats.OCTANT_EXPECTED]
    with pcpx.Index(kats.SELF_ONLY_POINTS, voxel_grid=box) as ix:
        assert knn_list(ix, kats.SELF_ONLY_QUERY, 1) == [[]]
    with pcpx.Index(kats.SELF_PAIR_POINTS, voxel_grid=box) as ix:
        assert knn_list(ix, kats.SELF_PAIR_QUERY, 2) == [kats.SELF_PAIR_EXPECTED]
    with pcpx.Index(kats.ORDER_POINTS, voxel_grid=box) as ix:
        assert knn_list(ix, kats.ORDER_QUERY, 4) == [kats.ORDER_EXPECTED_K4]
        assert knn_list(ix, kats.ORDER_QUERY, 3) == [kats.ORDER_EXPECTED_K3]
    with pcpx.Index(kats.RANGE_POINTS, voxel_grid=box) as ix:
        for centre, r, expected in kats.RANGE_SPHERES:
            off, idx = ix.radius_search(centre[None], float(r))
            assert sorted(idx) == expected
    inside, outside = kats.insertion_points()
    with pcpx.Index(np.concatenate([inside, outside]), voxel_grid=box) as ix:
        assert ix.info()["n_indexed"] == len(inside)
    for seed in (1, 2, 3):
        cloud, q, k, planted, pbox = kats.planted_corner_case(seed)
        with pcpx.Index(cloud, voxel_grid=pbox) as ix:
            assert set(knn_list(ix, q, k)[0]) == planted
    pts, k, expected = kats.mean_distance_case()
    with pcpx.Index(pts) as ix:
        per, mean = ix.mean_knn_distance(k)
        assert abs(mean - float(expected)) < 1e-5
    nrm = pcpx.normals_from_neighbourhoods(kats.PCA_CROSS, [0, len(kats.PCA_CROSS)])
    assert np.allclose(np.abs(nrm[0]), kats.PCA_EXPECTED, atol=1e-5)
    assert abs(np.linalg.norm(nrm[0]) - 1) < 1e-5


# ---- (ii) fixtures made by the unmodified reference -----------------------------------------
@pytest.mark.parametrize("name", CLOUDS)
def test_reference_fixtures(pcpx, fix, name):
    xyz, q = fix[name + "_xyz"], fix[name + "_queries"]
    with pcpx.Index(xyz) as ix:
        for k in (1, 8, 15):
            for qq, qn in ((None, "self"), (q, "ext")):
                idx, d2, cnt = ix.knn(qq, k)
                assert np.array_equal(as_i64(idx), fix["%s_%s_k%d_idx" % (name, qn, k)].astype(np.int64))
                assert np.array_equal(d2, fix["%s_%s_k%d_d2" % (name, qn, k)])
        for frac in ("0.02", "0.1"):
            r = float(fix["%s_radius_%s_r" % (name, frac)])
            assert np.array_equal(ix.radius_count(None, r), fix["%s_radius_%s_count" % (name, frac)])
            assert np.array_equal(ix.radius_count(q, r), fix["%s_radius_%s_ext_count" % (name, frac)])
        off, idx = ix.radius_search(None, float(fix[name + "_radius_0.02_r"]))
        seg = np.repeat(np.arange(len(xyz)), np.diff(off.astype(np.int64)))
        order = np.lexsort((idx, seg))
        assert np.array_equal(idx[order].astype(np.int32), fix[name + "_radius_0.02_idx"])
        # PCPX_RADIUS_SORTED: the same lists, each ascending by index, straight from the device
        off2, idx2 = ix.radius_search(None, float(fix[name + "_radius_0.02_r"]), sorted=True)
        assert np.array_equal(off2, off)
        assert np.array_equal(idx2.astype(np.int32), fix[name + "_radius_0.02_idx"])
        per, mean = ix.mean_knn_distance(15)
        assert np.array_equal(per, fix[name + "_mean15"], equal_nan=True)


SCANS = ["stanford_bunny", "fandisk", "detergent", "spray"]


@pytest.mark.parametrize("name", SCANS)
def test_reference_real_scans(pcpx, oracle, name):
    """the reference's example scans (examples/data/*.ply), answers produced by the
    UNMODIFIED reference (tests/golden/ref_scans.npz, make_scan_fixtures.py): full-cloud kNN
    k = 15 rows, per-point mean 15-NN distances and sphere-range counts bit-exact — through the
    default flow (tile pass + warp-per-query kernel), the block-search flow and the warp-only
    flow; normals against the oracle's on the same neighbourhoods."""
    fix = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_scans.npz"))
    xyz = fix[name + "_xyz"]
    want_idx, want_d2 = fix[name + "_k15_idx"].astype(np.int64), fix[name + "_k15_d2"]
    with pcpx.Index(xyz) as ix:
        for knob, val in (("tile", 1), ("tile", 0), ("warp_all", 1)):
            pcpx.set_tuning(knob, val)
            try:
                idx, d2, cnt = ix.knn(None, 15)
                per, _ = ix.mean_knn_distance(15)
            finally:
                pcpx.set_tuning("tile", 1), pcpx.set_tuning("warp_all", 0)
            assert np.array_equal(as_i64(idx), want_idx), (name, knob, val)
            assert np.array_equal(d2, want_d2), (name, knob, val)
            assert np.all(cnt == 15)
            assert np.array_equal(per, fix[name + "_mean15"], equal_nan=True), (name, knob, val)
        assert np.array_equal(ix.radius_count(None, float(fix[name + "_radius_r"])),
                              fix[name + "_radius_count"])
        nrm = ix.estimate_normals(None, 15)
        onrm, gap = oracle.cloud(xyz).normals(None, 15)
        err = 1 - np.abs((nrm * onrm).sum(1))
        assert err[gap > 1e-3].max() <= 1e-4
        assert (gap > 1e-3).mean() > 0.9


# ---- (iii) the oracle on seeded clouds and edge cases ---------------------------------------
@pytest.mark.parametrize("k", [1, 4, 8, 10, 15, 16, 21, 30, 32])
def test_knn_every_list_size(pcpx, oracle, k):
    rng = np.random.default_rng(k)
    xyz = rng.uniform(0, 1, (20_000, 3)).astype(np.float32) * np.array([1, 1, 0.05], np.float32)
    q = rng.uniform(-0.2, 1.2, (2_000, 3)).astype(np.float32)
    oc = oracle.cloud(xyz)
    with pcpx.Index(xyz) as ix:
        assert same_knn(ix.knn(None, k), oc.knn(None, k))
        assert same_knn(ix.knn(q, k), oc.knn(q, k))
        nrm = ix.estimate_normals(None, k)
        onrm, gap = oc.normals(None, k)
        err = 1 - np.abs((nrm * onrm).sum(1))
        well = gap > 1e-3
        if k >= 4:
            assert err[well].max() <= 1e-4
        per, _ = ix.mean_knn_distance(k)
        assert np.array_equal(per, oc.mean_knn_distance(k)[0], equal_nan=True)


@pytest.mark.parametrize("k", [33, 64, 100, 256])
def test_knn_beyond_the_register_list(pcpx, oracle, k):
    """32 < k <= 256 takes the heap kernel (csrc/big_k.cuh): exact, same contract."""
    rng = np.random.default_rng(k)
    xyz = rng.uniform(0, 1, (6_000, 3)).astype(np.float32) * np.array([1, 1, 0.05], np.float32)
    xyz[:200] = xyz[200:400]  # duplicates
    q = rng.uniform(-0.2, 1.2, (500, 3)).astype(np.float32)
    oc = oracle.cloud(xyz)
    with pcpx.Index(xyz) as ix:
        assert same_knn(ix.knn(None, k), oc.knn(None, k))
        assert same_knn(ix.knn(q, k), oc.knn(q, k))
        nrm = ix.estimate_normals(None, k)
        onrm, gap = oc.normals(None, k)
        assert (1 - np.abs((nrm * onrm).sum(1)))[gap > 1e-3].max() <= 1e-4
        per, _ = ix.mean_knn_distance(k)
        assert np.array_equal(per, oc.mean_knn_distance(k)[0], equal_nan=True)
    with pcpx.Index(xyz[:50]) as ix:  # k > n
        assert same_knn(ix.knn(None, k), oracle.cloud(xyz[:50]).knn(None, k))


def test_unsupported_k_fails_loudly(pcpx):
    with pcpx.Index(np.zeros((10, 3), np.float32)) as ix:
        with pytest.raises(pcpx.PcpxError) as e:
            ix.knn(None, 257)
        assert e.value.code == -5
        idx, d2, cnt = ix.knn(None, 0)  # k == 0 -> {} (octree/linked_octree_node.hpp:464)
        assert idx.shape == (10, 0)


def test_exact_ties_take_the_exact_path(pcpx, oracle):
    g = np.stack(np.meshgrid(*[np.arange(16)] * 3, indexing="ij"), -1).reshape(-1, 3)
    lattice = (g * 0.25).astype(np.float32)
    oc = oracle.cloud(lattice)
    with pcpx.Index(lattice) as ix:
        for k in (6, 7, 15, 26):
            assert same_knn(ix.knn(None, k), oc.knn(None, k))
            # bit-equal distances were detected: handed on by the tile pass (to the exact 64-bit
            # keys of the warp-per-query kernel) or answered by the block search's exact tie path
            t = ix.timings()
            assert t["deferred_queries"] + t["retry_queries"] > 0
        for knob in ("tile", "warp_retry"):  # the block search and its per-thread retry kernels
            pcpx.set_tuning(knob, 0)
        try:
            for k in (6, 15):
                assert same_knn(ix.knn(None, k), oc.knn(None, k))
                assert ix.timings()["retry_queries"] > 0
        finally:
            pcpx.set_tuning("tile", 1), pcpx.set_tuning("warp_retry", 1)
        assert np.array_equal(ix.radius_count(None, 0.25), oc.radius_count(None, 0.25))
        nrm = ix.estimate_normals(None, 7)
        onrm, gap = oc.normals(None, 7)
        well = gap > 1e-3
        if well.any():
            assert (1 - np.abs((nrm * onrm).sum(1)))[well].max() <= 1e-4


def test_near_duplicates_and_eps(pcpx, oracle):
    rng = np.random.default_rng(21)
    base = rng.uniform(0, 1, (5000, 3)).astype(np.float32)
    xyz = np.concatenate([base, base + np.float32(4e-6), base[:1000]], 0)
    oc = oracle.cloud(xyz)
    with pcpx.Index(xyz) as ix:
        for eps in (1e-5, 1e-3, 0.0):
            assert same_knn(ix.knn(None, 5, eps=eps), oc.knn(None, 5, eps=eps))
        assert np.array_equal(ix.radius_count(None, 1e-5), oc.radius_count(None, 1e-5))


def test_small_clouds(pcpx, oracle):
    for n in (0, 1, 2, 3, 17, 200):
        rng = np.random.default_rng(n)
        xyz = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
        q = rng.uniform(-2, 2, (9, 3)).astype(np.float32)
        oc = oracle.cloud(xyz)
        with pcpx.Index(xyz) as ix:
            assert ix.info()["n_indexed"] == n
            for k in (1, 5, 32):
                assert same_knn(ix.knn(q, k), oc.knn(q, k)), (n, k)
                assert same_knn(ix.knn(None, k), oc.knn(None, k)), (n, k)
            assert np.array_equal(ix.radius_count(q, 0.7), oc.radius_count(q, 0.7))
            mask, pts, kept = ix.density_filter(0.5, 2)
            okeep, _, okept = oc.density_filter(0.5, 2)
            assert kept == okept and np.array_equal(mask, okeep)


def test_degenerate_geometry(pcpx, oracle):
    rng = np.random.default_rng(4)
    line = np.zeros((5000, 3), np.float32)
    line[:, 0] = rng.uniform(0, 1, 5000)
    same = np.tile(np.array([[0.3, -2.0, 5.0]], np.float32), (400, 1))
    far = (rng.uniform(0, 1, (8000, 3)) + 1e4).astype(np.float32)
    for xyz in (line, same, far):
        oc = oracle.cloud(xyz)
        with pcpx.Index(xyz) as ix:
            assert same_knn(ix.knn(None, 6), oc.knn(None, 6))
            assert np.array_equal(ix.radius_count(None, 0.05),
                                  oc.radius_count(None, 0.05, exact_prune=1))


def test_radius_sweep(pcpx, oracle):
    rng = np.random.default_rng(8)
    xyz = rng.uniform(-1, 1, (30_000, 3)).astype(np.float32)
    q = rng.uniform(-3, 3, (2_000, 3)).astype(np.float32)
    oc = oracle.cloud(xyz)
    with pcpx.Index(xyz) as ix:
        for r in (0.0, 1e-4, 0.05, 0.3, 1.0):
            assert np.array_equal(ix.radius_count(None, r), oc.radius_count(None, r))
            assert np.array_equal(ix.radius_count(q, r), oc.radius_count(q, r))
        # r > 1: exact set (the reference's own prune undercounts there)
        assert np.array_equal(ix.radius_count(q[:200], 2.5),
                              oracle.radius_count_bruteforce(xyz, q[:200], 2.5))
        radii = rng.uniform(0, 0.5, len(q)).astype(np.float32)
        assert np.array_equal(ix.radius_count(q, 0.0, radii=radii),
                              oc.radius_count(q, 0.0, radii=radii))


def test_voxel_grid_and_stride(pcpx, oracle):
    rng = np.random.default_rng(6)
    xyz = rng.uniform(-2, 2, (50_000, 3)).astype(np.float32)
    box = (np.array([-1, -1, -1], np.float32), np.array([1, 1, 1], np.float32))
    oc = oracle.cloud(xyz, bbox=np.concatenate(box))
    with pcpx.Index(xyz, voxel_grid=box) as ix:
        assert ix.info()["n_indexed"] == oc.size() < len(xyz)
        assert same_knn(ix.knn(None, 8), oc.knn(None, 8))  # un-indexed points are valid queries
        assert np.array_equal(ix.radius_count(None, 0.3), oc.radius_count(None, 0.3))
    padded = np.zeros((len(xyz), 5), np.float32)  # 20-byte stride (e.g. xyz + 2 attributes)
    padded[:, :3] = xyz
    oc2 = oracle.cloud(xyz)
    with pcpx.Index(padded, stride_bytes=20) as ix:
        ix.n = len(xyz)
        assert same_knn(ix.knn(None, 8), oc2.knn(None, 8))


def test_device_resident_buffers(pcpx, oracle):
    import torch

    xyz = pcpx.synth.noisy_sphere(50_000, seed=1)
    oc = oracle.cloud(xyz)
    d_xyz = torch.from_numpy(xyz).cuda()
    d_idx = torch.empty((len(xyz), 15), dtype=torch.int32, device="cuda")
    d_cnt = torch.empty(len(xyz), dtype=torch.int32, device="cuda")
    d_nrm = torch.empty((len(xyz), 3), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    with pcpx.Index(d_xyz) as ix:
        ix.knn(None, 15, out_idx=d_idx, out_d2=None, out_count=d_cnt, want_d2=False)
        ix.estimate_normals(None, 15, out=d_nrm)
        oi, od, ocnt = oc.knn(None, 15)
        assert np.array_equal(d_idx.cpu().numpy().astype(np.int64), oi)
        assert np.array_equal(d_cnt.cpu().numpy().astype(np.uint32), ocnt)
        onrm, gap = oc.normals(None, 15)
        err = 1 - np.abs((d_nrm.cpu().numpy() * onrm).sum(1))
        assert err[gap > 1e-3].max() <= 1e-4
        pts, nrm = ix.estimate_tangent_planes(None, 15)
        assert np.allclose(nrm, d_nrm.cpu().numpy())
        # tangent-plane point = neighbourhood centroid (estimate_tangent_planes.hpp:82-94)
        cen = xyz[oi[:100]].astype(np.float64).mean(1)
        assert np.allclose(pts[:100], cen, atol=1e-5)


def test_tangent_planes_reference_scenario(pcpx, oracle):
    """test/algorithm/estimate_tangent_planes.cpp:9-76 restated: uniform cloud in [-10, 10]^3, the
    octree over the voxel grid {(-10,-10,-10), (10,10,10)}, k = 5; EVERY plane's point must equal
    the centre of geometry of the point's 5 nearest neighbours and its normal +- estimate_normal
    of them, both within the reference's are_vectors_equal tolerance (1e-5 per component,
    common/vector3d_queries.hpp:48-64) — the normal where the neighbourhood's eigengap makes it
    well defined (five random points can be near-degenerate; the reference compares two runs of
    the same Eigen code, here two different solvers are compared).  A larger, surface-like case
    (k = 15, every row) follows."""
    rng = np.random.default_rng(12)
    xyz = rng.uniform(-10, 10, (1000, 3)).astype(np.float32)
    box = (np.full(3, -10, np.float32), np.full(3, 10, np.float32))
    oc = oracle.cloud(xyz, bbox=box)
    with pcpx.Index(xyz, voxel_grid=box) as ix:
        pts, nrm = ix.estimate_tangent_planes(None, 5)
        oi, _, ocnt = oc.knn(None, 5)
        assert np.all(ocnt == 5)
        centre = xyz[oi].astype(np.float64).mean(1)
        assert np.abs(pts - centre).max() <= 1e-5
        onrm, gap = oc.normals(None, 5)
        assert np.allclose(np.linalg.norm(nrm, axis=1), 1, atol=1e-5)
        well = gap > 1e-2
        diff = np.minimum(np.abs(nrm - onrm).max(1), np.abs(nrm + onrm).max(1))
        assert well.mean() > 0.8 and diff[well].max() <= 1e-4
        cos = np.abs((nrm * onrm).sum(1))
        assert (1 - cos[gap > 1e-3]).max() <= 1e-4
    surf = pcpx.synth.noisy_sphere(150_000, seed=3)
    oc = oracle.cloud(surf)
    with pcpx.Index(surf) as ix:
        pts, nrm = ix.estimate_tangent_planes(None, 15)
        oi, _, _ = oc.knn(None, 15)
        centre = surf[oi].astype(np.float64).mean(1)
        assert np.abs(pts - centre).max() <= 1e-5 * max(1.0, float(np.abs(surf).max()))
        onrm, gap = oc.normals(None, 15)
        cos = np.abs((nrm * onrm).sum(1))
        assert (1 - cos[gap > 1e-3]).max() <= 1e-4
        assert np.array_equal(nrm, ix.estimate_normals(None, 15))


def test_outliers_and_density_filter(pcpx, oracle):
    xyz = pcpx.synth.noise_mix(200_000, seed=11)
    oc = oracle.cloud(xyz)
    with pcpx.Index(xyz) as ix:
        assert same_knn(ix.knn(None, 15), oc.knn(None, 15))
        per, mean = ix.mean_knn_distance(15)
        operp, omean = oc.mean_knn_distance(15)
        assert np.array_equal(per, operp)
        assert abs(mean - float(omean)) < 1e-5 * float(omean) + 1e-7
        radius = float(np.float32(mean))
        mask, pts, kept = ix.density_filter(radius, 5)
        okeep, ocnt, okept = oc.density_filter(radius, 5)
        assert kept == okept and np.array_equal(mask, okeep)
        assert np.array_equal(pts, xyz[okeep.astype(bool)])
        assert 0.9 * len(xyz) < kept < len(xyz)


# ---- (iv) BASELINE sizes ---------------------------------------------------------------------
@pytest.fixture(scope="module")
def big_plane(pcpx):
    return pcpx.synth.noisy_plane(10_000_000)


@pytest.fixture(scope="module")
def big_oracle(oracle, big_plane):
    return oracle.cloud(big_plane)  # the reference's octree insertion over all 10 M points


def host_d2(xyz, idx, rows):
    """the reference's squared_distance re-evaluated on the host in fp32, no FMA"""
    p = xyz[idx]
    t = xyz[rows][:, None, :]
    d = p - t
    return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]


def test_10m_knn_properties_and_sample(pcpx, big_plane, big_oracle):
    n, k = len(big_plane), 15
    with pcpx.Index(big_plane) as ix:
        idx, d2, cnt = ix.knn(None, k)
        assert (cnt == k).all()
        assert (np.diff(d2, axis=1) >= 0).all()  # nearest -> furthest
        rows = np.random.default_rng(0).choice(n, 200_000, replace=False)
        assert np.array_equal(host_d2(big_plane, idx[rows].astype(np.int64), rows), d2[rows])
        assert (idx != np.arange(n, dtype=np.uint32)[:, None]).all()  # self excluded
        sample = np.sort(np.random.default_rng(1).choice(n, 100_000, replace=False))
        oi, od, ocnt = big_oracle.knn(big_plane[sample], k)
        assert np.array_equal(as_i64(idx[sample]), oi)
        assert np.array_equal(d2[sample], od)
        # radius search r = 0.01 on the same cloud (configs[1])
        cnt_r = ix.radius_count(None, 0.01)
        assert np.array_equal(cnt_r[sample], big_oracle.radius_count(big_plane[sample], 0.01))
        assert 25 < cnt_r.mean() < 40


def test_10m_normals_properties_and_sample(pcpx, big_plane, big_oracle):
    n, k = len(big_plane), 15
    with pcpx.Index(big_plane) as ix:
        nrm = ix.estimate_normals(None, k)
        assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-5)
        assert np.median(np.abs(nrm[:, 2])) > 0.95  # a noisy z = 0 plane
        sample = np.sort(np.random.default_rng(2).choice(n, 100_000, replace=False))
        onrm, gap = big_oracle.normals(big_plane[sample], k)
        err = 1 - np.abs((nrm[sample] * onrm).sum(1))
        well = gap > 1e-3
        assert well.mean() > 0.99
        assert err[well].max() <= 1e-4


def test_10m_density_filter(pcpx, oracle):
    xyz = pcpx.synth.noise_mix(10_000_000, seed=11)  # configs[2]: 5 % uniform noise
    with pcpx.Index(xyz) as ix:
        per, mean = ix.mean_knn_distance(15)
        radius = float(np.float32(mean))
        mask, pts, kept = ix.density_filter(radius, 5)
        assert kept == int(mask.sum()) == len(pts)
        assert np.array_equal(pts, xyz[mask.astype(bool)])  # stable compaction
        cnt = ix.radius_count(None, radius)
        assert np.array_equal(mask.astype(bool), cnt >= 5)  # threshold + count agree
        oc = oracle.cloud(xyz)
        sample = np.sort(np.random.default_rng(3).choice(len(xyz), 100_000, replace=False))
        assert np.array_equal(cnt[sample], oc.radius_count(xyz[sample], radius))
        operp = oc.knn(xyz[sample], 15)[1]
        # idempotence-like: filtering the kept set with threshold 1 keeps everything
        with pcpx.Index(pts) as ix2:
            assert ix2.density_filter(radius, 1)[2] == kept


def test_short_code_build_and_its_fallback(pcpx, oracle):
    """Clouds of up to 2^24 points are first indexed with 10 levels (32-bit sort keys); when the
    cell counts say an 11th level would still hold >= 4 points per cell the build repeats with
    the full code length.  Either way the answers are the oracle's."""
    rng = np.random.default_rng(41)
    n = 300_000
    plain = pcpx.synth.noisy_sphere(n, seed=4)
    # a dense clump plus one far point: at level 10 the clump still sits in a handful of cells
    clump = (rng.uniform(0, 0.002, (n, 3))).astype(np.float32)
    clump[0] = (1.0, 1.0, 1.0)
    line = np.zeros((n, 3), np.float32)
    line[:, 0] = rng.uniform(0, 1, n).astype(np.float32)
    line[:, 1:] = (1e-6 * rng.standard_normal((n, 2))).astype(np.float32)
    for name, xyz, want_bits in (("plain", plain, 30), ("clump", clump, 33), ("line", line, 33)):
        ix = pcpx.Index(xyz)
        info = ix.info()
        assert info["code_bits"] == want_bits, (name, info["code_bits"])
        idx, d2, cnt = ix.knn(None, 8)
        oi, od2, oc = oracle.cloud(xyz).knn(None, 8)
        assert np.array_equal(cnt, oc), name
        assert np.array_equal(idx.astype(np.int64), oi), name
        assert np.array_equal(d2, od2), name
        ix.close()


def test_100m_scan_properties_and_brute_force_sample(pcpx):
    """configs[3] / configs[4] size: 100 M-point synthetic scan on one GPU.  No oracle tree at
    this size; instead the size-independent properties of the answer and an exact brute-force
    check (numpy, fp32, the reference's operation order) of sixteen queries against all
    100 M points.  Results stay on the device; only samples come back."""
    import torch

    n, k = 100_000_000, 8
    xyz = pcpx.synth.scan(n)
    d_xyz = torch.from_numpy(xyz).cuda()
    with pcpx.Index(d_xyz) as ix:
        info = ix.info()
        assert info["n_indexed"] == n and info["code_bits"] > 32  # the full-length code path
        d_idx = torch.empty((n, k), dtype=torch.int32, device="cuda")
        d_d2 = torch.empty((n, k), dtype=torch.float32, device="cuda")
        d_cnt = torch.empty((n,), dtype=torch.int32, device="cuda")
        ix.knn(None, k, out_idx=d_idx, out_d2=d_d2, out_count=d_cnt)
        assert bool((d_cnt == k).all())
        assert bool((d_d2[:, 1:] >= d_d2[:, :-1]).all())  # nearest -> furthest
        rows = torch.arange(n, device="cuda", dtype=torch.int32)[:, None]
        assert bool((d_idx != rows).all())  # self excluded
        del rows
        # equal distances are ordered by original index
        tie = d_d2[:, 1:] == d_d2[:, :-1]
        assert bool((d_idx[:, 1:][tie] > d_idx[:, :-1][tie]).all())
        del tie
        # distances are the reference's squared_distance of the returned indices (1 M sample)
        rng = np.random.default_rng(5)
        sample = np.sort(rng.choice(n, 1_000_000, replace=False))
        s_idx = d_idx[torch.from_numpy(sample).cuda()].cpu().numpy().astype(np.int64)
        s_d2 = d_d2[torch.from_numpy(sample).cuda()].cpu().numpy()
        assert np.array_equal(host_d2(xyz, s_idx, sample), s_d2)
        # exact brute force for 16 queries over all 100 M points
        for q in rng.choice(n, 16, replace=False):
            d = xyz - xyz[q]
            dd = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            excluded = (np.abs(d[:, 0]) < np.float32(1e-5)) & (np.abs(d[:, 1]) < np.float32(1e-5)) \
                & (np.abs(d[:, 2]) < np.float32(1e-5))
            dd[excluded] = np.inf
            cand = np.argpartition(dd, k + 8)[: k + 8]
            order = cand[np.lexsort((cand, dd[cand]))][:k]
            assert np.array_equal(d_idx[int(q)].cpu().numpy().astype(np.int64), order), int(q)
            assert np.array_equal(d_d2[int(q)].cpu().numpy(), dd[order])
        del d_idx, d_d2, d_cnt
        # normals k = 30 (configs[3]): unit length, finite, mostly vertical on the height field
        d_nrm = torch.empty((n, 3), dtype=torch.float32, device="cuda")
        ix.estimate_normals(None, 30, out=d_nrm)
        norms = torch.linalg.vector_norm(d_nrm, dim=1)
        assert bool(torch.isfinite(d_nrm).all()) and float((norms - 1).abs().max()) < 1e-5
        assert float(d_nrm[:, 2].abs().median()) > 0.9


def test_concurrent_calls_on_one_index(pcpx, oracle):
    """include/pcpx.h: the query entry points may be called from several host threads on ONE
    index at once (each call borrows its own stream).  Eight threads run different
    calls — self and external kNN at several k (different tile lists built lazily and
    concurrently), normals, radius counts, the density filter, mean distances — three times over;
    every result must equal the serial answer bit for bit."""
    import threading

    xyz = pcpx.synth.noise_mix(150_000, seed=21)
    rng = np.random.default_rng(2)
    q = (xyz[rng.choice(len(xyz), 30_000)] + rng.normal(0, 0.01, (30_000, 3))).astype(np.float32)
    jobs = [lambda: ix.knn(None, 15), lambda: ix.knn(q, 15), lambda: ix.knn(None, 4),
            lambda: ix.knn(q, 30), lambda: (ix.estimate_normals(None, 15),),
            lambda: (ix.radius_count(None, 0.02),), lambda: ix.density_filter(0.02, 5)[:2],
            lambda: (ix.mean_knn_distance(8)[0],)]
    # serial answers on a fresh index, concurrent ones on another (its tile lists are built
    # inside the race)
    with pcpx.Index(xyz) as ix:
        want = [job() for job in jobs]
    with pcpx.Index(xyz) as ix:
        got = [[None] * 3 for _ in jobs]
        errs = []

        def work(j):
            try:
                for rep in range(3):
                    got[j][rep] = jobs[j]()
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        th = [threading.Thread(target=work, args=(j,)) for j in range(len(jobs))]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs
        for j in range(len(jobs)):
            for rep in range(3):
                for a, b in zip(got[j][rep], want[j]):
                    assert np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True), (j, rep)
    oi, od2, _ = oracle.cloud(xyz).knn(q, 15)
    assert np.array_equal(as_i64(want[1][0]), oi) and np.array_equal(want[1][1], od2)
