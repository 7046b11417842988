"""The N > 1 DEVICE path against the single-GPU answer: two NCCL ranks (one per GPU) run slab
mode — device strip cut, NCCL halo exchange, local index, kNN + normals, repair rounds — and the
assembled rows must equal what one index over the whole cloud returns: k-th... all k distances
and neighbour coordinates bit-exactly, normals bit-exactly too (same neighbour sets through the
same kernels).  Skipped below 2 GPUs (run it with `gpurun --gpus 2`)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_ranks_equal_one_gpu(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "mgpu_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    parts = [np.load(os.path.join(tmp_path, "rank%d.npz" % i)) for i in range(2)]
    for name in ("plane", "mix"):
        ref_d2, ref_nbr = parts[0][name + "_ref_d2"], parts[0][name + "_ref_nbr"]
        ref_nrm = parts[0][name + "_ref_normals"]
        seen = np.zeros(len(ref_d2), bool)
        for p in parts:
            rows = p[name + "_rows"]
            assert not seen[rows].any()
            seen[rows] = True
            assert np.array_equal(p[name + "_d2"], ref_d2[rows]), name
            assert np.array_equal(p[name + "_nbr"], ref_nbr[rows]), name
            cos = np.abs((p[name + "_normals"] * ref_nrm[rows]).sum(1))
            assert np.all(1 - cos[np.isfinite(cos)] <= 1e-4), name
        assert seen.all()
    # the noise mix cannot be answered from the first strip: the repair rounds ran
    assert int(parts[0]["mix_rounds"][0]) > 1


def test_sharded_calls_on_one_gpu(pcpx, oracle):
    """The sharded code path of a multi-device handle on a ONE-GPU box: with
    PCPX_TEST_SAME_DEVICE_REPLICAS set a device may be listed several times, so three replicas
    on device 0 answer three shares of every kNN-shaped call — tile ranges, sorted-query ranges,
    their own hand-on queues.  Rows against the oracle and against the plain one-device handle."""
    import importlib

    os.environ["PCPX_TEST_SAME_DEVICE_REPLICAS"] = "1"
    synth = importlib.import_module("point-cloud-processing_b200.synth")
    rng = np.random.default_rng(4)
    xyz = synth.noise_mix(120_000, seed=8)
    q = (xyz[rng.choice(len(xyz), 20_000)] + rng.normal(0, 0.02, (20_000, 3))).astype(np.float32)
    oc = oracle.cloud(xyz)
    with pcpx.Index(xyz) as one, pcpx.Index(xyz, devices=[0, 0, 0]) as many:
        assert many.info()["n_devices"] == 3
        for k in (8, 15):
            for qq in (None, q):
                a, b = one.knn(qq, k), many.knn(qq, k)
                assert all(np.array_equal(x, y) for x, y in zip(a, b)), (k, qq is None)
            oi, od2, ocnt = oc.knn(q, k)
            idx = b[0].astype(np.int64)
            idx[b[0] == 0xFFFFFFFF] = -1
            assert np.array_equal(idx, oi) and np.array_equal(b[1], od2) and np.array_equal(b[2], ocnt)
            assert np.array_equal(one.estimate_normals(None, k), many.estimate_normals(None, k),
                                  equal_nan=True)
            pa, ma = one.mean_knn_distance(k)
            pb, mb = many.mean_knn_distance(k)
            assert np.array_equal(pa, pb, equal_nan=True) and ma == mb
        for tile in (0, 1):  # the block-search path shards by sorted-query range
            pcpx.set_tuning("tile", tile)
            a, b = one.knn(None, 15), many.knn(None, 15)
            assert all(np.array_equal(x, y) for x, y in zip(a, b)), tile
        pcpx.set_tuning("tile", 1)


def test_replicated_index_equals_one_gpu(pcpx):
    """Several GPUs behind ONE C-ABI handle (pcpx_index_params.devices: replicated index, kNN-shaped
    calls sharded by tile range / sorted-query range, rows written into the primary's buffers over
    NVLink): every output equals the one-device answer bit for bit — self queries (tile pass +
    warp-per-query kernel on every device), external queries, normals, tangent planes, mean
    distances; host and device-resident outputs; plane and noise mix."""
    import importlib

    import torch

    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs 2 GPUs")
    synth = importlib.import_module("point-cloud-processing_b200.synth")
    devices = list(range(min(ndev, 4)))
    rng = np.random.default_rng(9)
    for name, xyz in (("plane", synth.noisy_plane(400_000, seed=5)), ("mix", synth.noise_mix(300_000, seed=6))):
        q = (xyz[rng.choice(len(xyz), 50_000)] + rng.normal(0, 0.01, (50_000, 3))).astype(np.float32)
        with pcpx.Index(xyz) as one, pcpx.Index(xyz, devices=devices) as many:
            assert many.info()["n_devices"] == len(devices) and one.info()["n_devices"] == 1
            for k in (1, 15, 30):
                for qq in (None, q):
                    a, b = one.knn(qq, k), many.knn(qq, k)
                    for x, y in zip(a, b):
                        assert np.array_equal(x, y), (name, k, qq is None)
                assert np.array_equal(one.estimate_normals(None, k), many.estimate_normals(None, k),
                                      equal_nan=True), (name, k)
                pa, ma = one.mean_knn_distance(k)
                pb, mb = many.mean_knn_distance(k)
                assert np.array_equal(pa, pb, equal_nan=True) and (ma == mb or (ma != ma and mb != mb))
            ca, na = one.estimate_tangent_planes(None, 15)
            cb, nb = many.estimate_tangent_planes(None, 15)
            assert np.array_equal(ca, cb, equal_nan=True) and np.array_equal(na, nb, equal_nan=True)
            # k beyond the register lists is answered by the primary alone
            a, b = one.knn(q[:2000], 40), many.knn(q[:2000], 40)
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
            # device-resident outputs on the primary
            d_xyz = torch.from_numpy(xyz).cuda(devices[0])
            out = torch.empty((len(xyz), 3), dtype=torch.float32, device="cuda:%d" % devices[0])
            many.estimate_normals(None, 15, out=out)
            assert np.array_equal(out.cpu().numpy(), one.estimate_normals(None, 15), equal_nan=True)
        # the cloud itself in device memory: replicas fetch it by peer copy
        with pcpx.Index(d_xyz, devices=devices) as many, pcpx.Index(xyz) as one:
            assert np.array_equal(many.estimate_normals(None, 15), one.estimate_normals(None, 15),
                                  equal_nan=True)


def test_sharded_calls_edge_cases(pcpx, oracle):
    """Multi-device handles at the edges: an empty cloud, clouds and query sets smaller than the
    number of shares, k beyond the sharded kernels, a one-entry device list, bad device lists."""
    os.environ["PCPX_TEST_SAME_DEVICE_REPLICAS"] = "1"
    rng = np.random.default_rng(1)
    with pcpx.Index(np.zeros((0, 3), np.float32), devices=[0, 0]) as ix:
        assert ix.info()["n_indexed"] == 0 and ix.info()["n_devices"] == 2
        idx, d2, cnt = ix.knn(rng.uniform(0, 1, (5, 3)).astype(np.float32), 3)
        assert np.all(cnt == 0) and np.all(idx == 0xFFFFFFFF)
    tiny = rng.uniform(0, 1, (5, 3)).astype(np.float32)
    q1 = rng.uniform(0, 1, (1, 3)).astype(np.float32)
    with pcpx.Index(tiny, devices=[0, 0, 0, 0]) as many, pcpx.Index(tiny) as one:
        for qq in (None, q1, tiny[:2] + np.float32(0.01)):
            for k in (1, 3, 8, 40):
                a, b = one.knn(qq, k), many.knn(qq, k)
                assert all(np.array_equal(x, y) for x, y in zip(a, b)), k
        assert np.array_equal(one.estimate_normals(None, 4), many.estimate_normals(None, 4), equal_nan=True)
        pa, ma = one.mean_knn_distance(3)
        pb, mb = many.mean_knn_distance(3)
        assert np.array_equal(pa, pb, equal_nan=True) and ma == mb
    mid = rng.uniform(0, 1, (3000, 3)).astype(np.float32)  # below the tile path's minimum size
    oc = oracle.cloud(mid)
    with pcpx.Index(mid, devices=[0]) as single, pcpx.Index(mid, devices=[0, 0, 0]) as many:
        assert single.info()["n_devices"] == 1
        oi, od2, ocnt = oc.knn(None, 15)
        for ix in (single, many):
            idx, d2, cnt = ix.knn(None, 15)
            assert np.array_equal(idx.astype(np.int64), oi) and np.array_equal(d2, od2)
    for bad in ([0, 99], [-2], list(range(9))):
        with pytest.raises(pcpx.PcpxError) as e:
            pcpx.Index(mid, devices=bad)
        assert e.value.code == -1
    del os.environ["PCPX_TEST_SAME_DEVICE_REPLICAS"]
    with pytest.raises(pcpx.PcpxError) as e:  # the same device twice is an error outside tests
        pcpx.Index(mid, devices=[0, 0])
    assert e.value.code in (-1,)
