"""The N > 1 host logic on CPU: two gloo ranks, spatial slabs + halo, no data-path collective.
The oracle stands in for the GPU compute (this test is about ownership, halo sufficiency and
assembly, not about kernels)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib

    from oracle_lib import Oracle

    pcpx = importlib.import_module("point-cloud-processing_b200")
    sh = pcpx.sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        k = 8
        xyz = pcpx.synth.noisy_plane(40_000, seed=7)
        L = pcpx.synth.plane_extent(40_000)
        edges = sh.slab_edges(0.0, L, world)
        halo = 0.05
        local, owned, gidx = sh.local_cloud(xyz, 0, edges, rank, halo)
        # the same local cloud through the exchange step (what bench.py does over NCCL)
        own_mask = sh.owner_of(xyz[:, 0], edges) == rank
        exch, n_own = sh.exchange_halo(torch.from_numpy(xyz[own_mask]), 0, float(edges[rank]),
                                       float(edges[rank + 1]), halo, rank, world, dist)
        # ... and the in-place variant (strips received into the rows behind the owned slab)
        buf = torch.empty((int(own_mask.sum()) + 5000, 3), dtype=torch.float32)
        buf[: int(own_mask.sum())] = torch.from_numpy(xyz[own_mask])
        exch2, _ = sh.exchange_halo(buf[: int(own_mask.sum())], 0, float(edges[rank]),
                                    float(edges[rank + 1]), halo, rank, world, dist, buffer=buf)
        assert exch2.data_ptr() == buf.data_ptr() and torch.equal(exch2, exch)
        exch = exch.numpy()
        same_set = (n_own == int(owned.sum()) and len(exch) == len(local) and np.array_equal(
            exch[np.lexsort(exch.T)], local[np.lexsort(local.T)]))
        o = Oracle()
        idx, d2, cnt = o.cloud(local).knn(None, k, nthreads=2)
        ok = sh.halo_is_sufficient(local, owned, d2[:, k - 1], 0, edges, rank, halo)
        # local neighbour indices -> global indices, owned rows only
        g_rows = gidx[owned]
        g_nbrs = gidx[idx[owned]]
        # every rank contributes a disjoint set of rows; sizes are exchanged, payloads gathered
        n_own = torch.tensor([int(owned.sum())])
        sizes = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
        dist.all_gather(sizes, n_own)
        total = int(sum(int(s) for s in sizes))
        ok_t = torch.tensor([1 if (ok and same_set) else 0])
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        gathered = [None] * world
        dist.all_gather_object(gathered, (g_rows, g_nbrs))
        if rank == 0:
            full = np.full((len(xyz), k), -1, np.int64)
            for rows, nbrs in gathered:
                assert (full[rows] == -1).all()  # disjoint ownership
                full[rows] = nbrs
            ref_idx, _, _ = o.cloud(xyz).knn(None, k, nthreads=2)
            ret["total"] = total
            ret["halo_ok"] = int(ok_t)
            ret["equal"] = bool(np.array_equal(full, ref_idx))
            # a halo that is too thin must be detected, not silently accepted
            thin_local, thin_owned, _ = sh.local_cloud(xyz, 0, edges, rank, 1e-4)
            _, td2, _ = o.cloud(thin_local).knn(None, k, nthreads=2)
            ret["thin_detected"] = not sh.halo_is_sufficient(thin_local, thin_owned,
                                                             td2[:, k - 1], 0, edges, rank, 1e-4)
    finally:
        dist.destroy_process_group()


def test_two_rank_slabs_match_global():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret["total"] == 40_000
    assert ret["halo_ok"] == 1
    assert ret["equal"]
    assert ret["thin_detected"]


def test_query_slices_partition(pcpx):
    for n in (0, 1, 10, 1_000_003):
        for w in (1, 2, 3, 8):
            parts = [pcpx.sharding.query_slice(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
