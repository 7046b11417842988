"""Worker of tests/test_gpu_multi.py: one rank per GPU (torchrun), slab mode through NCCL.
Every rank answers the points it owns (kNN + normals, with the repair round); rank 0 also answers
the WHOLE cloud on its own GPU.  Results go to <outdir>/rank<r>.npz for the test to compare."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    outdir = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", local_rank))
    pcpx = importlib.import_module("point-cloud-processing_b200")
    sh = pcpx.sharding
    k = 15
    out = {}
    clouds = {
        "plane": (pcpx.synth.noisy_plane(400_000, seed=7), 0.05),
        # 5 % uniform noise: outliers' neighbourhoods are far wider than the first strip,
        # so this cloud needs the repair rounds
        "mix": (pcpx.synth.noise_mix(300_000, seed=11), 0.02),
    }
    for name, (xyz, halo) in clouds.items():
        lo_all, hi_all = float(xyz[:, 0].min()), float(xyz[:, 0].max())
        edges = sh.slab_edges(lo_all, hi_all, world)
        owner = sh.owner_of(xyz[:, 0], edges)
        mine = np.flatnonzero(owner == rank)
        own = torch.from_numpy(np.ascontiguousarray(xyz[mine])).cuda()
        r = sh.sharded_self_queries(own, float(edges[rank]), float(edges[rank + 1]), rank, world,
                                    dist, k, halo, local_rank)
        out[name + "_rows"] = mine
        out[name + "_d2"] = r["d2"].cpu().numpy()
        out[name + "_nbr"] = r["nbr_xyz"].cpu().numpy()
        out[name + "_normals"] = r["normals"].cpu().numpy()
        out[name + "_rounds"] = np.array([r["rounds"]])
        if rank == 0:
            with pcpx.Index(xyz, device=local_rank) as ix:
                idx, d2, _ = ix.knn(None, k)
                out[name + "_ref_d2"] = d2
                out[name + "_ref_nbr"] = xyz[idx.astype(np.int64)]
                out[name + "_ref_normals"] = ix.estimate_normals(None, k)
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), **out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
