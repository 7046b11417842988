"""Ad-hoc GPU measurement (not the bench contract): build + kNN / normals / radius timings and
search-work statistics on a synthetic cloud.  Usage: python tools/gpu_probe.py [n] [k]"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcpx = importlib.import_module("point-cloud-processing_b200")


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 15
    cloud = sys.argv[3] if len(sys.argv) > 3 else "plane"
    import torch

    xyz = {"plane": pcpx.synth.noisy_plane, "sphere": pcpx.synth.noisy_sphere,
           "mix": pcpx.synth.noise_mix, "cube": pcpx.synth.uniform_cube}[cloud](n)
    d_xyz = torch.from_numpy(xyz).cuda()
    torch.cuda.synchronize()
    out = {"n": n, "k": k, "cloud": cloud}
    for occ in (0,):
        t0 = time.time()
        ix = pcpx.Index(d_xyz, min_cell_occupancy=occ)
        out["build_wall_ms"] = (time.time() - t0) * 1e3
        info = ix.info()
        out["info"] = {a: (b.tolist() if hasattr(b, "tolist") else b) for a, b in info.items()}
        out["build_timings"] = ix.timings()
        d_nrm = torch.empty((n, 3), dtype=torch.float32, device="cuda")
        d_idx = torch.empty((n, k), dtype=torch.int32, device="cuda")
        d_cnt = torch.empty((n,), dtype=torch.int32, device="cuda")
        for lf in (1.0, 1.15, 1.5):
            pcpx.set_tuning("success_margin", lf)
            st = ix.knn_stats(k)
            res = {"stats_per_query": (st / n).tolist()}
            for rep in range(3):
                ix.estimate_normals(None, k, out=d_nrm)
                res.setdefault("normals_kernel_ms", []).append(ix.timings()["kernel_ms"])
            for rep in range(3):
                ix.knn(None, k, out_idx=d_idx, out_d2=None, out_count=d_cnt, want_d2=False)
                res.setdefault("knn_kernel_ms", []).append(ix.timings()["kernel_ms"])
            out["level_factor_%g" % lf] = res
        pcpx.set_tuning("success_margin", 1.15)
        for rep in range(2):
            ix.radius_count(None, 0.01, out_count=d_cnt)
            out.setdefault("radius_count_kernel_ms", []).append(ix.timings()["kernel_ms"])
        out["radius_mean_count"] = float(d_cnt.float().mean().item())
        # external (host) queries end to end
        q = xyz[: min(n, 1_000_000)]
        t0 = time.time()
        nrm = ix.estimate_normals(q, k)
        out["external_1M_normals_wall_ms"] = (time.time() - t0) * 1e3
        out["external_timings"] = ix.timings()
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe_%s_%d_k%d.json" % (cloud, n, k)), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
