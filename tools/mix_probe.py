import importlib, json, os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
pcpx = importlib.import_module("point-cloud-processing_b200")
n = 10_000_000
for cloud, k in (("noise_mix", 15), ("noisy_sphere", 8), ("noisy_plane", 15)):
    xyz = torch.from_numpy(getattr(pcpx.synth, cloud)(n)).cuda()
    idx = torch.empty((n, k), dtype=torch.int32, device="cuda"); cnt = torch.empty((n,), dtype=torch.int32, device="cuda")
    nrm = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    with pcpx.Index(xyz) as ix:
        for what in ("knn", "normals", "mean"):
            ms = []
            for _ in range(3):
                if what == "knn": ix.knn(None, k, out_idx=idx, out_d2=None, out_count=cnt, want_d2=False)
                elif what == "normals": ix.estimate_normals(None, k, out=nrm)
                else: ix.mean_knn_distance(k)
                t = ix.timings(); ms.append(t["kernel_ms"])
            print(cloud, k, what, "ms", round(min(ms[1:]), 3), "deferred", t["deferred_queries"], "expanded", t["expanded_queries"], flush=True)
    del xyz, idx, cnt, nrm
