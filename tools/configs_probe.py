"""The other BASELINE.json configurations, measured once each (not the bench contract):
configs[1] kNN k=15 + radius r=0.01 on the 10M plane, configs[2] density filter on 10M with 5 %
noise, configs[3] normals k=30 on a 100M-point scan, configs[4] build + kNN k=8 sweep over
1M..200M points.  Kernel times are CUDA events on the library's stream; clouds resident in HBM.
Usage: python tools/configs_probe.py [max_points]"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcpx = importlib.import_module("point-cloud-processing_b200")
PEAK = 6454.0e9


def roof(units, bytes_per_unit, ms):
    ach = units * bytes_per_unit / (ms * 1e-3)
    return {"units_per_s": units / (ms * 1e-3), "achieved_GBs": ach / 1e9, "frac": ach / PEAK}


def best(f, reps=3):
    out = []
    for _ in range(reps):
        out.append(f())
    return min(out)


def main():
    import torch

    max_n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
    res = {}

    # configs[1]
    xyz = pcpx.synth.noisy_plane(10_000_000)
    d = torch.from_numpy(xyz).cuda()
    n = len(xyz)
    ix = pcpx.Index(d)
    d_idx = torch.empty((n, 15), dtype=torch.int32, device="cuda")
    d_cnt = torch.empty(n, dtype=torch.int32, device="cuda")

    def knn():
        ix.knn(None, 15, out_idx=d_idx, out_d2=None, out_count=d_cnt, want_d2=False)
        return ix.timings()["kernel_ms"]

    def rad():
        ix.radius_count(None, 0.01, out_count=d_cnt)
        return ix.timings()["kernel_ms"]

    t_knn, t_rad = best(knn), best(rad)
    mbar = float(d_cnt.float().mean().item())
    res["knn_k15_plane10M"] = dict(kernel_ms=t_knn, build_ms=ix.info()["build_ms"],
                                   **roof(n, 12 + 12 * 15 + 4 * 15, t_knn))
    res["radius_r0.01_plane10M"] = dict(kernel_ms=t_rad, mean_count=mbar,
                                        **roof(n, 12 + 12 * mbar + 4, t_rad))
    ix.close()
    del d_idx

    # configs[2]
    xyz = pcpx.synth.noise_mix(10_000_000, seed=11)
    d = torch.from_numpy(xyz).cuda()
    ix = pcpx.Index(d)
    per, mean = ix.mean_knn_distance(15)
    t_mean = ix.timings()["kernel_ms"]
    radius = float(np.float32(mean))
    d_mask = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_pts = torch.empty((n, 3), dtype=torch.float32, device="cuda")

    def filt():
        ix.density_filter(radius, 5, out_mask=d_mask, out_xyz=d_pts)
        return ix.timings()["kernel_ms"]

    t_f = best(filt)
    kept = int(d_mask.sum().item())
    res["density_filter_mix10M"] = dict(kernel_ms=t_f, mean_knn_distance_ms=t_mean, radius=radius,
                                        kept=kept, points_per_s=n / (t_f * 1e-3))
    ix.close()
    del d_mask, d_pts

    # configs[3]
    if max_n >= 100_000_000:
        xyz = pcpx.synth.scan(100_000_000)
        d = torch.from_numpy(xyz).cuda()
        n3 = len(xyz)
        del xyz
        t0 = time.time()
        ix = pcpx.Index(d)
        info = ix.info()
        d_nrm = torch.empty((n3, 3), dtype=torch.float32, device="cuda")

        def nrm30():
            ix.estimate_normals(None, 30, out=d_nrm)
            return ix.timings()["kernel_ms"]

        t = best(nrm30, 2)
        res["normals_k30_scan100M"] = dict(kernel_ms=t, build_ms=info["build_ms"],
                                           device_bytes=info["device_bytes"],
                                           finest_level=info["finest_level"],
                                           **roof(n3, 12 + 12 * 30 + 12, t))
        ix.close()
        del d, d_nrm
        torch.cuda.empty_cache()

    # configs[4]
    sweep = {}
    for m in (1, 2, 5, 10, 20, 50, 100, 200):
        nn = m * 1_000_000
        if nn > max_n:
            break
        xyz = pcpx.synth.noisy_sphere(nn)
        d = torch.from_numpy(xyz).cuda()
        del xyz
        builds = []
        for _ in range(2):
            ix = pcpx.Index(d)
            builds.append(ix.info()["build_ms"])
            if _ == 0:
                ix.close()
        d_idx = torch.empty((nn, 8), dtype=torch.int32, device="cuda")

        def knn8():
            ix.knn(None, 8, out_idx=d_idx, out_d2=None, out_count=None, want_d2=False,
                   want_count=False)
            return ix.timings()["kernel_ms"]

        t = best(knn8, 2)
        info = ix.info()
        bits = info["code_bits"]
        passes = (bits + 7) // 8
        kb = 4 if bits + 1 <= 32 else 8
        a_build = 12 + 12 + (kb + 4) + kb + 2 * passes * (kb + 4) + (4 + 12 + 16) + kb
        sweep[str(nn)] = dict(build_ms=min(builds), knn_ms=t, code_bits=bits,
                              finest_level=info["finest_level"],
                              build=roof(nn, a_build, min(builds)),
                              knn=roof(nn, 12 + 12 * 8 + 4 * 8, t))
        ix.close()
        del d, d_idx
        torch.cuda.empty_cache()
    res["sweep_sphere_k8"] = sweep
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "configs_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
