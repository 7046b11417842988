"""Times the §8f rows on the GPU (10 M-point noisy plane unless n is given) and, beside them,
the reference's own CPU code on a bounded sample (oracle/_ref bridges, when built).
Usage: python tools/next_rows_probe.py [n] > gpurun_out/next_rows.json"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pcpx = importlib.import_module("point-cloud-processing_b200")


def main():
    import torch

    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    out = {"n": n}
    xyz = pcpx.synth.noisy_plane(n)
    d_xyz = torch.from_numpy(xyz).cuda()
    ix = pcpx.Index(d_xyz)
    d_nrm = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    ix.estimate_normals(None, 15, out=d_nrm)
    _, mean = ix.mean_knn_distance(15)
    sigma = float(mean)
    out["sigmaf"] = sigma

    # orientation
    best = 1e9
    for _ in range(2):
        work = d_nrm.clone()
        _, levels, reached = ix.orient_normals(work, 15, want_stats=True)
        best = min(best, ix.timings()["kernel_ms"])
    out["orient_k15"] = {"ms": best, "levels": levels, "reached": reached,
                         "launches": ix.timings()["kernel_launches"]}
    d_nrm = work

    # bilateral filter, K = 1 and K = 3 (support radius 2 sigma)
    for K in (1, 3):
        best = 1e9
        for _ in range(2):
            _, ms = pcpx.bilateral_filter_points(d_xyz, d_nrm, sigma, sigma / 4, K, want_ms=True)
            best = min(best, ms)
        out["bilateral_points_K%d" % K] = {"ms": best, "points_per_s": n * K / best * 1e3}
    best = 1e9
    for _ in range(2):
        _, ms = pcpx.bilateral_filter_normals(d_xyz, d_nrm, sigma, sigma / 4, 1, want_ms=True)
        best = min(best, ms)
    out["bilateral_normals_K1"] = {"ms": best, "points_per_s": n / best * 1e3}
    out["bilateral_mean_neighbours"] = float(np.asarray(ix.radius_count(None, 2 * sigma)).mean())

    # WLOP: I = n / 10, 5 iterations, h = 4 sigma
    best = 1e9
    for _ in range(2):
        _, ms = pcpx.wlop(d_xyz, n // 10, 4 * sigma, iterations=5, seed=1, want_ms=True)
        best = min(best, ms)
    out["wlop_I_n10_k5"] = {"ms": best, "h": 4 * sigma}
    ix.close()

    # reference CPU code on a sample (same density: a sub-square of the plane)
    try:
        from oracle_lib import RefOrient, RefSmoothing, have_ref_orient, have_ref_smoothing
        m = 200_000
        L = pcpx.synth.plane_extent(n) * np.sqrt(m / n)
        sub = xyz[(xyz[:, 0] < L) & (xyz[:, 1] < L)][:m]
        sub_ix = pcpx.Index(sub)
        nrm = sub_ix.estimate_normals(None, 15)
        sub_ix.close()
        if have_ref_smoothing():
            ref = RefSmoothing()
            t = time.time()
            ref.bilateral_filter_points(sub, nrm, sigma, sigma / 4, 1)
            dt = time.time() - t
            out["ref_cpu_bilateral_points_K1"] = {"sample": len(sub), "s": dt,
                                                  "points_per_s": len(sub) / dt, "threads": 1}
            init = np.random.default_rng(1).permutation(len(sub))[: len(sub) // 10].astype(np.uint32)
            t = time.time()
            ref.wlop(sub, init, 0.45, 4 * sigma, 5, True)
            out["ref_cpu_wlop_I_n10_k5"] = {"sample": len(sub), "s": time.time() - t, "threads": 1}
        if have_ref_orient():
            t = time.time()
            RefOrient().propagate_normal_orientations(sub, 15, nrm)
            out["ref_cpu_orient_k15"] = {"sample": len(sub), "s": time.time() - t, "threads": 1,
                                         "note": "includes the reference's kd-tree build and kNN"}
    except Exception as e:  # the bridges are optional
        out["ref_cpu_error"] = repr(e)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
