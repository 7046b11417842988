"""Smoke-sized pass over the hot path for compute-sanitizer (one tool per run):
index build, tile + queue + retry kernels (kNN, normals, mean distance), radius count / search,
density filter, orientation propagation — on clouds small enough for the instrumented run.

    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_target.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcpx = importlib.import_module("point-cloud-processing_b200")


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 40_000
    for cloud, k in (("noisy_plane", 15), ("noise_mix", 8), ("noisy_sphere", 30)):
        xyz = getattr(pcpx.synth, cloud)(n)
        with pcpx.Index(xyz) as ix:
            idx, d2, cnt = ix.knn(None, k)
            t = ix.timings()
            nrm = ix.estimate_normals(None, k)
            per, mean = ix.mean_knn_distance(k)
            q = xyz[:2000] + np.float32(1e-3)
            ix.knn(q, k)
            ix.estimate_normals(q, k)
            c = ix.radius_count(None, float(mean) * 2)
            off, lst = ix.radius_search(q, float(mean) * 2)
            mask, kept_xyz, kept = ix.density_filter(float(mean), 5)
            ix.orient_normals(nrm, min(k, 15))
            print(cloud, "k", k, "deferred", t["deferred_queries"], "expanded", t["expanded_queries"],
                  "kept", kept, "mean count", float(c.mean()))
    print("sanitize_target: done")


if __name__ == "__main__":
    main()
