import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pcpx = importlib.import_module("point-cloud-processing_b200")
from oracle_lib import Oracle
for k in (15, 16, 21):
    rng = np.random.default_rng(k)
    xyz = rng.uniform(0, 1, (20_000, 3)).astype(np.float32) * np.array([1, 1, 0.05], np.float32)
    oc = Oracle().cloud(xyz)
    for tile in (0, 1):
        pcpx.set_tuning("tile", tile)
        with pcpx.Index(xyz) as ix:
            idx, d2, cnt = ix.knn(None, k)
            t = ix.timings()
            oi, od2, ocnt = oc.knn(None, k)
            bad = np.flatnonzero((idx.astype(np.int64) != oi).any(1) | (d2 != od2).any(1))
            per, _ = ix.mean_knn_distance(k)
            t2 = ix.timings()
            operm = oc.mean_knn_distance(k)[0]
            badm = np.flatnonzero(~((per == operm) | (np.isnan(per) & np.isnan(operm))))
            print("k", k, "tile", tile, "knn bad rows", len(bad), "deferred", t["deferred_queries"], t["expanded_queries"],
                  "| mean bad", len(badm), "deferred", t2["deferred_queries"], t2["expanded_queries"], "info", ix.info()["finest_level"])
            for b in badm[:3]:
                print("   row", b, per[b], operm[b], "d2 row", d2[b][-3:], od2[b][-3:])
for k in (15, 16, 21):
    rng = np.random.default_rng(k)
    xyz = rng.uniform(0, 1, (20_000, 3)).astype(np.float32) * np.array([1, 1, 0.05], np.float32)
    oc = Oracle().cloud(xyz)
    onrm, gap = oc.normals(None, k)
    for tile in (0, 1):
        pcpx.set_tuning("tile", tile)
        with pcpx.Index(xyz) as ix:
            nrm = ix.estimate_normals(None, k)
            err = 1 - np.abs((nrm * onrm).sum(1))
            well = gap > 1e-3
            w = np.flatnonzero(well & ~(err <= 1e-4))
            print("k", k, "tile", tile, "normals max err", err[well].max(), "bad", len(w), "nan", np.isnan(nrm).any())
            for b in w[:4]:
                print("   row", b, nrm[b], onrm[b], "gap", gap[b], "err", err[b])
