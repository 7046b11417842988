"""Tile kernel against the per-thread kernel on one cloud: kernel times (CUDA events, through the
C ABI with device-resident buffers), retry counts and result equality.

    python tools/tile_probe.py [n] [k] [cloud] > gpurun_out/tile_probe.json
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcpx = importlib.import_module("point-cloud-processing_b200")


def timed(fn, ix, reps=5):
    ms, retries = [], []
    for _ in range(reps):
        fn()
        t = ix.timings()
        ms.append(t["kernel_ms"])
        retries.append(t["retry_queries"])
    return dict(first_ms=ms[0], best_ms=min(ms[1:]), median_ms=float(np.median(ms[1:])),
                retries=int(retries[-1]))


def main():
    import torch

    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 15
    cloud = sys.argv[3] if len(sys.argv) > 3 else "noisy_plane"
    variants = [dict(tile=0),
                dict(tile=1, tile_sub=0, tile_cap=1.0),
                dict(tile=1, tile_sub=2, tile_cap=1.0),
                dict(tile=1, tile_sub=1, tile_cap=1.0)]
    if os.environ.get("PCPX_PROBE_VARIANTS"):
        variants = [variants[0]] + json.loads(os.environ["PCPX_PROBE_VARIANTS"])
    xyz = getattr(pcpx.synth, cloud)(n)
    d_xyz = torch.from_numpy(xyz).cuda()
    d_nrm = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    d_idx = torch.empty((n, k), dtype=torch.int32, device="cuda")
    d_d2 = torch.empty((n, k), dtype=torch.float32, device="cuda")
    d_cnt = torch.empty((n,), dtype=torch.int32, device="cuda")
    out = dict(n=n, k=k, cloud=cloud, variants=[])
    ref = {}
    for v in variants:
        for name, val in v.items():
            pcpx.set_tuning(name, val)
        ix = pcpx.Index(d_xyz)
        rec = dict(v)
        rec["info"] = {a: ix.info()[a] for a in ("finest_level", "n_cells", "build_ms")}
        rec["normals"] = timed(lambda: ix.estimate_normals(None, k, out=d_nrm), ix)
        nrm = d_nrm.clone()
        rec["knn"] = timed(lambda: ix.knn(None, k, out_idx=d_idx, out_d2=d_d2, out_count=d_cnt), ix)
        idx, d2, cnt = d_idx.clone(), d_d2.clone(), d_cnt.clone()
        mean = ix.mean_knn_distance(k)
        rec["mean_ms"] = ix.timings()["kernel_ms"]
        if not ref:
            ref = dict(nrm=nrm, idx=idx, d2=d2, cnt=cnt, mean=mean)
        else:
            cos = (nrm * ref["nrm"]).sum(1).abs()
            rec["normals_max_1_minus_cos"] = float((1 - cos).max())
            rec["normals_rows_off_1e-4"] = int(((1 - cos) > 1e-4).sum())
            rec["knn_idx_equal"] = bool(torch.equal(idx, ref["idx"]))
            rec["knn_d2_equal"] = bool(torch.equal(d2, ref["d2"]))
            rec["knn_cnt_equal"] = bool(torch.equal(cnt, ref["cnt"]))
            rec["mean_equal"] = bool(np.array_equal(np.asarray(mean[0]), np.asarray(ref["mean"][0])))
        ix.close()
        out["variants"].append(rec)
        print(json.dumps(rec), file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
