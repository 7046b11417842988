"""Off-headline clouds through the kNN-shaped calls: kernel time and how many queries the tile
pass handed on, tile path on / off.  python tools/offhead_probe.py > gpurun_out/offhead.json"""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcpx = importlib.import_module("point-cloud-processing_b200")

def main():
    import torch
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    out = []
    cases = [("noisy_plane", 15), ("noisy_sphere", 8), ("noise_mix", 15), ("noisy_sphere", 15), ("uniform_cube", 8), ("scan", 30)]
    for cloud, k in cases:
        xyz = torch.from_numpy(getattr(pcpx.synth, cloud)(n)).cuda()
        idx = torch.empty((n, k), dtype=torch.int32, device="cuda")
        cnt = torch.empty((n,), dtype=torch.int32, device="cuda")
        nrm = torch.empty((n, 3), dtype=torch.float32, device="cuda")
        for tile, extra in ((0, {"warp_retry": 0}), (0, {"warp_retry": 1}), (1, {"warp_retry": 0}), (1, {"warp_retry": 1})):
            pcpx.set_tuning("tile", tile)
            pcpx.set_tuning("tile_min_queries", 24)
            pcpx.set_tuning("tile_margin", 1.15)
            for a, b in extra.items():
                pcpx.set_tuning(a, b)
            with pcpx.Index(xyz) as ix:
                rec = dict(cloud=cloud, k=k, tile=tile, **extra, finest=ix.info()["finest_level"])
                for what in ("knn", "normals"):
                    ms = []
                    for _ in range(3):
                        if what == "knn":
                            ix.knn(None, k, out_idx=idx, out_d2=None, out_count=cnt, want_d2=False)
                        else:
                            ix.estimate_normals(None, k, out=nrm)
                        t = ix.timings()
                        ms.append(t["kernel_ms"])
                    rec[what + "_ms"] = min(ms[1:])
                    rec[what + "_deferred"] = t["deferred_queries"]
                    rec[what + "_expanded"] = t["expanded_queries"]
            out.append(rec)
            print(json.dumps(rec), file=sys.stderr, flush=True)
        del xyz, idx, cnt, nrm
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))

if __name__ == "__main__":
    main()
