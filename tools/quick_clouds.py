"""kNN / normals kernel times on the non-planar synthetic clouds (sphere shell, uniform cube,
noise mix).  Usage: python tools/quick_clouds.py [n] [k...]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcpx = importlib.import_module("point-cloud-processing_b200")


def main():
    import torch

    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    ks = [int(a) for a in sys.argv[2:]] or [8, 15]
    for name, gen in (("sphere", pcpx.synth.noisy_sphere), ("cube", pcpx.synth.uniform_cube),
                      ("mix", pcpx.synth.noise_mix), ("scan", pcpx.synth.scan)):
        xyz = gen(n)
        d = torch.from_numpy(xyz).cuda()
        ix = pcpx.Index(d)
        info = ix.info()
        d_nrm = torch.empty((n, 3), dtype=torch.float32, device="cuda")
        for k in ks:
            st = ix.knn_stats(k) / n
            d_idx = torch.empty((n, k), dtype=torch.int32, device="cuda")
            t_n, t_k = [], []
            for _ in range(2):
                ix.estimate_normals(None, k, out=d_nrm)
                t_n.append(ix.timings()["kernel_ms"])
                ix.knn(None, k, out_idx=d_idx, out_d2=None, out_count=None, want_d2=False,
                       want_count=False)
                t_k.append(ix.timings()["kernel_ms"])
            print("%s n=%d k=%d lfine=%d build=%.2fms normals=%.2fms knn=%.2fms cand/q=%.1f "
                  "lookups/q=%.1f attempts/q=%.3f warpmax/q=%.1f"
                  % (name, n, k, info["finest_level"], info["build_ms"], min(t_n), min(t_k),
                     st[0], st[1], st[2], st[3]))
            del d_idx
        ix.close()
        del d, d_nrm
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
