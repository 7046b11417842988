"""kNN kernel time against the index's min_cell_occupancy on volumetric-ish clouds.
Usage: python tools/occupancy_probe.py"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcpx = importlib.import_module("point-cloud-processing_b200")


def main():
    import torch

    for name, gen, n in (("sphere", pcpx.synth.noisy_sphere, 10_000_000),
                         ("sphere", pcpx.synth.noisy_sphere, 2_000_000),
                         ("sphere", pcpx.synth.noisy_sphere, 50_000_000),
                         ("cube", pcpx.synth.uniform_cube, 10_000_000),
                         ("plane", pcpx.synth.noisy_plane, 10_000_000)):
        xyz = gen(n)
        d = torch.from_numpy(xyz).cuda()
        for occ in (4, 2, 1):
            ix = pcpx.Index(d, min_cell_occupancy=occ)
            info = ix.info()
            for k in (8, 15):
                d_idx = torch.empty((n, k), dtype=torch.int32, device="cuda")
                ts = []
                for _ in range(3):
                    ix.knn(None, k, out_idx=d_idx, out_d2=None, out_count=None, want_d2=False,
                           want_count=False)
                    ts.append(ix.timings()["kernel_ms"])
                st = ix.knn_stats(k) / n
                print("%s n=%d occ>=%d lfine=%d k=%d knn=%.2fms build=%.2fms cand/q=%.1f lookups/q=%.1f "
                      "attempts/q=%.3f" % (name, n, occ, info["finest_level"], k, min(ts),
                                           info["build_ms"], st[0], st[1], st[2]), flush=True)
                del d_idx
            ix.close()
        del d
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
