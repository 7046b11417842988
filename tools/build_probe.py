"""Index build times (device events) for the synthetic clouds, repeated, to expose allocator or
hash-table anomalies.  Usage: python tools/build_probe.py [n]"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcpx = importlib.import_module("point-cloud-processing_b200")


def main():
    import torch

    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    for name, gen in (("plane", pcpx.synth.noisy_plane), ("mix", pcpx.synth.noise_mix),
                      ("sphere", pcpx.synth.noisy_sphere), ("cube", pcpx.synth.uniform_cube)):
        d = torch.from_numpy(gen(n)).cuda()
        torch.cuda.synchronize()
        out = []
        for _ in range(4):
            t0 = time.perf_counter()
            ix = pcpx.Index(d)
            wall = (time.perf_counter() - t0) * 1e3
            t = ix.timings()
            info = ix.info()
            out.append("build %.2f (sort %.2f) wall %.2f" % (t["build_ms"], t["sort_ms"], wall))
            ix.close()
        print(name, "lfine", info["finest_level"], "cells", info["n_cells"], "bits",
              info["code_bits"], "|", " | ".join(out))
        del d


if __name__ == "__main__":
    main()
