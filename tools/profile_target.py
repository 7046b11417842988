"""Small, fixed workload for ncu: build a 10 M-point noisy-plane index and run each hot kernel a
few times.  Usage: python tools/profile_target.py [n] [k] [what]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcpx = importlib.import_module("point-cloud-processing_b200")


def main():
    import torch

    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 15
    what = sys.argv[3] if len(sys.argv) > 3 else "normals"
    lf = float(sys.argv[4]) if len(sys.argv) > 4 else None
    if lf is not None:
        pcpx.set_tuning("success_margin", lf)
    for kv in filter(None, os.environ.get("PCPX_TUNING", "").split(",")):  # e.g. tile_sub=2,tile_cap=1.25
        pcpx.set_tuning(kv.split("=")[0], float(kv.split("=")[1]))
    xyz = getattr(pcpx.synth, os.environ.get('PCPX_CLOUD', 'noisy_plane'))(n)
    d_xyz = torch.from_numpy(xyz).cuda()
    torch.cuda.synchronize()
    ix = pcpx.Index(d_xyz, min_cell_occupancy=int(os.environ.get('PCPX_MIN_OCC', '0')),
                    max_level=int(os.environ.get('PCPX_MAX_LEVEL', '0')))
    for _ in range(3):
        ix.close()
        ix = pcpx.Index(d_xyz, min_cell_occupancy=int(os.environ.get('PCPX_MIN_OCC', '0')),
                        max_level=int(os.environ.get('PCPX_MAX_LEVEL', '0')))
    print('build ms', ix.info()['build_ms'], 'code bits', ix.info()['code_bits'])
    print('finest level', ix.info()['finest_level'], 'cells', ix.info()['n_cells'])
    d_nrm = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    d_idx = torch.empty((n, k), dtype=torch.int32, device="cuda")
    d_cnt = torch.empty((n,), dtype=torch.int32, device="cuda")
    for _ in range(3):
        if what in ("normals", "all"):
            ix.estimate_normals(None, k, out=d_nrm)
            print("normals kernel ms", ix.timings()["kernel_ms"])
        if what in ("knn", "all"):
            ix.knn(None, k, out_idx=d_idx, out_d2=None, out_count=d_cnt, want_d2=False)
            print("knn kernel ms", ix.timings()["kernel_ms"])
        if what in ("radius", "all"):
            ix.radius_count(None, 0.01, out_count=d_cnt)
            print("radius kernel ms", ix.timings()["kernel_ms"])
    ix.close()


if __name__ == "__main__":
    main()
