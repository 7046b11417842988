"""Orientation search time on a 5 M-point thin shell."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcpx = importlib.import_module("point-cloud-processing_b200")
import torch
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5_000_000
xyz = pcpx.synth.noisy_sphere(n, seed=2, sigma=1e-4)
ix = pcpx.Index(torch.from_numpy(xyz).cuda())
nrm = torch.empty((n, 3), dtype=torch.float32, device="cuda")
ix.estimate_normals(None, 15, out=nrm)
d_idx = torch.empty((n, 15), dtype=torch.int32, device="cuda")
ix.knn(None, 15, out_idx=d_idx, out_d2=None, out_count=None, want_d2=False, want_count=False)
print("knn ms", ix.timings()["kernel_ms"])
for _ in range(3):
    w = nrm.clone()
    _, levels, reached = ix.orient_normals(w, 15, want_stats=True)
    t = ix.timings()
    print("orient ms", round(t["kernel_ms"], 2), "levels", levels, "launches", t["kernel_launches"], flush=True)
