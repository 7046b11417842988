import importlib, os, sys
os.environ["PCPX_TEST_SAME_DEVICE_REPLICAS"] = "1"
sys.path.insert(0, "/root/repo")
import numpy as np, torch
pcpx = importlib.import_module("point-cloud-processing_b200")
synth = pcpx.synth
ndev = torch.cuda.device_count()
devs = [0, 1] if ndev > 1 else [0, 0]
xyz0 = synth.noisy_plane(400_000, seed=5)
with pcpx.Index(xyz0) as warm:
    warm.estimate_normals(None, 15); warm.knn(None, 15)
xyz = synth.noise_mix(300_000, seed=6)
n = len(xyz)
with pcpx.Index(xyz) as one, pcpx.Index(xyz, devices=devs) as many:
    for k in (1, 15):
        out1 = torch.full((n, k), -7, dtype=torch.int32, device="cuda:0")
        out2 = torch.full((n, k), -7, dtype=torch.int32, device="cuda:0")
        d1 = torch.full((n, k), -7.0, dtype=torch.float32, device="cuda:0")
        d2 = torch.full((n, k), -7.0, dtype=torch.float32, device="cuda:0")
        one.knn(None, k, out_idx=out1, out_d2=d1, out_count=None)
        many.knn(None, k, out_idx=out2, out_d2=d2, out_count=None)
        a, b = out1.cpu().numpy(), out2.cpu().numpy()
        print("k", k, "untouched rows one", int((a == -7).all(1).sum()), "many", int((b == -7).all(1).sum()))
        bad = np.flatnonzero((a != b).any(1))
        print("   rows differing", len(bad), bad[:6])
        da, db = d1.cpu().numpy(), d2.cpu().numpy()
        for r in bad[:5]:
            print("   ", r, xyz[r], a[r][:4], b[r][:4], da[r][:4], db[r][:4])
