import importlib, os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
pcpx = importlib.import_module("point-cloud-processing_b200")
n, k = 10_000_000, 15
xyz = torch.from_numpy(pcpx.synth.noisy_plane(n)).cuda()
nrm = torch.empty((n, 3), dtype=torch.float32, device="cuda")
with pcpx.Index(xyz) as ix:
    ms = []
    for _ in range(8):
        ix.estimate_normals(None, k, out=nrm); ms.append(ix.timings()["kernel_ms"])
    print(os.environ.get("PCPX_LIB", "product"), "normals kernel ms best %.3f median %.3f" % (min(ms[1:]), float(np.median(ms[1:]))))
