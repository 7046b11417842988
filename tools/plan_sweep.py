"""kNN / normals kernel time over clouds and k (A/B of plan choices: PCPX_LIB selects the library)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcpx = importlib.import_module("point-cloud-processing_b200")
import torch
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
for name, gen in (("plane", pcpx.synth.noisy_plane), ("sphere", pcpx.synth.noisy_sphere),
                  ("cube", pcpx.synth.uniform_cube), ("mix", pcpx.synth.noise_mix), ("scan", pcpx.synth.scan)):
    xyz = gen(n)
    d = torch.from_numpy(xyz).cuda()
    ix = pcpx.Index(d)
    nrm = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    out = []
    for k in (4, 8, 12, 15, 20, 24, 28, 30, 32):
        ts = []
        for _ in range(2):
            ix.estimate_normals(None, k, out=nrm)
            ts.append(ix.timings()["kernel_ms"])
        out.append("k%d %.2f" % (k, min(ts)))
    print(name, " ".join(out), flush=True)
    ix.close()
    del d, nrm
    torch.cuda.empty_cache()
