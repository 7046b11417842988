"""Replicated-index mode (pcpx_index_params.devices) on one 10 M-point cloud: build and kNN-shaped
call times for 1 .. N devices behind one handle (strong scaling of the one cloud).
    python tools/replica_probe.py [n] > gpurun_out/replica_probe.json"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcpx = importlib.import_module("point-cloud-processing_b200")


def main():
    import torch
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    k = 15
    ndev = torch.cuda.device_count()
    xyz = pcpx.synth.noisy_plane(n)
    d_xyz = torch.from_numpy(xyz).cuda(0)
    d_nrm = torch.empty((n, 3), dtype=torch.float32, device="cuda:0")
    out = []
    ref = None
    for nd in [d for d in (1, 2, 4, 8) if d <= ndev]:
        devices = list(range(nd))
        rec = dict(n=n, k=k, devices=nd)
        t0 = time.perf_counter()
        ix = pcpx.Index(d_xyz, devices=devices)
        rec["first_build_wall_ms"] = (time.perf_counter() - t0) * 1e3
        ix.close()
        t0 = time.perf_counter()
        ix = pcpx.Index(d_xyz, devices=devices)
        torch.cuda.synchronize()
        rec["build_wall_ms"] = (time.perf_counter() - t0) * 1e3
        rec["build_ms"] = ix.info()["build_ms"]
        wall, kern = [], []
        for _ in range(6):
            t0 = time.perf_counter()
            ix.estimate_normals(None, k, out=d_nrm)
            wall.append((time.perf_counter() - t0) * 1e3)
            kern.append(ix.timings()["kernel_ms"])
        rec["normals_wall_ms"] = min(wall[1:])
        rec["normals_kernel_ms_slowest_device"] = min(kern[1:])
        rec["normals_per_s"] = n / (min(wall[1:]) * 1e-3)
        got = d_nrm.cpu().numpy()
        if ref is None:
            ref = got
        rec["equal_to_one_device"] = bool(np.array_equal(got, ref, equal_nan=True))
        # whole step: build + normals
        t0 = time.perf_counter()
        for _ in range(3):
            ix.close()
            ix = pcpx.Index(d_xyz, devices=devices)
            ix.estimate_normals(None, k, out=d_nrm)
        rec["step_wall_ms"] = (time.perf_counter() - t0) * 1e3 / 3
        ix.close()
        out.append(rec)
        print(json.dumps(rec), file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
