"""Where does the end-to-end step go?  Copy bandwidths alone and together, then whole blocking
C-ABI calls (host in, host out) from 1 / 2 / 3 host threads.  python tools/e2e_probe.py"""
import importlib, os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcpx = importlib.import_module("point-cloud-processing_b200")
import torch

n, K = 10_000_000, 15
xyz = pcpx.synth.noisy_plane(n)
h_in = torch.from_numpy(xyz).pin_memory()
h_out = [torch.empty((n, 3), dtype=torch.float32).pin_memory() for _ in range(6)]
d_a = torch.empty((n, 3), dtype=torch.float32, device="cuda")
d_b = torch.empty((n, 3), dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out[0].copy_(d_b, non_blocking=True)


print("H2D 120 MB alone %.2f ms, D2H alone %.2f ms, both at once %.2f ms" % (
    timed(h2d), timed(d2h), timed(lambda: (h2d(), d2h()))), flush=True)


def once(t):
    ix = pcpx.Index(h_in.numpy())
    ix.estimate_normals(None, K, out=h_out[t].numpy())
    ix.close()


for threads in (1, 2, 3, 4, 6):
    steps = 12

    def work(t):
        torch.cuda.set_device(0)
        for _ in range(steps // threads):
            once(t)

    for rep in range(2):
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        [x.start() for x in th]; [x.join() for x in th]
        dt = (time.perf_counter() - t0) / (steps // threads * threads) * 1e3
    print("%d host thread(s): %.2f ms per cloud" % (threads, dt), flush=True)

# device-resident call + caller-side copies, software-pipelined on three streams
d_in = [torch.empty((n, 3), dtype=torch.float32, device="cuda") for _ in range(2)]
d_out = [torch.empty((n, 3), dtype=torch.float32, device="cuda") for _ in range(2)]
ev_in = [torch.cuda.Event() for _ in range(2)]
ev_out = [torch.cuda.Event() for _ in range(2)]
ev_free = [torch.cuda.Event() for _ in range(2)]


def pipelined(steps):
    with torch.cuda.stream(s1):
        d_in[0].copy_(h_in, non_blocking=True); ev_in[0].record(s1)
    for i in range(steps):
        b = i & 1
        if i + 1 < steps:
            with torch.cuda.stream(s1):
                d_in[1 - b].copy_(h_in, non_blocking=True); ev_in[1 - b].record(s1)
        ev_in[b].synchronize()
        if i >= 2:
            ev_out[b].synchronize()  # the D2H that read d_out[b] two steps ago is done
        ix = pcpx.Index(d_in[b])
        ix.estimate_normals(None, K, out=d_out[b])
        ix.close()
        with torch.cuda.stream(s2):
            h_out[b].copy_(d_out[b], non_blocking=True); ev_out[b].record(s2)
    torch.cuda.synchronize()


pipelined(4)
t0 = time.perf_counter(); pipelined(12); dt = (time.perf_counter() - t0) / 12 * 1e3
print("device-resident calls, copies pipelined by the caller: %.2f ms per cloud" % dt)
