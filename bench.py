#!/usr/bin/env python
"""bench.py — the hot path's headline metric on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], the one the metric is quoted on): k = 15 normal estimation
over a 10 M-point synthetic noisy plane (x, y ~ U[0, 10), z = 1e-3 N(0,1), seed 7).  ONE STEP =
the whole path over one cloud: build the spatial index on the device (bbox -> Morton -> radix
sort -> cell tables) + the fused kNN -> PCA normal kernel over all of its points.

  value   normals/s with the cloud already resident in HBM (device pointers in, normals left in
          HBM), whole job over all ranks, bracketed by barrier + synchronize, max over ranks.
  e2e     the same step through the public C ABI with HOST buffers: pinned xyz in (H2D inside
          the timed region), host normals out (D2H inside).
  roofline  the dominant kernel (normals_kernel): algorithmic bytes (SURVEY.md §8d gather model:
          12 + 12 k + 12 = 204 B per normal at k = 15) / its CUDA-event duration on the
          library's launching stream, against the measured HBM peak in MEASURED_PEAKS.json.
  cpu_baseline  the reference's own octree (oracle/_ref, unmodified headers) + the restated
          estimate_normal, on this box's host cores, on a bounded sample (rank 0, N = 1 only).

N > 1 ("weak"): every rank owns one 10 M-point slab of an N x 10 M-point plane plus a 0.05-wide
halo of its neighbours' points; the halo strips are exchanged between neighbouring ranks over
NCCL inside every timed step (point-cloud-processing_b200/sharding.py: exchange_halo) — the one
real exchange of the sharded path; value = N x 10 M normals / max-over-ranks time.

--impl reference: times the reference's CPU implementation of the same step (full-size octree
build + a bounded query sample, extrapolated linearly) on all host threads; rank 0 only.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS = 10_000_000
K = 15
HALO = 0.05
ALGORITHMIC_BYTES_PER_NORMAL = 12 + 12 * K + 12  # query + k winners + normal out
METRIC = "normals/sec (estimate_normals k=15, 10M-pt noisy plane; index build + fused kNN->PCA per step)"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes of the dominant kernel from the committed ncu capture, if one was summarised"""
    p = os.path.join(ROOT, "profiles", "normals_kernel_traffic.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region: one long-running
    `nvidia-smi -lms 50` whose lines are collected while the context is open."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None
        self.t = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) >= 9:
                self.samples.append(f)

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(0.3)  # let the first samples arrive before the load starts
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            if self.t is not None:
                self.t.join(timeout=5)

    def summary(self):
        def num(x):
            try:
                return float(x)
            except ValueError:
                return None

        sm = [num(s[1]) for s in self.samples if num(s[1]) is not None]
        mx = [num(s[2]) for s in self.samples if num(s[2]) is not None]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for name, v in zip(names, s[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def bind_to_gpu_numa_node(local_rank):
    """Best effort: run this rank (and first-touch its pinned buffers) on the CPUs of the NUMA
    node its GPU hangs off, so that N ranks do not all stage through one socket."""
    try:
        q = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id",
                            "--format=csv,noheader"], capture_output=True, text=True, timeout=10)
        bus = q.stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = "0000:" + bus[9:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def own_slab(pcpx, rank):
    """this rank's 10 M-point slab of the world x 10 M plane: x in [rank L, (rank + 1) L)"""
    L = pcpx.synth.plane_extent(N_POINTS)
    pts = pcpx.synth.noisy_plane(N_POINTS, seed=7 + rank, extent=L)
    pts[:, 0] += np.float32(rank * L)
    return pts, L


# ------------------------------------------------------------------------------------------
def cpu_reference_step(xyz, k, sample_size, seed=123):
    """The reference's CPU path for one step: full-size octree build (single thread: the
    insertion cannot be parallelised) + estimate_normals' body on a fixed-seed sample over all
    host threads, extrapolated linearly to the whole cloud."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib

    threads = max(1, os.cpu_count() or 1)
    rng = np.random.default_rng(seed)
    sample = np.sort(rng.choice(len(xyz), size=min(sample_size, len(xyz)), replace=False))
    if oracle_lib.have_ref():
        kind = "reference"
        ref = oracle_lib.RefBridge()
        t0 = time.perf_counter()
        cloud = ref.cloud(xyz, which=0)
        t_build = time.perf_counter() - t0
        t0 = time.perf_counter()
        cloud.normals_sample(0, sample.astype(np.uint32), k, nthreads=threads)
        t_query = time.perf_counter() - t0
    else:
        kind = "port"
        orc = oracle_lib.Oracle()
        t0 = time.perf_counter()
        cloud = orc.cloud(xyz)
        t_build = time.perf_counter() - t0
        t0 = time.perf_counter()
        cloud.normals(xyz[sample], k, nthreads=threads)
        t_query = time.perf_counter() - t0
    t_full = t_build + t_query * (len(xyz) / len(sample))
    desc = ("octree build over all %d points (1 thread, %.2f s) + estimate_normals body on a "
            "fixed-seed sample of %d points (%d threads, %.2f s), extrapolated linearly"
            % (len(xyz), t_build, len(sample), threads, t_query))
    return {"value": len(xyz) / t_full, "unit": "normals/s", "cores": threads, "kind": kind,
            "sample": desc, "build_s": t_build, "query_sample_s": t_query}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    pcpx = importlib.import_module("point-cloud-processing_b200")
    xyz = pcpx.synth.noisy_plane(N_POINTS)
    times, last = [], None
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        last = cpu_reference_step(xyz, K, args.ref_sample)
        if it >= args.warmup:
            times.append(N_POINTS / last["value"])
    t = float(np.mean(times))
    value = N_POINTS / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "normals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "estimate_normals k=15, 10M-point noisy plane (seed 7, L=10), "
                               "reference CPU path: octree build + kNN + PCA normal",
                   "n_points": N_POINTS, "k": K},
        "cpu_baseline": {k_: last[k_] for k_ in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": value, "unit": "normals/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line["cpu_baseline"]["value"] = value
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local_rank))
    pcpx = importlib.import_module("point-cloud-processing_b200")
    pcpx.lib()  # fail loudly now if the CUDA library is missing

    xyz, L = own_slab(pcpx, rank)
    n_owned = len(xyz)
    h_xyz = torch.from_numpy(xyz).pin_memory()
    # room behind the owned slab for the neighbours' strips: they are received in place
    slack = 0 if world == 1 else 1_000_000
    d_buf = torch.empty((n_owned + slack, 3), dtype=torch.float32, device="cuda")
    d_stage_buf = torch.empty((n_owned + slack, 3), dtype=torch.float32, device="cuda")
    d_own, d_stage = d_buf[:n_owned], d_stage_buf[:n_owned]
    d_own.copy_(h_xyz)
    slab_lo, slab_hi = float(rank * L), float((rank + 1) * L)
    # strips hold about n * HALO / L points; twice that is the exchange capacity
    scratch = None
    if world > 1:
        scratch = pcpx.sharding.HaloScratch(int(2 * n_owned * HALO / L) + 4096, "cuda")

    def local_cloud(d_points, d_buffer):
        """N > 1: the exchange step — boundary strips go to the neighbouring ranks over NCCL"""
        if world == 1:
            return d_points
        return pcpx.sharding.exchange_halo(d_points, 0, slab_lo, slab_hi, HALO, rank, world,
                                           dist, buffer=d_buffer, scratch=scratch)[0]

    n_local = int(local_cloud(d_own, d_buf).shape[0])
    d_nrm = torch.empty((n_local, 3), dtype=torch.float32, device="cuda")
    h_nrm = torch.empty((n_local, 3), dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(stats=None):
        d_xyz = local_cloud(d_own, d_buf)
        torch.cuda.synchronize()  # the library runs on its own stream
        ix = pcpx.Index(d_xyz, device=local_rank)
        tb = ix.timings()
        ix.estimate_normals(None, K, out=d_nrm)
        tq = ix.timings()
        ix.close()
        if stats is not None:
            stats["build_ms"].append(tb["build_ms"])
            stats["sort_ms"].append(tb["sort_ms"])
            stats["kernel_ms"].append(tq["kernel_ms"])
            stats["launches"].append(tb["kernel_launches"] + tq["kernel_launches"])
            stats["retries"].append(tq["retry_queries"])

    def step_e2e():
        if world == 1:
            ix = pcpx.Index(h_xyz.numpy(), device=local_rank)  # host pointer: H2D inside
        else:
            d_stage.copy_(h_xyz, non_blocking=True)  # H2D of the owned slab, then the exchange
            d_xyz = local_cloud(d_stage, d_stage_buf)
            torch.cuda.synchronize()
            ix = pcpx.Index(d_xyz, device=local_rank)
        ix.estimate_normals(None, K, out=h_nrm.numpy()[: ix.n])  # host pointer: D2H inside
        ix.close()

    def timed(fn, steps, warmup, stats=None):
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn() if stats is None else fn(stats)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        barrier()
        return dt

    stats = {k_: [] for k_ in ("build_ms", "sort_ms", "kernel_ms", "launches", "retries")}
    with ClockSampler(local_rank) as clocks:
        dt_res = timed(step_resident, args.steps, args.warmup, stats)
        dt_e2e = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    total_owned = n_owned * world
    value = total_owned * args.steps / dt_res
    e2e_value = total_owned * args.steps / dt_e2e

    kernel_ms = float(np.mean(stats["kernel_ms"]))
    if world > 1:
        t = torch.tensor([kernel_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kernel_ms = float(t.item())
    peak, peak_src = measured_peak_gbs()
    achieved = ALGORITHMIC_BYTES_PER_NORMAL * n_local / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": recorded_traffic(),
                "kernel": "knn_main_kernel<15, MODE_NORMALS>", "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": ALGORITHMIC_BYTES_PER_NORMAL * n_local,
                "peak_source": peak_src,
                "note": "gather-model bytes (12 + 12k + 12 per normal); the kernel is "
                        "instruction-issue bound, DRAM traffic sits far below this figure"}

    # untimed: the halo is wide enough iff every owned point near an inner slab face has its k
    # nearest neighbours inside the local cloud, i.e. at least k + 1 local points (itself
    # included) within its distance to the outer face of the halo
    halo_ok = True
    if world > 1:
        own = xyz
        d_xyz = local_cloud(d_own, d_buf)
        torch.cuda.synchronize()
        lo_gap = own[:, 0] - np.float32(rank * L - HALO) if rank > 0 else None
        hi_gap = np.float32((rank + 1) * L + HALO) - own[:, 0] if rank < world - 1 else None
        gap = np.full(n_owned, np.inf, np.float32)
        if lo_gap is not None:
            gap = np.minimum(gap, lo_gap)
        if hi_gap is not None:
            gap = np.minimum(gap, hi_gap)
        near = np.flatnonzero(gap < 4 * HALO)
        with pcpx.Index(d_xyz, device=local_rank) as ix:
            cnt = ix.radius_count(own[near], 0.0, radii=(gap[near] * np.float32(0.999)))
        halo_ok = bool((cnt >= K + 1).all())
        t = torch.tensor([1 if halo_ok else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        halo_ok = bool(t.item())

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_step(xyz, K, args.ref_sample)
        cpu = {k_: cpu[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "normals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt_res / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": "estimate_normals k=15 over a 10M-point noisy plane per GPU (seed 7+rank, "
                            "L=10, sigma_z=1e-3): device index build + fused kNN->PCA normal kernel",
                "n_points_per_gpu": N_POINTS, "k": K, "halo": HALO if world > 1 else 0.0,
                "local_points": n_local,
                "cache": "inputs larger than L2 (160 MB sorted float4 SoA + cell table vs 126 MB "
                         "L2); the index is rebuilt from scratch every step",
                "parallelism": "one process per GPU, spatial slabs; halo strips exchanged with the neighbouring ranks over NCCL (isend/irecv) inside the timed step",
                "numa_node_rank0": numa_node,
            },
            "e2e": {"value": e2e_value, "unit": "normals/s",
                    "h2d_bytes_per_step": int(n_owned * 12), "d2h_bytes_per_step": int(n_local * 12),
                    "ms_per_step": dt_e2e / args.steps * 1e3},
            "gpu_launches": int(np.sum(stats["launches"])),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks.summary(),
            "breakdown_ms": {"build": float(np.mean(stats["build_ms"])),
                             "sort": float(np.mean(stats["sort_ms"])),
                             "normals_kernel": float(np.mean(stats["kernel_ms"]))},
            "exact_fallback_queries_per_step": float(np.mean(stats["retries"])),
            "halo_sufficient": halo_ok,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-sample", type=int, default=100_000,
                    help="queries in the CPU reference's bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
