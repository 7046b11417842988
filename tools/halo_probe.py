"""Phase times of the multi-GPU halo exchange (run under torchrun, one rank per GPU)."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    pcpx = importlib.import_module("point-cloud-processing_b200")
    n, halo = 10_000_000, 0.05
    L = pcpx.synth.plane_extent(n)
    xyz = pcpx.synth.noisy_plane(n, seed=7 + rank)
    xyz[:, 0] += rank * L
    buf = torch.empty((n + 1_000_000, 3), dtype=torch.float32, device="cuda")
    own = buf[:n]
    own.copy_(torch.from_numpy(xyz))
    scratch = pcpx.sharding.HaloScratch(int(2 * n * halo / L) + 4096, "cuda")
    lo, hi = rank * L, (rank + 1) * L

    def sync():
        torch.cuda.synchronize()

    def timeit(fn, reps=20):
        for _ in range(3):
            fn()
        sync(); dist.barrier(); sync()
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        sync()
        return (time.perf_counter() - t) / reps * 1e3

    t_new = timeit(lambda: pcpx.sharding.exchange_halo(own, 0, lo, hi, halo, rank, world, dist, buffer=buf, scratch=scratch))
    t_old = timeit(lambda: pcpx.sharding.exchange_halo(own, 0, lo, hi, halo, rank, world, dist, buffer=buf))
    counts = torch.zeros(2, dtype=torch.int64, device="cuda")
    t_extract = timeit(lambda: pcpx.extract_bands(own, 0, lo + halo, hi - halo, scratch.send[-1], scratch.send[+1], counts, stream=torch.cuda.current_stream().cuda_stream))
    c = own[:, 0]
    t_mask = timeit(lambda: (own[c < lo + halo].contiguous(), own[c > hi - halo].contiguous()))
    peer = rank ^ 1

    def p2p(t_out, t_in):
        ops = [dist.P2POp(dist.isend, t_out, peer), dist.P2POp(dist.irecv, t_in, peer)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    small_o, small_i = torch.zeros(1, dtype=torch.int64, device="cuda"), torch.zeros(1, dtype=torch.int64, device="cuda")
    t_small = timeit(lambda: p2p(small_o, small_i)) if peer < world else -1
    t_big = timeit(lambda: p2p(scratch.send[+1], scratch.recv[+1])) if peer < world else -1
    t_item = timeit(lambda: counts.tolist())
    print("rank %d: exchange new %.3f ms, old %.3f ms | extract %.3f, masking %.3f, p2p 8B %.3f, "
          "p2p %d rows %.3f, tolist %.3f" % (rank, t_new, t_old, t_extract, t_mask, t_small,
                                             scratch.capacity, t_big, t_item), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
