"""Host <-> device copy bandwidth of the box with every rank copying at once (run under torchrun):
what bounds the end-to-end leg of bench.py at N > 1.  Each rank moves 120 MB up and 120 MB down
per iteration, on two streams, from / to pinned memory."""
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 10_000_000
    h_in = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    h_out = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    h_in.uniform_(); h_out.zero_()
    d_a = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    d_b = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, down, reps=20):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt * 1e3

    for name, up, down in (("H2D only", 1, 0), ("D2H only", 0, 1), ("both", 1, 1)):
        run(up, down, 3)
        ms = run(up, down)
        if rank == 0:
            gb = 0.12 * (up + down) * world
            print("%d ranks, %s: %.2f ms per 120 MB copy round (slowest rank) -> %.0f GB/s aggregate"
                  % (world, name, ms, gb / (ms * 1e-3)), flush=True)


if __name__ == "__main__":
    main()
