"""Host-side multi-GPU plan: one process per GPU, spatial slabs with a halo, no data-path
collective.

The cloud is cut into `world` slabs along one axis.  Rank r indexes its own slab plus the points
of the neighbouring slabs that lie within `halo` of its faces, and answers the queries it OWNS.
A kNN answer computed this way is the global answer iff the k-th neighbour distance does not
reach past the halo (`halo_is_sufficient`); radius search / the density filter need halo >= r.
Results are disjoint per rank, so assembling them is a concatenation in rank order."""
import numpy as np


def slab_edges(lo, hi, world):
    return np.linspace(float(lo), float(hi), world + 1)


def owner_of(coord, edges):
    """rank owning each coordinate (half-open slabs, the last one closed)"""
    world = len(edges) - 1
    r = np.searchsorted(edges, coord, side="right") - 1
    return np.clip(r, 0, world - 1)


def local_cloud(xyz, axis, edges, rank, halo):
    """Returns (local points, boolean mask of the ones this rank owns, their global indices)."""
    c = xyz[:, axis]
    own = owner_of(c, edges) == rank
    near = (c >= edges[rank] - halo) & (c <= edges[rank + 1] + halo)
    sel = np.flatnonzero(own | near)
    return np.ascontiguousarray(xyz[sel]), own[sel], sel


def halo_is_sufficient(local_xyz, owned, kth_d2, axis, edges, rank, halo):
    """True iff every owned query's k-th neighbour is provably the global one: its distance is
    smaller than the distance to the outer faces of the halo (faces on the global boundary do
    not constrain)."""
    world = len(edges) - 1
    c = local_xyz[owned, axis].astype(np.float64)
    reach = np.sqrt(kth_d2[owned].astype(np.float64))
    lo_gap = np.where(rank == 0, np.inf, c - (edges[rank] - halo))
    hi_gap = np.where(rank == world - 1, np.inf, (edges[rank + 1] + halo) - c)
    return bool(np.all(reach < np.minimum(lo_gap, hi_gap)))


def query_slice(n, rank, world):
    """contiguous slice of n queries for replicated-index query sharding"""
    b = n * rank // world
    e = n * (rank + 1) // world
    return b, e


class HaloScratch:
    """Device buffers of the fast exchange path, allocated once per rank: fixed-capacity strip
    buffers (so that the strip sizes travel in the same message round as the strips) and the
    two size counters."""

    def __init__(self, capacity, device):
        import torch

        self.capacity = int(capacity)
        f = dict(dtype=torch.float32, device=device)
        self.send = {-1: torch.empty((self.capacity, 3), **f), +1: torch.empty((self.capacity, 3), **f)}
        self.recv = {-1: torch.empty((self.capacity, 3), **f), +1: torch.empty((self.capacity, 3), **f)}
        self.counts_out = torch.zeros(2, dtype=torch.int64, device=device)
        self.counts_in = torch.zeros(2, dtype=torch.int64, device=device)


def _exchange_halo_device(own_xyz, axis, lo, hi, halo, rank, world, dist, group, buffer, scratch):
    """CUDA path: strips cut by one ordered compaction kernel of the library
    (pcpx_extract_bands), sizes and strips exchanged in ONE round of point-to-point messages,
    strips copied behind the owned rows of `buffer`."""
    import torch

    from . import capi

    inf = float("inf")
    below = lo + halo if rank > 0 else -inf
    above = hi - halo if rank < world - 1 else inf
    stream = torch.cuda.current_stream().cuda_stream
    capi.extract_bands(own_xyz, axis, below, above, scratch.send[-1], scratch.send[+1],
                       scratch.counts_out, stream=stream)
    ops = []
    for side, peer in ((-1, rank - 1), (+1, rank + 1)):
        if peer < 0 or peer >= world:
            continue
        k = 0 if side < 0 else 1
        ops.append(dist.P2POp(dist.isend, scratch.counts_out[k:k + 1], peer, group))
        ops.append(dist.P2POp(dist.irecv, scratch.counts_in[k:k + 1], peer, group))
        ops.append(dist.P2POp(dist.isend, scratch.send[side], peer, group))
        ops.append(dist.P2POp(dist.irecv, scratch.recv[side], peer, group))
    if rank == 0:
        scratch.counts_in[0] = 0
    if rank == world - 1:
        scratch.counts_in[1] = 0
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    sent = scratch.counts_out.tolist()
    got = scratch.counts_in.tolist()  # one device -> host read for both neighbours
    if max(sent + got) > scratch.capacity:
        raise RuntimeError("halo strip of %d points exceeds the exchange capacity %d"
                           % (max(sent + got), scratch.capacity))
    n_own = own_xyz.shape[0]
    at = n_own
    if buffer.shape[0] < n_own + got[0] + got[1]:
        raise RuntimeError("local-cloud buffer too small for the received strips")
    for k, side in ((0, -1), (1, +1)):
        if got[k]:
            buffer[at:at + got[k]].copy_(scratch.recv[side][:got[k]])
            at += got[k]
    return buffer[:at], n_own


def exchange_halo(own_xyz, axis, lo, hi, halo, rank, world, dist, group=None, buffer=None,
                  scratch=None):
    """The one exchange step of the sharded path: every rank sends the points of its slab that
    lie within `halo` of an inner face to the neighbour across that face and receives the
    neighbour's strip (torch tensors on any device; NCCL moves device tensors over NVLink, gloo
    CPU tensors in the tests).  Returns the local cloud [own ; from the left ; from the right]
    and the number of owned points.

    With CUDA tensors, `own_xyz` the leading rows of `buffer` and a `HaloScratch`, the device
    path above runs.  Otherwise (CPU tensors, the gloo tests): strips by boolean masking, strip
    sizes exchanged first (two 8-byte messages), then the strips; when `own_xyz` is the leading
    rows of a larger `buffer` the strips are received straight into the rows behind it."""
    import torch

    if (scratch is not None and own_xyz.is_cuda and buffer is not None
            and buffer.data_ptr() == own_xyz.data_ptr()):
        return _exchange_halo_device(own_xyz, axis, lo, hi, halo, rank, world, dist, group,
                                     buffer, scratch)

    c = own_xyz[:, axis]
    send = {}
    if rank > 0:
        send[rank - 1] = own_xyz[c < lo + halo].contiguous()
    if rank < world - 1:
        send[rank + 1] = own_xyz[c > hi - halo].contiguous()
    dev = own_xyz.device
    sizes_out = {p: torch.tensor([t.shape[0]], dtype=torch.int64, device=dev) for p, t in send.items()}
    sizes_in = {p: torch.zeros(1, dtype=torch.int64, device=dev) for p in send}
    ops = []
    for p in send:
        ops.append(dist.P2POp(dist.isend, sizes_out[p], p, group))
        ops.append(dist.P2POp(dist.irecv, sizes_in[p], p, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    n_own = own_xyz.shape[0]
    counts = {p: int(sizes_in[p].item()) for p in send}
    in_place = (buffer is not None and buffer.data_ptr() == own_xyz.data_ptr()
                and buffer.shape[0] >= n_own + sum(counts.values()))
    recv, at = {}, n_own
    for p in sorted(send):
        if in_place:
            recv[p] = buffer[at:at + counts[p]]
            at += counts[p]
        else:
            recv[p] = torch.empty((counts[p], 3), dtype=own_xyz.dtype, device=dev)
    ops = []
    for p in send:
        if send[p].shape[0]:
            ops.append(dist.P2POp(dist.isend, send[p], p, group))
        if recv[p].shape[0]:
            ops.append(dist.P2POp(dist.irecv, recv[p], p, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if in_place:
        return buffer[:at], n_own
    parts = [own_xyz] + [recv[p] for p in sorted(recv)]
    return torch.cat(parts, 0), n_own


def sharded_self_queries(own_xyz, lo, hi, rank, world, dist, k, halo, device, axis=0,
                         max_rounds=8, group=None):
    """Slab mode end to end on CUDA tensors: kNN (distances + neighbour coordinates) and PCA
    normals of the points this rank OWNS, equal to what one index over the whole cloud returns.

    Round 0 exchanges `halo`-wide strips, builds the local index and answers every owned point.
    A row is final iff its k-th neighbour distance is smaller than the distance to the outer
    faces of the halo (`halo_is_sufficient`, per row).  While ANY rank has open rows, every rank
    doubles the strip width, exchanges again, rebuilds, and re-answers only its open rows as
    external queries against the wider local cloud (the repair round): a halo that is too thin
    costs a second pass over a few rows instead of an error.

    Returns dict(d2 [n_own, k], nbr_xyz [n_own, k, 3], normals [n_own, 3], rounds, halo)."""
    import torch

    from . import capi

    n_own = own_xyz.shape[0]
    inf = float("inf")
    d2 = torch.empty((n_own, k), dtype=torch.float32, device=own_xyz.device)
    nbr = torch.empty((n_own, k, 3), dtype=torch.float32, device=own_xyz.device)
    nrm = torch.empty((n_own, 3), dtype=torch.float32, device=own_xyz.device)
    open_rows = None
    rounds = 0
    width = float(hi) - float(lo)
    while True:
        cap = n_own + 16
        buf = torch.empty((n_own + 2 * cap, 3), dtype=torch.float32, device=own_xyz.device)
        buf[:n_own].copy_(own_xyz)
        scratch = HaloScratch(cap, own_xyz.device)
        local, _ = exchange_halo(buf[:n_own], axis, lo, hi, halo, rank, world, dist, group=group,
                                 buffer=buf, scratch=scratch)
        torch.cuda.synchronize()
        with capi.Index(local, device=device) as ix:
            if open_rows is None:
                nl = local.shape[0]
                idx = torch.empty((nl, k), dtype=torch.int32, device=own_xyz.device)
                dd = torch.empty((nl, k), dtype=torch.float32, device=own_xyz.device)
                nn = torch.empty((nl, 3), dtype=torch.float32, device=own_xyz.device)
                ix.knn(None, k, out_idx=idx, out_d2=dd, out_count=None)
                ix.estimate_normals(None, k, out=nn)
                rows = slice(0, n_own)
                d2.copy_(dd[rows]), nrm.copy_(nn[rows])
                nbr.copy_(local[idx[rows].long().clamp_(min=0)])
                check = torch.arange(n_own, device=own_xyz.device)
            else:
                q = own_xyz[open_rows].contiguous()
                m = q.shape[0]
                idx = torch.empty((max(m, 1), k), dtype=torch.int32, device=own_xyz.device)
                dd = torch.empty((max(m, 1), k), dtype=torch.float32, device=own_xyz.device)
                nn = torch.empty((max(m, 1), 3), dtype=torch.float32, device=own_xyz.device)
                if m:
                    ix.knn(q, k, out_idx=idx[:m], out_d2=dd[:m], out_count=None)
                    ix.estimate_normals(q, k, out=nn[:m])
                    d2[open_rows] = dd[:m]
                    nrm[open_rows] = nn[:m]
                    nbr[open_rows] = local[idx[:m].long().clamp_(min=0)]
                check = open_rows
        # rows whose k-th neighbour could be beaten by a point beyond the strips
        c = own_xyz[check, axis].double()
        reach = d2[check, k - 1].double().sqrt()
        gap_lo = (c - (lo - halo)) if rank > 0 else torch.full_like(c, inf)
        gap_hi = ((hi + halo) - c) if rank < world - 1 else torch.full_like(c, inf)
        bad = ~(reach < torch.minimum(gap_lo, gap_hi))
        # a strip as wide as the neighbour's whole slab cannot be widened further: the
        # neighbour's neighbour would have to contribute (not needed for slabs wider than the
        # k-neighbourhoods, which is what slab mode is for)
        open_rows = check[bad]
        flag = torch.tensor([int(open_rows.numel())], device=own_xyz.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        rounds += 1
        if int(flag.item()) == 0 or rounds >= max_rounds or halo >= width:
            if int(flag.item()) != 0:
                raise RuntimeError("slab mode: %d rows still open after %d rounds (halo %g)"
                                   % (int(flag.item()), rounds, halo))
            return dict(d2=d2, nbr_xyz=nbr, normals=nrm, rounds=rounds, halo=halo)
        halo = min(2.0 * halo, width)
