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


def exchange_halo(own_xyz, axis, lo, hi, halo, rank, world, dist, group=None, buffer=None):
    """The one exchange step of the sharded path: every rank sends the points of its slab that
    lie within `halo` of an inner face to the neighbour across that face and receives the
    neighbour's strip (torch tensors on any device; NCCL moves device tensors over NVLink, gloo
    CPU tensors in the tests).  Returns the local cloud [own ; from the left ; from the right]
    and the number of owned points.  Strip sizes are exchanged first (two 8-byte messages).
    When `own_xyz` is the leading rows of a larger `buffer`, the strips are received straight
    into the rows behind it and the local cloud is a view of `buffer` (no concatenation copy)."""
    import torch

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
