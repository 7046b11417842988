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
