"""One process per GPU: row partition by dof ownership + halo plan.

The reference is shared-memory only; this is the B200-native scale-out of the same
algorithm.  Ownership = contiguous ranges of the reference's DoF numbering (first-touch numbering of
a structured grid fills it plane by plane, so contiguous ranges cut at plane boundaries are z-slabs).
Each rank keeps the cells that touch its dofs, assembles exactly its rows, and before every SpMV
receives the entries of x its rows reference but other ranks own ("ghosts").  torch.distributed is
used only as host plumbing (exchanging the NCCL id and the index lists at setup); the per-iteration
traffic is NCCL send/recv + all-reduce issued by the library on its own stream.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .core import B200CSRMatrix, B200Device, DeviceMesh


def dof_bounds(ndofs: int, nranks: int, plane: int | None = None) -> np.ndarray:
    """Ownership boundaries (nranks+1 entries).  With `plane` (dofs per grid plane) cuts fall on plane
    boundaries so that each halo is exactly one plane."""
    if plane and ndofs % plane == 0 and ndofs // plane >= nranks:
        nplanes = ndofs // plane
        cuts = np.round(np.arange(nranks + 1) * nplanes / nranks).astype(np.int64) * plane
    else:
        cuts = np.round(np.arange(nranks + 1) * ndofs / nranks).astype(np.int64)
    cuts[0], cuts[-1] = 0, ndofs
    return cuts


def ghosts_of(celldofs: np.ndarray, lo: int, hi: int) -> np.ndarray:
    """Sorted global ids referenced by cells touching [lo, hi) but owned elsewhere (host twin of
    tb_mesh_extract_local, used by the CPU tests)."""
    cd = np.asarray(celldofs)
    touch = ((cd >= lo) & (cd < hi)).any(axis=1)
    d = np.unique(cd[touch])
    return d[(d < lo) | (d >= hi)]


@dataclass
class HaloPlan:
    neigh_ranks: np.ndarray   # sorted neighbour ranks
    send_ptr: np.ndarray      # nneigh+1
    send_rows: np.ndarray     # local row ids (global - lo), grouped by neighbour
    recv_ptr: np.ndarray      # nneigh+1, offsets into the ghost block (ghosts are sorted by global id)


def build_halo_plan(rank: int, bounds: np.ndarray, ghost_global: np.ndarray, all_ghosts: list[np.ndarray]) -> HaloPlan:
    """all_ghosts[q] = sorted ghost ids of rank q (all-gathered).  Pure host logic."""
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    owner = np.searchsorted(bounds, ghost_global, side="right") - 1
    recv_from = {int(q): int((owner == q).sum()) for q in np.unique(owner)}
    send_to = {}
    for q, g in enumerate(all_ghosts):
        if q == rank:
            continue
        mine = g[(g >= lo) & (g < hi)]
        if mine.size:
            send_to[q] = mine - lo
    neigh = np.array(sorted(set(recv_from) | set(send_to)), dtype=np.int32)
    send_ptr, recv_ptr, rows = [0], [0], []
    for q in neigh:
        s = send_to.get(int(q), np.empty(0, dtype=np.int64))
        rows.append(s)
        send_ptr.append(send_ptr[-1] + s.size)
        recv_ptr.append(recv_ptr[-1] + recv_from.get(int(q), 0))
    send_rows = np.concatenate(rows) if rows else np.empty(0, dtype=np.int64)
    assert recv_ptr[-1] == ghost_global.size
    return HaloPlan(neigh, np.array(send_ptr, dtype=np.int64), send_rows.astype(np.int64), np.array(recv_ptr, dtype=np.int64))


def init_comm(dev: B200Device, dist) -> None:
    """Create the library's NCCL communicator: rank 0 makes the id, torch.distributed ships it."""
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [B200Device.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    dev.comm_init(rank, world, box[0])


@dataclass
class Partition:
    mesh: DeviceMesh          # local cells, owned dofs first then ghosts
    bounds: np.ndarray
    plan: HaloPlan
    rank: int
    nranks: int

    def attach_halo(self, A: B200CSRMatrix):
        """M, K and A share one pattern object, so attaching to one attaches to all."""
        p = self.plan
        A.set_halo(p.neigh_ranks, p.send_ptr, p.send_rows, p.recv_ptr)


def partition_mesh(dev: B200Device, mesh: DeviceMesh, dist, plane: int | None = None) -> Partition:
    """Split `mesh` (the full grid, present on every rank) by dof ownership and build the halo plan."""
    rank, world = dist.get_rank(), dist.get_world_size()
    bounds = dof_bounds(mesh.ndofs, world, plane)
    local = mesh.extract_local(int(bounds[rank]), int(bounds[rank + 1]))
    gathered = [None] * world
    dist.all_gather_object(gathered, local.ghost_global)
    plan = build_halo_plan(rank, bounds, local.ghost_global, gathered)
    return Partition(local, bounds, plan, rank, world)
