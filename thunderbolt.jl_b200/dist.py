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

import os

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


def peer_targets(rank: int, plan: HaloPlan, all_plans: list, all_nrows: list):
    """For each of our neighbours q: where our boundary entries go inside q's vector (q's owned block, then q's
    ghost block in q's neighbour order) and which of q's halo flags is ours.  Pure host logic."""
    dst_off, dst_slot = [], []
    for q in plan.neigh_ranks:
        pq = all_plans[int(q)]
        idx = int(np.flatnonzero(np.asarray(pq.neigh_ranks) == rank)[0])
        dst_off.append(int(all_nrows[int(q)]) + int(pq.recv_ptr[idx]))
        dst_slot.append(idx)
    return np.array(dst_off, dtype=np.int64), np.array(dst_slot, dtype=np.int32)


@dataclass
class Partition:
    mesh: DeviceMesh          # local cells, owned dofs first then ghosts
    bounds: np.ndarray
    plan: HaloPlan
    rank: int
    nranks: int
    dist: object = None
    peer: bool = False        # True once the NVLink peer-memory path is active
    fused: bool = False       # all ranks agreed on the fused (3 launches per iteration) communication path

    def attach_halo(self, A: B200CSRMatrix, p2p: bool | None = None):
        """M, K and A share one pattern object, so attaching to one attaches to all.

        p2p (default: env TB_P2P != "0"): map every rank's CG work vectors and mailbox window into every other rank
        (CUDA IPC) so that the halo of p and the dot-product reductions are plain NVLink stores issued by the
        kernels themselves; NCCL stays for everything else.  Falls back to NCCL if the mapping fails."""
        p = self.plan
        A.set_halo(p.neigh_ranks, p.send_ptr, p.send_rows, p.recv_ptr)
        if p2p is None:
            p2p = os.environ.get("TB_P2P", "1") != "0"
        if not p2p or self.dist is None or self.nranks < 2:
            return
        dev, dist = A.dev, self.dist
        ok = True
        try:
            if not dev.peer_enabled():
                blob = dev.peer_export(A.ncols)
            else:
                blob = None
        except Exception as e:                      # noqa: BLE001 - any failure means "no IPC here"
            print(f"[tb dist] rank {self.rank}: peer export failed ({e}); staying on NCCL", flush=True)
            blob, ok = None, False
        gathered = [None] * self.nranks
        dist.all_gather_object(gathered, (ok, blob, p, int(self.mesh.ndofs_owned)))
        if not all(g[0] for g in gathered):
            return
        if blob is not None:
            try:
                dev.peer_attach(b"".join(g[1] for g in gathered), self.nranks)
            except Exception as e:                  # noqa: BLE001
                print(f"[tb dist] rank {self.rank}: peer attach failed ({e}); staying on NCCL", flush=True)
                ok = False
        flags = [None] * self.nranks
        dist.all_gather_object(flags, ok)
        if not all(flags):
            raise RuntimeError("peer attach succeeded on some ranks only; set TB_P2P=0")
        off, slot = peer_targets(self.rank, p, [g[2] for g in gathered], [g[3] for g in gathered])
        A.set_halo_peer(off, slot)
        self.peer = True
        # fused vs unfused consume different numbers of halo epochs per solve: the choice must be collective
        caps = [None] * self.nranks
        dist.all_gather_object(caps, bool(A.halo_fused_capable()))
        self.fused = all(caps)
        A.set_halo_fused(self.fused)


def partition_structured_grid(dev: B200Device, celltype, nel, left, right, dist, plane: int | None = None) -> Partition:
    """partition_mesh for a structured Quadrilateral / Hexahedron grid WITHOUT the global grid in HBM: every rank generates
    only its own cells from the closed-form first-touch numbering (tb_mesh_generate_grid_local)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    ndofs = int(np.prod([n + 1 for n in nel]))
    bounds = dof_bounds(ndofs, world, plane)
    local = DeviceMesh.generate_grid_local(dev, celltype, nel, left, right, int(bounds[rank]), int(bounds[rank + 1]))
    gathered = [None] * world
    dist.all_gather_object(gathered, local.ghost_global)
    plan = build_halo_plan(rank, bounds, local.ghost_global, gathered)
    return Partition(local, bounds, plan, rank, world, dist)


def partition_mesh(dev: B200Device, mesh: DeviceMesh, dist, plane: int | None = None) -> Partition:
    """Split `mesh` (the full grid, present on every rank) by dof ownership and build the halo plan."""
    rank, world = dist.get_rank(), dist.get_world_size()
    bounds = dof_bounds(mesh.ndofs, world, plane)
    local = mesh.extract_local(int(bounds[rank]), int(bounds[rank + 1]))
    gathered = [None] * world
    dist.all_gather_object(gathered, local.ghost_global)
    plan = build_halo_plan(rank, bounds, local.ghost_global, gathered)
    return Partition(local, bounds, plan, rank, world, dist)


# =====================================================================================================================
# General meshes ("general meshes via a host-side graph partition"): ownership by recursive coordinate
# bisection of the dof coordinates, a global renumbering that makes every rank's dofs one contiguous range (so the halo
# plan, the peer windows and the kernels above are reused unchanged), and a HOST-side cut of the mesh so that no rank ever
# holds the global mesh in HBM.  The renumbering is a symmetric permutation of the reference's operator: iteration counts
# agree within the +-1 rule, vectors map back through `GeneralPartition.gids_old`.
# =====================================================================================================================
def rcb_partition(x: np.ndarray, nparts: int) -> np.ndarray:
    """Recursive coordinate bisection: part id per point, sizes equal up to one, deterministic (stable sorts, ties by
    index) so that every rank computes the same partition from the same coordinates."""
    x = np.asarray(x, dtype=np.float64)
    part = np.empty(x.shape[0], dtype=np.int32)

    def rec(idx, nparts, first):
        if nparts == 1:
            part[idx] = first
            return
        nl = nparts // 2
        ext = x[idx].max(axis=0) - x[idx].min(axis=0)
        ax = int(np.argmax(ext))
        order = np.argsort(x[idx, ax], kind="stable")
        k = (idx.size * nl) // nparts
        rec(idx[order[:k]], nl, first)
        rec(idx[order[k:]], nparts - nl, first + nl)

    rec(np.arange(x.shape[0]), int(nparts), 0)
    return part


def renumber_by_part(part: np.ndarray, nparts: int):
    """(new_of_old, old_of_new, bounds): dofs of part 0 first, then part 1, ... (old order kept inside a part)."""
    old_of_new = np.argsort(part, kind="stable").astype(np.int64)
    new_of_old = np.empty_like(old_of_new)
    new_of_old[old_of_new] = np.arange(part.size, dtype=np.int64)
    bounds = np.concatenate([[0], np.cumsum(np.bincount(part, minlength=nparts))]).astype(np.int64)
    return new_of_old, old_of_new, bounds


def cut_local_mesh(conn: np.ndarray, coords: np.ndarray, celldofs_new: np.ndarray, lo: int, hi: int):
    """Host twin of tb_mesh_extract_local for renumbered dofs: the cells touching [lo, hi) (ascending global cell id, so the
    ordered-gather assembly adds element contributions in the reference's order), local node / dof numbering (owned dofs
    first, ghosts after in ascending global id).  Returns (cells, lconn, lcoords, lcelldofs, ghost_global)."""
    cd = np.asarray(celldofs_new)
    touch = ((cd >= lo) & (cd < hi)).any(axis=1)
    cells = np.flatnonzero(touch)
    cdl = cd[cells]
    out = (cdl < lo) | (cdl >= hi)
    ghosts = np.unique(cdl[out])
    ldofs = np.where(out, (hi - lo) + np.searchsorted(ghosts, cdl), cdl - lo)
    cl = np.asarray(conn)[cells]
    used = np.unique(cl)
    lconn = np.searchsorted(used, cl)
    return cells, lconn.astype(np.int64), np.asarray(coords)[used], ldofs.astype(np.int64), ghosts.astype(np.int64)


@dataclass
class GeneralPartition(Partition):
    cells: np.ndarray = None        # global ids of the local cells (per-cell coefficient data is subset with this)
    gids_old: np.ndarray = None     # local dof -> dof id in the REFERENCE's numbering (owned first, then ghosts)
    new_of_old: np.ndarray = None
    part: np.ndarray = None

    def local_vector(self, v_old: np.ndarray, ncols: int = 1) -> np.ndarray:
        """state-blocked host vector in the reference's numbering -> this rank's local image (owned + ghosts)"""
        n = v_old.size // ncols
        return np.concatenate([v_old[c * n:(c + 1) * n][self.gids_old] for c in range(ncols)])


def host_cut(rank: int, nranks: int, conn, coords, celldofs, ndofs: int, part: np.ndarray | None = None):
    """Pure host logic of partition_host_mesh (CPU-testable): ownership, renumbering and this rank's local cut."""
    conn, celldofs = np.asarray(conn), np.asarray(celldofs)
    if part is None:
        xd = np.empty((ndofs, np.asarray(coords).shape[1]))
        xd[celldofs.ravel()] = np.asarray(coords)[conn.ravel()]          # Lagrange-1: a dof sits on its vertex
        part = rcb_partition(xd, nranks)
    new_of_old, old_of_new, bounds = renumber_by_part(part, nranks)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    cells, lconn, lcoords, ldofs, ghosts = cut_local_mesh(conn, coords, new_of_old[celldofs], lo, hi)
    gids_old = old_of_new[np.concatenate([np.arange(lo, hi), ghosts])]
    return dict(part=part, new_of_old=new_of_old, old_of_new=old_of_new, bounds=bounds, lo=lo, hi=hi, cells=cells, lconn=lconn,
                lcoords=lcoords, ldofs=ldofs, ghosts=ghosts, gids_old=gids_old)


def partition_host_mesh(dev: B200Device, celltype, conn, coords, celldofs, ndofs: int, dist=None, rank: int | None = None,
                        nranks: int | None = None, part: np.ndarray | None = None) -> GeneralPartition:
    """conn / coords / celldofs: the whole mesh in HOST memory with the reference's dof numbering (what Ferrite hands
    over).  `part` (owner rank per dof) defaults to RCB of the dof coordinates.  Only the local cut is uploaded."""
    rank = dist.get_rank() if rank is None else rank
    nranks = dist.get_world_size() if nranks is None else nranks
    c = host_cut(rank, nranks, conn, coords, celldofs, ndofs, part)
    lo, hi, ghosts = c["lo"], c["hi"], c["ghosts"]
    mesh = DeviceMesh.from_host(dev, celltype, c["lconn"], c["lcoords"], c["ldofs"], (hi - lo) + ghosts.size)
    mesh.set_ownership(hi - lo, lo, ghosts)
    if dist is not None and nranks > 1:
        gathered = [None] * nranks
        dist.all_gather_object(gathered, ghosts)
        plan = build_halo_plan(rank, c["bounds"], ghosts, gathered)
    else:
        plan = HaloPlan(np.empty(0, np.int32), np.zeros(1, np.int64), np.empty(0, np.int64), np.zeros(1, np.int64))
    return GeneralPartition(mesh, c["bounds"], plan, rank, nranks, dist, cells=c["cells"], gids_old=c["gids_old"],
                            new_of_old=c["new_of_old"], part=c["part"])
