"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: ownership ranges, ghost lists,
halo plan, and that exchanging by the plan reproduces the global SpMV and dot products."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, ret):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle as O
        from thunderbolt_jl_b200 import dist as tbd
        nel = (5, 4, 7)
        m = O.generate_grid(O.HEX8, nel, (0, 0, 0), (1, 1, 1))
        plane = (nel[0] + 1) * (nel[1] + 1)
        bounds = tbd.dof_bounds(m.ndofs, world, plane)
        assert bounds[0] == 0 and bounds[-1] == m.ndofs and np.all(np.diff(bounds) % plane == 0)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        ghosts = tbd.ghosts_of(m.celldofs, lo, hi)
        # with plane-aligned cuts of the first-touch numbering every halo is exactly one plane per side
        nsides = (rank > 0) + (rank < world - 1)
        assert ghosts.size == nsides * plane
        gathered = [None] * world
        dist.all_gather_object(gathered, ghosts)
        plan = tbd.build_halo_plan(rank, bounds, ghosts, gathered)
        assert list(plan.neigh_ranks) == [q for q in (rank - 1, rank + 1) if 0 <= q < world]
        # exchange a global vector's owned parts by the plan
        rng = np.random.default_rng(0)
        xg = rng.standard_normal(m.ndofs)
        x_local = np.concatenate([xg[lo:hi], np.zeros(ghosts.size)])
        reqs, recv_bufs = [], []
        for i, q in enumerate(plan.neigh_ranks):
            s = torch.from_numpy(np.ascontiguousarray(x_local[plan.send_rows[plan.send_ptr[i]:plan.send_ptr[i + 1]]]))
            r = torch.empty(int(plan.recv_ptr[i + 1] - plan.recv_ptr[i]), dtype=torch.float64)
            recv_bufs.append((i, r))
            reqs += [dist.isend(s, int(q)), dist.irecv(r, int(q))]
        for rq in reqs:
            rq.wait()
        for i, r in recv_bufs:
            x_local[(hi - lo) + plan.recv_ptr[i]:(hi - lo) + plan.recv_ptr[i + 1]] = r.numpy()
        assert np.array_equal(x_local[hi - lo:], xg[ghosts])
        # local rows of A (global columns -> local numbering) reproduce the global SpMV bitwise
        rp, ci = m.pattern()
        A = O.axpby_values(O.assemble_mass(m), O.assemble_diffusion(m, 2, O.D_SCALAR, [0.1]), 0.5)
        yg = O.spmv(rp, ci, A, xg)
        g2l = {int(g): (hi - lo) + k for k, g in enumerate(ghosts)}
        y = np.empty(hi - lo)
        for r_ in range(lo, hi):
            v = 0.0
            for k in range(rp[r_], rp[r_ + 1]):
                c = int(ci[k])
                v += A[k] * x_local[c - lo if lo <= c < hi else g2l[c]]
            y[r_ - lo] = v
        assert np.array_equal(y, yg[lo:hi])
        # distributed dot = all-reduce of local dots
        d = torch.tensor([float(xg[lo:hi] @ yg[lo:hi])], dtype=torch.float64)
        dist.all_reduce(d)
        assert abs(d.item() - float(xg @ yg)) <= 1e-12 * abs(float(xg @ yg))
        ret[rank] = "ok"
    except Exception as e:      # pragma: no cover
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partition_and_halo_plan_gloo(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + world + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)


def test_bounds_without_plane():
    sys.path.insert(0, str(ROOT))
    from thunderbolt_jl_b200 import dist as tbd
    b = tbd.dof_bounds(1001, 4)
    assert b[0] == 0 and b[-1] == 1001 and np.all(np.diff(b) > 0)
    b = tbd.dof_bounds(100, 3, plane=7)       # not divisible: falls back to an even split
    assert b[-1] == 100


def test_peer_targets_point_into_the_neighbours_ghost_blocks():
    """Host logic of the NVLink peer path: our boundary entries must land, inside the neighbour's vector, exactly
    where the neighbour's halo plan expects the entries it receives from us (owned block first, then ghosts in
    neighbour order), and the flag slot is our position in the neighbour's list."""
    sys.path.insert(0, str(ROOT))
    import oracle as O
    from thunderbolt_jl_b200 import dist as tbd
    m = O.generate_grid(O.HEX8, (3, 3, 11), (0, 0, 0), (1, 1, 1))
    world = 4
    bounds = tbd.dof_bounds(m.ndofs, world, plane=16)
    ghosts = [tbd.ghosts_of(m.celldofs, int(bounds[r]), int(bounds[r + 1])) for r in range(world)]
    plans = [tbd.build_halo_plan(r, bounds, ghosts[r], ghosts) for r in range(world)]
    nrows = [int(bounds[r + 1] - bounds[r]) for r in range(world)]
    # a global vector; every rank's local image = owned block + ghost block
    x = np.arange(m.ndofs, dtype=np.float64) * 1.5
    local = [np.concatenate([x[bounds[r]:bounds[r + 1]], np.full(ghosts[r].size, np.nan)]) for r in range(world)]
    for r in range(world):
        off, slot = tbd.peer_targets(r, plans[r], plans, nrows)
        for i, q in enumerate(plans[r].neigh_ranks):
            rows = plans[r].send_rows[plans[r].send_ptr[i]:plans[r].send_ptr[i + 1]]
            assert plans[int(q)].neigh_ranks[slot[i]] == r
            local[int(q)][off[i]:off[i] + rows.size] = local[r][rows]          # what k_halo_push stores
    for r in range(world):
        assert np.array_equal(local[r][nrows[r]:], x[ghosts[r]])                # every ghost received its owner's value
