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


# ---- general meshes: RCB ownership + renumbering + host-side cut (dist.host_cut) ---------------------------------------
def test_rcb_partition_is_balanced_and_deterministic():
    sys.path.insert(0, str(ROOT))
    from thunderbolt_jl_b200 import dist as tbd
    x = np.random.default_rng(0).random((1003, 3)) * [4.0, 1.0, 2.0]
    for k in (1, 2, 3, 5, 8):
        p = tbd.rcb_partition(x, k)
        cnt = np.bincount(p, minlength=k)
        assert cnt.sum() == 1003 and cnt.max() - cnt.min() <= k and np.array_equal(p, tbd.rcb_partition(x.copy(), k))
    # two parts of a box elongated in x are cut across x
    p = tbd.rcb_partition(x, 2)
    assert x[p == 0, 0].max() <= x[p == 1, 0].min()
    n2o, o2n, b = tbd.renumber_by_part(p, 2)
    assert np.array_equal(np.sort(n2o), np.arange(1003)) and np.all(p[o2n[:b[1]]] == 0) and np.all(p[o2n[b[1]:]] == 1)


@pytest.mark.parametrize("world", [2, 3, 5])
def test_host_cut_of_an_lv_mesh_reproduces_the_global_operator(world):
    """Every rank's local cut (host logic only): assembling the LOCAL cells with the oracle's element loop gives, on the owned
    rows, exactly the global operator permuted to the new numbering -- bitwise, because the cut keeps the global cell
    order -- and a halo exchange by the plan reproduces the global SpMV."""
    sys.path.insert(0, str(ROOT))
    import oracle as O
    from thunderbolt_jl_b200 import dist as tbd, lv
    nodes, hexes, wedges, prm = lv.generate_ideal_lv_mesh(12, 2, 6)
    tets = lv.tetrahedralize(nodes, hexes, wedges)
    mg = O.Mesh(O.TET4, tets, nodes)
    rp, ci = mg.pattern()
    Ag = O.axpby_values(O.assemble_mass(mg, 2), O.assemble_diffusion(mg, 2, O.D_SCALAR, [0.3]), 0.05)
    xg = np.random.default_rng(1).standard_normal(mg.ndofs)
    yg = O.spmv(rp, ci, Ag, xg)
    cuts = [tbd.host_cut(r, world, mg.conn, mg.coords, mg.celldofs, mg.ndofs) for r in range(world)]
    assert all(np.array_equal(c["part"], cuts[0]["part"]) for c in cuts)
    owned = np.concatenate([c["gids_old"][:c["hi"] - c["lo"]] for c in cuts])
    assert np.array_equal(np.sort(owned), np.arange(mg.ndofs))                      # every dof owned exactly once
    ghosts = [c["ghosts"] for c in cuts]
    for r, c in enumerate(cuts):
        plan = tbd.build_halo_plan(r, c["bounds"], c["ghosts"], ghosts)
        no = c["hi"] - c["lo"]
        # "exchange": what the neighbours would send by THEIR plans
        xl = np.concatenate([xg[c["gids_old"][:no]], np.full(c["ghosts"].size, np.nan)])
        for i, q in enumerate(plan.neigh_ranks):
            pq = tbd.build_halo_plan(int(q), c["bounds"], ghosts[int(q)], ghosts)
            j = int(np.flatnonzero(pq.neigh_ranks == r)[0])
            rows = pq.send_rows[pq.send_ptr[j]:pq.send_ptr[j + 1]]
            cq = cuts[int(q)]
            xl[no + plan.recv_ptr[i]:no + plan.recv_ptr[i + 1]] = xg[cq["gids_old"][:cq["hi"] - cq["lo"]]][rows]
        assert np.array_equal(xl, xg[c["gids_old"]])
        # local mesh through the oracle: its own first-touch numbering differs, so feed the local dof ids explicitly
        ml = O.Mesh.__new__(O.Mesh)
        ml.celltype, ml.nv, ml.dim = O.TET4, 4, 3
        ml.conn, ml.coords, ml.celldofs = c["lconn"], np.ascontiguousarray(c["lcoords"]), np.ascontiguousarray(c["ldofs"])
        ml.ncells, ml.nnodes, ml.ndofs, ml._pattern = c["lconn"].shape[0], c["lcoords"].shape[0], c["gids_old"].size, None
        rpl, cil = ml.pattern()
        Al = O.axpby_values(O.assemble_mass(ml, 2), O.assemble_diffusion(ml, 2, O.D_SCALAR, [0.3]), 0.05)
        yl = O.spmv(rpl, cil, Al, xl)[:no]
        ref = yg[c["gids_old"][:no]]
        assert np.abs(yl - ref).max() <= 1e-14 * np.abs(yg).max()                   # row sums in a different column order
        # owned rows hold the same VALUES as the global rows (same cells, same order of element contributions)
        for row in (0, no // 2, no - 1):
            g = c["gids_old"][row]
            lv_ = dict(zip(c["gids_old"][cil[rpl[row]:rpl[row + 1]]], Al[rpl[row]:rpl[row + 1]]))
            gv = dict(zip(ci[rp[g]:rp[g + 1]], Ag[rp[g]:rp[g + 1]]))
            assert lv_ == gv
