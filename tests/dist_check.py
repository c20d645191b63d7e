"""Run under torchrun (one rank per GPU): the row-partitioned multi-GPU path against the single-GPU path
on the same mesh and inputs.  Rank 0 also solves the whole problem alone; every rank compares its owned
rows.  Used by tests/test_gpu_multi.py and by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import thunderbolt_jl_b200 as tb  # noqa: E402
from thunderbolt_jl_b200 import dist as tbd  # noqa: E402


def build(dev, mesh, ion, D):
    M = tb.B200CSRMatrix.from_mesh(dev, mesh)
    K = M.like()
    tb.core.assemble_mass(dev, mesh, M, 2, 1.0)
    tb.core.assemble_diffusion(dev, mesh, K, 2, tb._lib.D_TENSOR, D, 1.0)
    return M, K


def main_general(rank, world, local, dev):
    """DIST_MESH=lv: an unstructured mesh (idealized LV, tetrahedra, fibre tensor) partitioned by RCB on the host
    (dist.partition_host_mesh: renumbering + local cut, no global mesh in HBM) against the single-GPU solve."""
    from thunderbolt_jl_b200 import lv
    nodes, hexes, wedges, prm = lv.generate_ideal_lv_mesh(24, 3, 12)
    tets = lv.tetrahedralize(nodes, hexes, wedges)
    fsn = np.ascontiguousarray(lv.odb25lt_fibres(prm, tets)).reshape(tets.shape[0], -1)
    kap = [0.17 * 0.62 / (0.17 + 0.62), 0.019 * 0.24 / (0.019 + 0.24), 0.019 * 0.24 / (0.019 + 0.24)]
    ion = tb.FHNModel()
    dev1 = tb.B200Device(local)
    full1 = tb.to_mesh(tb.Tetrahedron, tets, nodes, device=dev1)
    N = full1.ndofs

    def build_lv(d, mesh, cellsel):
        M = tb.B200CSRMatrix.from_mesh(d, mesh)
        K = M.like()
        tb.core.assemble_mass(d, mesh, M, 2, 1.0)
        tb.core.assemble_diffusion(d, mesh, K, 2, tb._lib.D_SPECTRAL, np.concatenate([kap, fsn[cellsel].ravel()]), 1.0)
        return M, K
    x = full1.dof_coords()
    u0 = np.concatenate([np.where(x[:, 2] > x[:, 2].max() - 0.4, 1.0, 0.0), np.zeros(N)])
    M1, K1 = build_lv(dev1, full1, slice(None))
    st1 = tb.MonodomainStepper(dev1, M1, K1, ion.model_id, ion.params())
    u1 = tb.B200Vector.from_host(dev1, u0, 2)
    it1 = [st1.step(u1, 0.05 * s, 0.05)[0] for s in range(5)]
    ref = u1.to_host()
    celldofs, nd = tb.api.close_dofs(tets)
    assert nd == N
    part = tbd.partition_host_mesh(dev, tb.Tetrahedron, tets, nodes, celldofs, N, dist)
    lm = part.mesh
    no, nl = lm.ndofs_owned, lm.ndofs
    assert np.array_equal(lm.dof_coords(), x[part.gids_old])
    M, K = build_lv(dev, lm, part.cells)
    part.attach_halo(M)
    if rank == 0:
        print(("peer path on" if part.peer else "peer path off (NCCL)") + (", fused" if part.fused else ", unfused"), flush=True)
    st = tb.MonodomainStepper(dev, M, K, ion.model_id, ion.params())
    u = tb.B200Vector.from_host(dev, part.local_vector(u0, 2), 2)
    its = [st.step(u, 0.05 * s, 0.05)[0] for s in range(5)]
    got = u.to_host()
    own = part.gids_old[:no]
    err_phi = np.abs(got[:no] - ref[own]).max() / np.abs(ref[:N]).max()
    err_s = np.abs(got[nl:nl + no] - ref[N + own]).max()
    ok = err_phi <= 1e-8 and err_s <= 1e-9 and max(abs(a - b) for a, b in zip(its, it1)) <= 1
    # Chebyshev-preconditioned CG across the partition (inner SpMVs exchange their halo through NCCL)
    for d_, s_ in ((dev1, st1), (dev, st)):
        d_.cg_set_chebyshev(6, 30.0)
        s_.set_preconditioner(tb._lib.PRECOND_CHEBYSHEV)
    u1.upload(u0)
    u.upload(part.local_vector(u0, 2))
    itc1 = [st1.step(u1, 0.05 * s, 0.05)[0] for s in range(3)]
    itc = [st.step(u, 0.05 * s, 0.05)[0] for s in range(3)]
    refc, gotc = u1.to_host(), u.to_host()
    err_c = np.abs(gotc[:no] - refc[own]).max() / np.abs(refc[:N]).max()
    ok = ok and err_c <= 1e-8 and max(abs(a - b) for a, b in zip(itc, itc1)) <= 1 and max(itc1) < max(it1)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print(f"rank {rank}/{world} [lv, rcb]: owned {no} ghosts {nl - no} neighbours {list(part.plan.neigh_ranks)} iters {its} vs {it1} "
          f"err_phi {err_phi:.2e} err_s {err_s:.2e} | chebyshev iters {itc} vs {itc1} err {err_c:.2e} {'OK' if ok else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = tb.B200Device(local)
    tbd.init_comm(dev, dist)
    if os.environ.get("DIST_MESH", "grid") == "lv":
        return main_general(rank, world, local, dev)
    nel = (24, 20, 4 * world + 3)
    lengths = tuple(0.25 * n for n in nel)
    D = np.diag([0.0295, 0.0131, 0.0131])
    ion = tb.FHNModel()
    dev1 = tb.B200Device(local)      # a second context WITHOUT a communicator for the single-GPU reference solve
    full = tb.generate_mesh(tb.Hexahedron, nel, (0, 0, 0), lengths, device=dev)
    full1 = tb.generate_mesh(tb.Hexahedron, nel, (0, 0, 0), lengths, device=dev1)
    N = full.ndofs
    x = full.dof_coords()
    rng = np.random.default_rng(0)
    u0 = np.concatenate([np.where(x[:, 0] <= 0.5 * lengths[0], 1.0, 0.0) + 0.01 * rng.standard_normal(N),
                         np.where(x[:, 1] >= 0.5 * lengths[1], 0.1, 0.0)])
    # --- reference: the whole problem on this rank's GPU alone -------------------------------------
    M1, K1 = build(dev1, full1, ion, D)
    st1 = tb.MonodomainStepper(dev1, M1, K1, ion.model_id, ion.params())
    st1.set_cell_solver(10, 0.1)
    u1 = tb.B200Vector.from_host(dev1, u0, 2)
    it1 = [st1.step(u1, float(s), 1.0)[0] for s in range(5)]
    ref = u1.to_host()
    # --- partitioned ---------------------------------------------------------------------------------
    # DIST_CUT=rows: ownership cut at equal row counts instead of grid planes (ragged halos of up to two planes)
    plane = None if os.environ.get("DIST_CUT", "planes") == "rows" else (nel[0] + 1) * (nel[1] + 1)
    part = tbd.partition_mesh(dev, full, dist, plane=plane)
    lm = part.mesh
    lo, hi = int(part.bounds[rank]), int(part.bounds[rank + 1])
    gids = np.concatenate([np.arange(lo, hi), lm.ghost_global])       # local -> global dof id
    assert lm.ndofs == gids.size and np.array_equal(lm.dof_coords(), x[gids])
    M, K = build(dev, lm, ion, D)
    part.attach_halo(M)
    if rank == 0:
        print("peer path on" if part.peer else "peer path off (NCCL)", flush=True)
    # owned rows of the distributed operators equal the same rows of the global ones (pattern bit exact)
    rp, ci = M.pattern()
    rp1, ci1 = M1.pattern()
    assert np.array_equal(np.diff(rp), np.diff(rp1[lo:hi + 1]))
    # local columns are sorted by LOCAL id (owned first, ghosts after), so map to global ids and re-sort per row
    rows = np.repeat(np.arange(hi - lo), np.diff(rp))
    order = np.lexsort((gids[ci], rows))
    assert np.array_equal(gids[ci][order], ci1[rp1[lo]:rp1[hi]])
    v, v1 = M.nonzeros()[order], M1.nonzeros()[rp1[lo]:rp1[hi]]
    assert np.abs(v - v1).max() <= 1e-13 * np.abs(v1).max()
    st = tb.MonodomainStepper(dev, M, K, ion.model_id, ion.params())
    st.set_cell_solver(10, 0.1)
    ul = np.concatenate([u0[:N][gids], u0[N:][gids]])
    u = tb.B200Vector.from_host(dev, ul, 2)
    its = [st.step(u, float(s), 1.0)[0] for s in range(5)]
    got = u.to_host()
    no = hi - lo
    nl = lm.ndofs
    err_phi = np.abs(got[:no] - ref[lo:hi]).max() / np.abs(ref[:N]).max()
    err_s = np.abs(got[nl:nl + no] - ref[N + lo:N + hi]).max()
    ok = err_phi <= 1e-9 and err_s <= 1e-10 and max(abs(a - b) for a, b in zip(its, it1)) <= 1
    if os.environ.get("TB_DOT_EXACT", "0") == "1" and part.peer:
        # order-independent dot products: what is left between the partitioned and the single-GPU solve is the summation
        # order INSIDE boundary rows (local columns are owned-first, ghosts-last, so a row with lower-numbered ghosts adds
        # them last) -- rounding of single row sums, no reduction-order noise: equal iteration counts, 1e-13 on phi
        ok = ok and err_phi <= 1e-13 and err_s <= 1e-13 and its == it1
    # Jacobi-preconditioned CG through the same (fused / unfused / NCCL) communication path
    st1.set_preconditioner(tb._lib.PRECOND_JACOBI)
    st.set_preconditioner(tb._lib.PRECOND_JACOBI)
    u1.upload(u0)
    u.upload(ul)
    itp1 = [st1.step(u1, float(s), 1.0)[0] for s in range(3)]
    itp = [st.step(u, float(s), 1.0)[0] for s in range(3)]
    refp, gotp = u1.to_host(), u.to_host()
    err_pc = np.abs(gotp[:no] - refp[lo:hi]).max() / np.abs(refp[:N]).max()
    ok = ok and err_pc <= 1e-9 and max(abs(a - b) for a, b in zip(itp, itp1)) <= 1 and max(itp1) <= max(it1)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print(f"rank {rank}/{world}: owned [{lo},{hi}) ghosts {lm.ghost_global.size} iters {its} vs {it1} "
          f"err_phi {err_phi:.2e} err_s {err_s:.2e} | jacobi iters {itp} vs {itp1} err {err_pc:.2e} {'OK' if ok else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
