#!/usr/bin/env python
"""bench.py -- monodomain DoF*steps/s on B200 (BASELINE.json metric) + roofline + CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c1] [--impl b200|reference]

A "step" is one LieTrotterGodunov step (backward-Euler diffusion solved with CG + ionic cell sweep)
over the whole mesh.  Default workload = BASELINE config 5, the configuration the headline target is
quoted on: 3D hexahedral slab 512x512x384 (101,320,065 nodes), FHN, fp64 -- it fits one 180 GB B200.
N > 1 (torchrun, one rank per GPU) partitions the SAME mesh by dof ownership (z-slabs): strong scaling.

Printed JSON line (rank 0): see the contract in the task statement; additional keys `roofline`
(dominant kernel = SpMV inside CG, timed live with CUDA events on the library's stream),
`cpu_baseline` (the oracle = CPU restatement of the reference's algorithm, timed on this box's host
cores on a bounded z-slab of the same mesh) and `e2e` (same metric through tb_monodomain_step_host:
pinned host buffers, H2D + D2H of the state every step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

SQRT_EPS = float(np.sqrt(np.finfo(np.float64).eps))

# kappa chosen so that kappa*dt/h^2 equals the tutorial's (ep01_spiral-wave.jl: 4.5e-5, 2e-5 on h = 2.5/256, dt = 1)
WORKLOADS = {
    "c5": dict(name="C5 3D hex slab 512x512x384 FHN (BASELINE config 5)", celltype="hex", nel=(512, 512, 384), h=0.25,
               model="fhn", kappa=(0.0295, 0.0131, 0.0131), dt=1.0, substeps=1),
    "c2": dict(name="C2 3D hex slab 128x128x32 PCG2019 (BASELINE config 2)", celltype="hex", nel=(128, 128, 32), h=0.25,
               model="pcg2019", kappa=(0.17 * 0.62 / (0.17 + 0.62), 0.019 * 0.24 / (0.019 + 0.24), 0.019 * 0.24 / (0.019 + 0.24)),
               dt=0.01, substeps=1),
    "c4": dict(name="C4 idealized LV, tetrahedralised (240 x 24 x 160 rings -> 5.5 M tets), ODB25LT fibres, spectral tensor, "
                    "PCG2019 (BASELINE config 4)", celltype="lv", nel=(240, 24, 160), h=0.0, model="pcg2019",
               kappa=(0.17 * 0.62 / (0.17 + 0.62), 0.019 * 0.24 / (0.019 + 0.24), 0.019 * 0.24 / (0.019 + 0.24)),
               dt=0.01, substeps=10),
    "c1": dict(name="C1 2D quad 256x256 FHN spiral wave (BASELINE config 1)", celltype="quad", nel=(256, 256), h=2.5 / 256,
               model="fhn", kappa=(4.5e-5, 2.0e-5), dt=1.0, substeps=1),
}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def initial_state(x, model, lengths, tb):
    """Spiral-wave style initial condition of the tutorial (ep01_spiral-wave.jl:113-118) scaled to the box."""
    n = x.shape[0]
    if model == "fhn":
        u = np.zeros(2 * n)
        u[:n] = np.where((x[:, 0] <= 0.5 * lengths[0]) & (x[:, 1] <= 0.5 * lengths[1]), 1.0, 0.0)
        u[n:] = np.where(x[:, 1] >= 0.5 * lengths[1], 0.1, 0.0)
        return u
    u = np.repeat(tb.default_initial_state(tb.PCG2019()), n)
    u[:n] = np.where(np.all(x < 1.5, axis=1), 20.0, u[:n])   # a depolarised corner instead of a stimulus current
    return u


if __name__ != "__main__":
    sys.modules.setdefault("bench", sys.modules[__name__])


def host_cores():
    """Physical cores this process may use (lscpu sockets x cores/socket, capped by the affinity mask).  OMP_NUM_THREADS is
    deliberately ignored: torchrun exports OMP_NUM_THREADS=1 to every rank, which silently made the CPU arm single-threaded."""
    phys = None
    try:
        out = subprocess.run(["lscpu"], capture_output=True, text=True, timeout=10).stdout
        f = {k.strip(): v.strip() for k, v in (ln.split(":", 1) for ln in out.splitlines() if ":" in ln)}
        phys = int(f["Socket(s)"]) * int(f["Core(s) per socket"])
    except Exception:
        pass
    try:
        aff = len(os.sched_getaffinity(0))
    except Exception:
        aff = os.cpu_count() or 1
    return max(1, min(phys, aff) if phys else aff)


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and therefore the first-touch placement of its pinned buffers) to the CPUs of the NUMA
    node its GPU hangs off: /sys/bus/pci/devices/<bdf>/local_cpulist.  Returns a short description for the JSON line."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = Path("/sys/bus/pci/devices") / bdf
        node = (base / "numa_node").read_text().strip()
        cpus = set()
        for part in (base / "local_cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"numa node {node}, {len(cpus)} cpus"
        return f"numa node {node}, affinity unchanged"
    except Exception as e:                                  # noqa: BLE001 - best effort, containers often hide sysfs
        return f"not bound ({type(e).__name__})"


def bytes_per_row(nnz, n):
    """SpMV = nnzr*12 + 8 (rowptr) + 8 (x) + 8 (y) per row."""
    return nnz / n * 12.0 + 24.0


def config_of(args, W, nel):
    """What is configured (identical for the b200 and the reference arm); measured quantities go to `run_info`."""
    return {"workload": W["name"] if not args.grid else f"{W['name']} [grid override {list(nel)}]", "nel": list(nel),
            "cell_model": W["model"], "dt": W["dt"],
            "cell_solver": "ForwardEulerCellSolver" if W["substeps"] == 1 else f"AdaptiveForwardEulerSubstepper({W['substeps']})",
            "cg": {"atol": SQRT_EPS, "rtol": SQRT_EPS, "x0": "zero", "preconditioner": args.precond,
                   **({"bj_rows": args.bj_rows} if args.precond == "block_jacobi" else {}),
                   **({"degree": args.cheb_degree, "ratio": args.cheb_ratio} if args.precond == "chebyshev" else {})}}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import thunderbolt_jl_b200 as tb

    W = WORKLOADS[args.workload]
    nel = tuple(int(v) for v in args.grid.split(",")) if args.grid else W["nel"]
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else "single GPU: not bound"
    dev = tb.B200Device(local)
    tb.set_default_device(dev)
    if world > 1:
        from thunderbolt_jl_b200 import dist as tbd
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        tbd.init_comm(dev, dist)

    ct = tb.Hexahedron if W["celltype"] == "hex" else tb.Quadrilateral
    dim = len(nel)
    lengths = tuple(n * W["h"] for n in nel)
    t_setup = time.time()
    lv_data = None
    part = None
    if W["celltype"] == "lv":
        from thunderbolt_jl_b200 import lv
        nodes, hexes, wedges, prm = lv.generate_ideal_lv_mesh(*nel)
        tets = lv.tetrahedralize(nodes, hexes, wedges)
        fsn = np.ascontiguousarray(lv.odb25lt_fibres(prm, tets)).reshape(tets.shape[0], -1)
        dim = 3
        lengths = tuple(float(v) for v in (nodes.max(axis=0) - nodes.min(axis=0)))
        if world > 1:
            # general mesh: RCB ownership of the dofs, renumbering, HOST-side cut -- only the local cells reach the GPU
            celldofs, N_global = tb.api.close_dofs(tets)
            part = tbd.partition_host_mesh(dev, tb.Tetrahedron, tets, nodes, celldofs, N_global, dist)
            mesh = part.mesh
            fsn = fsn[part.cells]
        else:
            mesh = tb.to_mesh(tb.Tetrahedron, tets, nodes, device=dev)
            N_global = mesh.ndofs
        lv_data = np.concatenate([np.asarray(W["kappa"], dtype=np.float64), fsn.reshape(-1)])
        del fsn
    elif world > 1:
        # structured grid: every rank generates only its own cells from the closed-form first-touch numbering -- the global
        # grid never exists in HBM (tb_mesh_generate_grid_local)
        plane = (nel[0] + 1) * (nel[1] + 1) if dim == 3 else nel[0] + 1
        N_global = int(np.prod([n + 1 for n in nel]))
        if os.environ.get("TB_BENCH_LOCAL_GRID", "1") != "0":
            part = tbd.partition_structured_grid(dev, ct, nel, (0.0,) * dim, lengths, dist, plane=plane if args.cut == "planes" else None)
        else:   # round-1 path: the global grid on every rank, cut on the device
            full = tb.generate_mesh(ct, nel, (0.0,) * dim, lengths, device=dev)
            part = tbd.partition_mesh(dev, full, dist, plane=plane if args.cut == "planes" else None)
            full.free()
        mesh = part.mesh
    else:
        mesh = tb.generate_mesh(ct, nel, (0.0,) * dim, lengths, device=dev)
        N_global = mesh.ndofs
    ion = tb.FHNModel() if W["model"] == "fhn" else tb.PCG2019()
    ns = tb.num_states(ion)
    if world > 1 and part is None:
        plane = (nel[0] + 1) * (nel[1] + 1) if dim == 3 else nel[0] + 1
        part = tbd.partition_mesh(dev, mesh, dist, plane=plane if args.cut == "planes" else None)
        mesh.free()
        mesh = part.mesh
    M = tb.B200CSRMatrix.from_mesh(dev, mesh)
    K = M.like()
    tb.core.assemble_mass(dev, mesh, M, 2, 1.0)
    if lv_data is not None:
        tb.core.assemble_diffusion(dev, mesh, K, 2, tb._lib.D_SPECTRAL, lv_data, 1.0)
        del lv_data
    else:
        tb.core.assemble_diffusion(dev, mesh, K, 2, tb._lib.D_TENSOR, np.diag(W["kappa"][:dim]), 1.0)
    peer_path = False
    if world > 1:
        part.attach_halo(M)
        peer_path = bool(part.peer)
    asm_info = dev.assembly_info()
    dev.assembly_release_scratch()
    st = tb.MonodomainStepper(dev, M, K, ion.model_id, ion.params())
    st.set_cg(SQRT_EPS, SQRT_EPS, None)
    PC = {"none": tb._lib.PRECOND_NONE, "jacobi": tb._lib.PRECOND_JACOBI, "block_jacobi": tb._lib.PRECOND_BLOCK_JACOBI,
          "chebyshev": tb._lib.PRECOND_CHEBYSHEV}[args.precond]
    if args.precond == "block_jacobi":
        dev.cg_set_block_jacobi(mesh.ndofs_owned, max(1, mesh.ndofs_owned // args.bj_rows), None)
    if args.precond == "chebyshev":
        dev.cg_set_chebyshev(args.cheb_degree, args.cheb_ratio)
    st.set_preconditioner(PC)
    if args.cell_substeps:
        W = dict(W, substeps=args.cell_substeps)
    st.set_cell_solver(W["substeps"], 0.1)
    x = mesh.dof_coords()
    n_local = mesh.ndofs                                    # owned + ghosts
    n_owned = mesh.ndofs_owned
    if W["celltype"] == "lv":   # rest, apex region depolarised
        u0 = np.repeat(tb.default_initial_state(tb.PCG2019()), mesh.ndofs)
        zmax = float(nodes[:, 2].max())                      # global, not this rank's
        u0[:mesh.ndofs] = np.where(x[:, 2] > zmax - 0.15 * lengths[2], 20.0, u0[:mesh.ndofs])
    else:
        u0 = initial_state(x, W["model"], lengths, tb)
    if args.no_parity:
        del x
    u = tb.B200Vector.from_host(dev, u0, ns)
    nnz = M.nnz
    dev.sync()
    t_setup = time.time() - t_setup

    def barrier():
        dev.sync()
        if world > 1:
            dist.barrier()
        dev.sync()

    t, dt = 0.0, W["dt"]
    iters = []
    for _ in range(args.warmup):
        it, rn, conv = st.step(u, t, dt)
        t += dt
    # ---- timed region: exactly K steps, device timer, max over ranks ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev.profile_enable(True)
    launches0 = dev.launch_count()
    if world > 1:
        dev.peer_stats(reset=True)
    barrier()
    dev.timer_start()
    conv_all = True
    for _ in range(args.steps):
        it, rn, conv = st.step(u, t, dt)
        t += dt
        iters.append(it)
        conv_all &= conv
    ms = dev.timer_stop()
    barrier()
    launches = dev.launch_count() - launches0
    comm = None
    if world > 1:
        ps = dev.peer_stats()
        cw = torch.tensor([ps["ar_wait_ms"], ps["halo_wait_ms"]], dtype=torch.float64, device="cuda")
        cmax, cmin = cw.clone(), cw.clone()
        dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(cmin, op=dist.ReduceOp.MIN)
        comm = {"what": "time CTA 0 of a rank spent waiting for the other ranks inside the CG kernels (ld.acquire.sys spins), per step",
                "allreduce_wait_ms_per_step_max_rank": float(cmax[0]) / args.steps, "allreduce_wait_ms_per_step_min_rank": float(cmin[0]) / args.steps,
                "halo_wait_ms_per_step_max_rank": float(cmax[1]) / args.steps, "halo_wait_ms_per_step_min_rank": float(cmin[1]) / args.steps,
                "waits_per_step": ps["ar_waits"] / args.steps}
    spmv_ms, spmv_n = dev.profile_get()
    dev.profile_enable(False)
    cg_path = dev.cg_last_path()
    persistent = cg_path != 0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = N_global * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (SpMV + fused p.Ap inside CG) ----
    peak, peak_src = peaks()
    rows_local = n_owned
    # persistent single-kernel CG (small operators): the timed unit is one whole iteration (SpMV + vector updates)
    spmv_bytes = (bytes_per_row(nnz, rows_local) + (72.0 if persistent else 0.0)) * rows_local
    avg_spmv_ms = spmv_ms / max(spmv_n, 1)
    achieved = spmv_bytes / (avg_spmv_ms * 1e-3) / 1e9 if spmv_n else None
    k_mean = float(np.mean(iters))
    step_bytes = (2 * ns * 8 + bytes_per_row(nnz, rows_local) * (1 + k_mean) + 72.0 * k_mean) * rows_local
    # bytes the kernel really streams: stored values + the (losslessly compressed) column stream + x, y, p
    stored, col_bytes, max_w = M.storage()
    stored_bytes = stored * 8.0 + col_bytes + ({0: 24.0, 1: 16.0, 2: 24.0 + 72.0}[cg_path]) * rows_local
    traffic = None
    tfile = ROOT / "profiles" / "traffic.json"
    if tfile.exists() and world == 1 and not args.grid:
        traffic = json.loads(tfile.read_text()).get(args.workload, {}).get("spmv_dram_bytes_per_launch")
    kname = (("k_cg_persistent" if cg_path == 1 else "k_cg_persistent_tma") +
             " (whole CG solve in one cooperative kernel; unit = one iteration: SpMV + x,r,p updates + 3 grid barriers)"
             if persistent else "k_cg_spmv_tma<1,false,true> (SELL-32 SpMV, TMA-staged, fused p.Ap)")
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic, "traffic_source": "ncu --set full dram__bytes_read.sum+write.sum, profiles/traffic.json" if traffic else None,
                "bytes_per_launch": spmv_bytes, "stored_bytes_per_launch": stored_bytes,
                "stored_achieved_gbs": stored_bytes / (avg_spmv_ms * 1e-3) / 1e9 if spmv_n else None,
                "avg_launch_ms": avg_spmv_ms, "launches_timed": spmv_n,
                "spmv_share_of_step": spmv_ms / ms if ms else None,
                "step_bytes_per_dof": step_bytes / rows_local,
                "step_achieved_gbs": step_bytes * args.steps / (ms * 1e-3) / 1e9 * 1.0,
                "step_frac_of_peak": step_bytes * args.steps / (ms * 1e-3) / 1e9 / peak}

    # ---- small / mid-size operators: the same K steps through tb_monodomain_run, which enqueues the whole run without a
    # read-back when the CG is one persistent kernel (iteration counts folded on the device, fetched once) ----
    run_api = None
    if persistent and world == 1:
        dev.sync()
        dev.timer_start()
        tot_run, conv_run = st.run(u, t, dt, args.steps)
        ms_run = dev.timer_stop()
        t += args.steps * dt
        run_api = {"api": "tb_monodomain_run (no read-back per step: persistent CG kernel, totals folded on the device)",
                   "value": N_global * args.steps / (ms_run * 1e-3), "unit": "DoF*steps/s", "ms_per_step": ms_run / args.steps,
                   "cg_iters_per_step_mean": tot_run / args.steps, "all_converged": bool(conv_run)}

    # ---- e2e: same steps through the host-buffer entry point: the state lives in PINNED HOST memory between steps; every
    # step uploads the whole state and downloads the whole result (tb_monodomain_run_host pipelines the copies) ----
    e2e = None
    if args.e2e_steps != 0:
        hin = torch.empty(ns * n_local, dtype=torch.float64, pin_memory=True)
        hout = torch.empty(ns * n_local, dtype=torch.float64, pin_memory=True)
        hin.numpy()[:] = u.to_host()
        ke = max(1, args.steps if args.e2e_steps < 0 else args.e2e_steps)   # default: the same K steps as the device-timed region
        st.run_host(u, hin.numpy(), hout.numpy(), t, dt, 1)          # warm-up (creates the copy streams)
        barrier()
        t0 = time.perf_counter()
        st.run_host(u, hout.numpy(), hin.numpy(), t + dt, dt, ke)
        dev.sync()
        te = time.perf_counter() - t0
        if world > 1:                                                 # max over ranks, like the device-timed value
            tt = torch.tensor([te], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            te = float(tt.item())
            nb = torch.tensor([float(ns * n_local * 8)], dtype=torch.float64, device="cuda")
            dist.all_reduce(nb, op=dist.ReduceOp.SUM)
            bytes_step = int(nb.item())
        else:
            bytes_step = ns * n_local * 8
        t += (ke + 1) * dt
        sweep = None
        if args.e2e_chunk_sweep:      # experiment: pieces of the phi column in the chunk-chased round trip
            sweep = {}
            for nch in (int(v) for v in args.e2e_chunk_sweep.split(",")):
                st.set_host_chunks(nch)
                st.run_host(u, hout.numpy(), hin.numpy(), t, dt, 1)
                barrier()
                t0 = time.perf_counter()
                st.run_host(u, hin.numpy(), hout.numpy(), t + dt, dt, ke)
                dev.sync()
                ts = time.perf_counter() - t0
                if world > 1:
                    tt = torch.tensor([ts], dtype=torch.float64, device="cuda")
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    ts = float(tt.item())
                sweep[str(nch)] = ts / ke * 1e3
                t += (ke + 1) * dt
            st.set_host_chunks(16)
        # what the host link gives each rank while ALL ranks copy at once, both directions at once (the platform bound of e2e)
        nb_probe = min(ns * n_local, 1 << 25)
        dprobe = [torch.empty(nb_probe, dtype=torch.float64, device="cuda") for _ in range(2)]
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        barrier()
        torch.cuda.synchronize()
        tp = time.perf_counter()
        for _ in range(4):
            with torch.cuda.stream(s1):
                dprobe[0].copy_(hin[:nb_probe], non_blocking=True)
            with torch.cuda.stream(s2):
                hout[:nb_probe].copy_(dprobe[1], non_blocking=True)
        torch.cuda.synchronize()
        tp = time.perf_counter() - tp
        link = 4 * nb_probe * 8 / tp / 1e9
        if world > 1:
            lk = torch.tensor([link], dtype=torch.float64, device="cuda")
            dist.all_reduce(lk, op=dist.ReduceOp.MIN)
            link = float(lk.item())
        del dprobe
        phi_bytes_rank = n_local * 8
        e2e = {"host_binding": numa, "link_gbs_per_direction_per_gpu_all_ranks_active": link,
               "exposed_ms_model": phi_bytes_rank / (link * 1e9) * 1e3,
               "exposed_ms_measured": te / ke * 1e3 - ms / args.steps,
               "value": N_global * ke / te, "unit": "DoF*steps/s", "h2d_bytes_per_step": bytes_step,
               "d2h_bytes_per_step": bytes_step, "steps": ke, "ms_per_step": te / ke * 1e3,
               **({"ms_per_step_by_phi_chunks": sweep} if sweep else {}),
               "api": "tb_monodomain_run_host (C ABI): state in pinned host buffers between steps, full state H2D + D2H every "
                      "step, copies pipelined on two copy streams (phi download/upload full duplex in chunks, other columns under CG); exposed time per step = the phi column's "
                      "round trip over the host link (download of step n, upload as step n+1's input), which no schedule can hide because "
                      "step n+1's solve needs it: exposed_ms_model = phi bytes / measured link rate with all ranks copying"}

    # ---- parity evidence on the benchmarked configuration, at this GPU count (scripts/parity_block.py) ----
    parity = None
    if not args.no_parity:
        from scripts import parity_block
        gname = ("c5" if args.workload == "c5" else args.workload) + "_checksum.json"
        parity = parity_block.run(tb, dev, mesh, M, K, st, ion, W, nel, u, u0, x, dist=dist if world > 1 else None, world=world,
                                  rank=rank, precond=PC,
                                  write_golden=args.write_golden, golden_path=ROOT / "tests" / "golden" / gname)
        del x

    # ---- CPU baseline: the oracle on this box's host cores, bounded z-slab of the same mesh ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and W["celltype"] != "lv":
        cpu = cpu_baseline(args, W, nel)

    if rank == 0:
        line = {
            "metric": "monodomain DoF*steps/s", "value": value, "unit": "DoF*steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args, W, nel),
            "run_info": {"dofs": N_global, "nnz": int(nnz) if world == 1 else None,
                         "cg_iters_per_step_mean": k_mean, "cg_iters_min": int(min(iters)), "cg_iters_max": int(max(iters)),
                         "all_converged": bool(conv_all),
                         "parallelism": "single GPU" if world == 1 else
                         (f"{'RCB dof ownership (general mesh, host-side cut)' if W['celltype'] == 'lv' else 'dof-ownership z-slabs'} x{world}, halo of p and dot products by NVLink peer stores from the CG kernels "
                          f"(CUDA IPC windows; {'collects and halo push fused into the update kernels' if part.fused else 'separate push/collect kernels'}), "
                          f"NCCL for the per-step phi halo" if peer_path else
                          f"dof-ownership z-slabs x{world}, NCCL halo + allreduce") + (f", cuts at {args.cut}" if world > 1 else ""),
                         "assembly": {2: "element matrices + ordered gather (deterministic)", 0: "fp64 atomic scatter"}.get(
                             asm_info["last_mode"], str(asm_info["last_mode"])) + f", {asm_info['last_chunks']} chunk(s)",
                         "l2": "working set (matrix + vectors) is far larger than the 126 MB L2, no flush needed"
                               if nnz * 12 > 4e8 else "working set fits L2: tb_l2_flush not applied between steps (steady-state regime)",
                         "setup_s": t_setup},
            "parity": parity, "comm": comm,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, **({"run_api": run_api} if run_api else {}),
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, W, nel, all_threads=True):
    """Oracle (CPU restatement of the reference's algorithm, threads over nodes/rows like the reference)
    on a thin z-slab (or sub-square) of the SAME mesh spacing/physics: ~10-30 s of CPU work."""
    import oracle as O
    cores = host_cores()
    O.set_num_threads(cores)
    dim = len(nel)
    if dim == 3:
        snel = (min(nel[0], 512), min(nel[1], 512), min(nel[2], args.cpu_inline_layers))
        ct = O.HEX8
    else:
        snel = (min(nel[0], 256), min(nel[1], 256))
        ct = O.QUAD4
    lengths = tuple(n * W["h"] for n in snel)
    m = O.generate_grid(ct, snel, (0.0,) * dim, lengths)
    Mv = O.assemble_mass(m, 2, threaded=True)
    Kv = O.assemble_diffusion(m, 2, O.D_TENSOR, np.diag(W["kappa"][:dim]), threaded=True)
    model = O.FHN if W["model"] == "fhn" else O.PCG2019

    class _T:
        default_initial_state = staticmethod(lambda ion: O.default_initial_state(O.PCG2019))
        PCG2019 = staticmethod(lambda: None)
    u = initial_state(m.dof_coords, W["model"], lengths, _T)
    # variant B of BASELINE.md: threaded vector ops (the stronger baseline); variant A (serial BLAS-1,
    # what Krylov.jl does on Vector{Float64}) is reported beside it
    out = {}
    for name, thr in (("threaded_blas1", True), ("serial_blas1", False)):
        orc = O.MonodomainOracle(m, model, O.default_params(model), Mv, Kv, substeps=W["substeps"], threaded_blas1=thr)
        v = u.copy()
        orc.step(v, 0.0, W["dt"])                       # warm-up (builds A)
        t0 = time.perf_counter()
        nst = args.cpu_steps
        for s in range(nst):
            orc.step(v, (s + 1) * W["dt"], W["dt"])
        dtw = time.perf_counter() - t0
        out[name] = m.ndofs * nst / dtw
        out[name + "_iters"] = float(np.mean(orc.iters[1:]))
    return {"value": out["threaded_blas1"], "unit": "DoF*steps/s", "cores": cores, "cores_physical": cores, "kind": "port",
            "sample": f"{'x'.join(map(str, snel))} cells ({m.ndofs} DoFs) of the same mesh, {args.cpu_steps} steps after 1 warm-up; "
                      f"oracle = C/OpenMP restatement of the reference's CPU algorithm (not Julia)",
            "serial_blas1_value": out["serial_blas1"], "cg_iters_mean": out["threaded_blas1_iters"]}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is pure Julia
    and there is no julia binary in this image (and nothing to compile into oracle/_ref), so this arm
    times the oracle port with all host threads on a bounded sample, as the task statement prescribes."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    W = WORKLOADS[args.workload]
    if W["celltype"] == "lv":
        print(json.dumps({"impl": "reference", "unavailable": "the CPU arm is implemented for the structured workloads (c1, c2, c5) only"}))
        return
    nel = tuple(int(v) for v in args.grid.split(",")) if args.grid else W["nel"]
    import oracle as O
    cores = host_cores()
    O.set_num_threads(cores)        # explicit: torchrun exports OMP_NUM_THREADS=1
    dim = len(nel)
    snel = (min(nel[0], 512), min(nel[1], 512), min(nel[2], args.cpu_layers)) if dim == 3 else (min(nel[0], 256), min(nel[1], 256))
    ct = O.HEX8 if dim == 3 else O.QUAD4
    lengths = tuple(n * W["h"] for n in snel)
    t_setup = time.time()
    m = O.generate_grid(ct, snel, (0.0,) * dim, lengths)
    Mv = O.assemble_mass(m, 2, threaded=True)
    Kv = O.assemble_diffusion(m, 2, O.D_TENSOR, np.diag(W["kappa"][:dim]), threaded=True)
    t_setup = time.time() - t_setup
    model = O.FHN if W["model"] == "fhn" else O.PCG2019

    class _T:
        default_initial_state = staticmethod(lambda ion: O.default_initial_state(O.PCG2019))
        PCG2019 = staticmethod(lambda: None)
    u = initial_state(m.dof_coords, W["model"], lengths, _T)
    orc = O.MonodomainOracle(m, model, O.default_params(model), Mv, Kv, substeps=W["substeps"], threaded_blas1=True)
    t = 0.0
    for _ in range(args.warmup):
        orc.step(u, t, W["dt"]); t += W["dt"]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.step(u, t, W["dt"]); t += W["dt"]
    el = time.perf_counter() - t0
    value = m.ndofs * args.steps / el
    frac = f"1/{nel[2] // snel[2]} slab" if dim == 3 and snel[2] < nel[2] else "whole mesh"
    sample = (f"{'x'.join(map(str, snel))} cells ({m.ndofs} DoFs, {frac}) of the same mesh spacing, physics and initial condition; "
              f"oracle = C/OpenMP restatement of the reference's CPU algorithm (threads over nodes / rows, threaded BLAS-1), not Julia; "
              f"DoF*steps/s is size-intensive once the working set ({(Mv.nbytes * 2 + m.pattern()[1].nbytes) / 1e9:.1f} GB) is out of cache")
    print(json.dumps({
        "impl": "reference", "metric": "monodomain DoF*steps/s", "value": value, "unit": "DoF*steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args, W, nel),
        "run_info": {"sample_dofs": m.ndofs, "cg_iters_per_step_mean": float(np.mean(orc.iters[-args.steps:])), "setup_s": t_setup},
        "cpu_baseline": {"value": value, "unit": "DoF*steps/s", "cores": cores, "cores_physical": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "DoF*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--grid", default="", help="override cells per direction, e.g. 128,128,32")
    ap.add_argument("--e2e-steps", type=int, default=-1, help="steps of the end-to-end (host-buffer) measurement; -1 = same as --steps, 0 = skip")
    ap.add_argument("--e2e-chunk-sweep", default="", help="experiment: also time the e2e loop with these phi chunk counts, e.g. 2,4,8,32")
    ap.add_argument("--cpu-layers", type=int, default=48,
                    help="z-layers of cells in the --impl reference sample (48 = the 1/8 slab of C5 BASELINE.md states)")
    ap.add_argument("--cpu-inline-layers", type=int, default=16, help="z-layers of the cpu_baseline sample printed by the b200 arm")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block")
    ap.add_argument("--cell-substeps", type=int, default=0, help="override the workload's cell solver: N > 1 = AdaptiveForwardEulerSubstepper(N)")
    ap.add_argument("--write-golden", action="store_true", help="N=1 only: store this run's parity sequence as the gpu_n1 section of the golden file")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cut", default="rows", choices=["planes", "rows"],
                    help="multi-GPU ownership cuts: at grid-plane boundaries (one-plane halos) or at equal row counts (balanced)")
    ap.add_argument("--precond", default="none", choices=["none", "jacobi", "block_jacobi", "chebyshev"], help="inner CG preconditioner (SURVEY 8f-2)")
    ap.add_argument("--bj-rows", type=int, default=64, help="block_jacobi: rows per block (contiguous ranges of the dof numbering)")
    ap.add_argument("--cheb-degree", type=int, default=8)
    ap.add_argument("--cheb-ratio", type=float, default=100.0)
    args = ap.parse_args()
    if args.warmup < 3:
        print("note: the timing rules ask for >= 3 warm-up steps", file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
