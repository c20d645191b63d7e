"""Generates the golden fixtures in this directory FROM THE ORACLE (the reference is pure Julia and
cannot run in this image, so there is nothing else to generate them from -- see DESIGN.md "Oracle").
They pin the oracle against accidental change and give the GPU tests input/output vectors that do
not need /root/reference or the oracle build at run time.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle as O  # noqa: E402

HERE = Path(__file__).resolve().parent


def activation_steps(phi_hist, thr):
    """first step index n with phi(t_n) >= thr after having been below it; -1 if never."""
    act = np.full(phi_hist.shape[1], -1, dtype=np.int64)
    below = phi_hist[0] < thr
    for n in range(1, phi_hist.shape[0]):
        hit = (phi_hist[n] >= thr) & below & (act < 0)
        act[hit] = n
        below |= phi_hist[n] < thr
    return act


def c1_small():
    """config 1 (ep01_spiral-wave.jl) on 32x32 quads: FHN, kappa scaled to keep kappa*dt/h^2 of the
    256x256 tutorial grid, spiral initial condition, dt = 1, default sqrt(eps) tolerances."""
    n = 32
    s = (256 / n) ** 2
    m = O.generate_grid(O.QUAD4, (n, n), (0.0, 0.0), (2.5, 2.5))
    M = O.assemble_mass(m, 2)
    K = O.assemble_diffusion(m, 2, O.D_TENSOR, [4.5e-5 * s, 0, 0, 2.0e-5 * s])
    x = m.dof_coords
    N = m.ndofs
    u = np.zeros(2 * N)
    u[:N] = np.where((x[:, 0] <= 1.25) & (x[:, 1] <= 1.25), 1.0, 0.0)
    u[N:] = np.where(x[:, 1] >= 1.25, 0.1, 0.0)
    out = {"u0": u.copy(), "kappa": np.array([4.5e-5 * s, 2.0e-5 * s])}
    for name, sub in (("fe", 1), ("adaptive", 10)):
        orc = O.MonodomainOracle(m, O.FHN, O.default_params(O.FHN), M, K, substeps=sub)
        v = u.copy()
        hist = [v[:N].copy()]
        for step in range(200):
            it, rn, conv = orc.step(v, float(step), 1.0)
            assert conv
            hist.append(v[:N].copy())
            if step == 0:
                out[f"u1_{name}"] = v.copy()
        out[f"u200_{name}"] = v.copy()
        out[f"iters_{name}"] = np.array(orc.iters)
        out[f"act_{name}"] = activation_steps(np.array(hist), 0.5)
    np.savez_compressed(HERE / "c1_small.npz", **out)


def c1_full():
    """config 1 at FULL size (ep01_spiral-wave.jl:30-65,113-141): 256x256 quads, FHN, dt = 1, 1000 steps, default
    sqrt(eps) tolerances.  Stored: phi_m after step 1 and after step 1000, CG iteration counts of all steps, and the
    activation step of every dof (threshold 0.5) -- the three things the parity rule is stated on."""
    n = 256
    m = O.generate_grid(O.QUAD4, (n, n), (0.0, 0.0), (2.5, 2.5))
    M = O.assemble_mass(m, 2)
    K = O.assemble_diffusion(m, 2, O.D_TENSOR, [4.5e-5, 0, 0, 2.0e-5])
    x = m.dof_coords
    N = m.ndofs
    u = np.zeros(2 * N)
    u[:N] = np.where((x[:, 0] <= 1.25) & (x[:, 1] <= 1.25), 1.0, 0.0)
    u[N:] = np.where(x[:, 1] >= 1.25, 0.1, 0.0)
    orc = O.MonodomainOracle(m, O.FHN, O.default_params(O.FHN), M, K, substeps=1)
    v = u.copy()
    act = np.full(N, -1, dtype=np.int16)
    below = v[:N] < 0.5
    out = {}
    for step in range(1000):
        it, rn, conv = orc.step(v, float(step), 1.0)
        assert conv
        hit = (v[:N] >= 0.5) & below & (act < 0)
        act[hit] = step + 1
        below |= v[:N] < 0.5
        if step == 0:
            out["phi1"] = v[:N].astype(np.float64).copy()
    out["phi1000"] = v[:N].copy()
    out["s1000"] = v[N:].copy()
    out["iters"] = np.array(orc.iters, dtype=np.int16)
    out["act"] = act
    np.savez_compressed(HERE / "c1_full.npz", **out)


def c2_full():
    """config 2 at FULL size (128x128x32 hexes, h = 0.25, PCG2019, dt = 0.01; conduction-velocity-benchmark.jl:29-53) with
    the corner stimulus, 40 steps of the adaptive substepper: phi_m sampled at every 61st dof after step 1 and step 40, all
    CG iteration counts, and the number of dofs above -40 mV (the stimulated corner)."""
    m = O.generate_grid(O.HEX8, (128, 128, 32), (0, 0, 0), (32.0, 32.0, 8.0))
    k1 = 0.17 * 0.62 / (0.17 + 0.62)
    kr = 0.019 * 0.24 / (0.019 + 0.24)
    M = O.assemble_mass(m, 2)
    K = O.assemble_diffusion(m, 2, O.D_TENSOR, np.diag([k1, kr, kr]))
    N = m.ndofs
    u = np.repeat(O.default_initial_state(O.PCG2019), N)
    orc = O.MonodomainOracle(m, O.PCG2019, O.default_params(O.PCG2019), M, K, substeps=10, threaded_blas1=False)
    out, t, dt = {}, 0.0, 0.01
    for step in range(40):
        orc.bS = O.assemble_source(m, 2, O.SRC_BOX, [1.5, 2.0, 0.5], t + dt)
        it, rn, conv = orc.step(u, t, dt)
        assert conv
        t += dt
        if step == 0:
            out["phi1"] = u[:N:61].copy()
    out["phi40"] = u[:N:61].copy()
    out["h40"] = u[N:2 * N:61].copy()
    out["iters"] = np.array(orc.iters, dtype=np.int16)
    out["n_above"] = np.array([(u[:N] > -84.0).sum()])
    np.savez_compressed(HERE / "c2_full.npz", **out)


def c2_full_1000():
    """config 2 at FULL size and FULL length: 1000 steps of dt = 0.01, adaptive substepper, corner stimulus
    with its [0, 2.1] window (needs_update gating, stale source afterwards: euler.jl:88-91).  Stored: phi_m at every 61st
    dof after steps 1 and 1000, the h gate after 1000, all CG iteration counts, activation steps (first step with
    phi_m >= 0 mV) of the sampled dofs."""
    m = O.generate_grid(O.HEX8, (128, 128, 32), (0, 0, 0), (32.0, 32.0, 8.0))
    k1 = 0.17 * 0.62 / (0.17 + 0.62)
    kr = 0.019 * 0.24 / (0.019 + 0.24)
    M = O.assemble_mass(m, 2, threaded=True)
    K = O.assemble_diffusion(m, 2, O.D_TENSOR, np.diag([k1, kr, kr]), threaded=True)
    N = m.ndofs
    u = np.repeat(O.default_initial_state(O.PCG2019), N)
    orc = O.MonodomainOracle(m, O.PCG2019, O.default_params(O.PCG2019), M, K, substeps=10, threaded_blas1=True)
    out, t, dt = {}, 0.0, 0.01
    act = np.full(u[:N:61].size, -1, dtype=np.int16)
    for step in range(1000):
        if 0.0 <= t + dt <= 2.1:
            orc.bS = O.assemble_source(m, 2, O.SRC_BOX, [1.5, 2.0, 0.5], t + dt)
        it, rn, conv = orc.step(u, t, dt)
        assert conv
        t += dt
        ph = u[:N:61]
        act[(act < 0) & (ph >= 0.0)] = step + 1
        if step == 0:
            out["phi1"] = ph.copy()
        if step % 100 == 99:
            print("step", step + 1, "iters", it, "activated", int((act > 0).sum()), flush=True)
    out["phi1000"] = u[:N:61].copy()
    out["h1000"] = u[N:2 * N:61].copy()
    out["iters"] = np.array(orc.iters, dtype=np.int16)
    out["act"] = act
    out["n_above"] = np.array([(u[:N] > -84.0).sum()])
    np.savez_compressed(HERE / "c2_full_1000.npz", **out)


def c4_mid_1000():
    """BASELINE config 4 at a realistic size and FULL length: idealized LV 120 x 12 x 80 rings, tetrahedralised (0.7 M tets,
    126 k dofs), ODB25LT fibres, spectral tensor, PCG2019, ForwardEulerCellSolver, endocardial stimulus, default CG
    tolerances, 1000 steps of dt = 0.01.  Stored: phi_m at every 13th dof after steps 1, 100 and 1000, all CG iteration
    counts, activation steps (phi_m >= 0 mV) of the sampled dofs."""
    sys.path.insert(0, str(HERE.parent.parent))
    from thunderbolt_jl_b200 import lv
    nodes, hexes, wedges, prm = lv.generate_ideal_lv_mesh(120, 12, 80)
    tets = lv.tetrahedralize(nodes, hexes, wedges)
    fsn = lv.odb25lt_fibres(prm, tets)
    k1, kr = 0.17 * 0.62 / (0.17 + 0.62), 0.019 * 0.24 / (0.019 + 0.24)
    data = np.concatenate([[k1, kr, kr], np.ascontiguousarray(fsn).reshape(tets.shape[0], 4, 9).ravel()])
    m = O.Mesh(O.TET4, tets, nodes)
    M = O.assemble_mass(m, 2, threaded=True)
    K = O.assemble_diffusion(m, 2, O.D_SPECTRAL, data, threaded=True)
    N = m.ndofs
    print("dofs", N, "tets", tets.shape[0], flush=True)
    u = np.repeat(O.default_initial_state(O.PCG2019), N)
    # dot mode 2: order-free (double-double) dot products, so that the iteration counts do not depend on a summation order
    # no other implementation could reproduce; the GPU test runs tb_cg_set_exact_dot against this golden
    orc = O.MonodomainOracle(m, O.PCG2019, O.default_params(O.PCG2019), M, K, threaded_blas1=2)
    SRC = [0.0, 0.2, 0.3, 0.25]
    out, t, dt = {}, 0.0, 0.01
    act = np.full(u[:N:13].size, -1, dtype=np.int16)
    for step in range(1000):
        orc.bS = O.assemble_source(m, 2, O.SRC_ENDO, SRC, t + dt)
        it, rn, conv = orc.step(u, t, dt)
        assert conv
        t += dt
        ph = u[:N:13]
        act[(act < 0) & (ph >= 0.0)] = step + 1
        if step == 0:
            out["phi1"] = ph.copy()
        if step == 99:
            out["phi100"] = ph.copy()
        if step % 50 == 49:
            print("step", step + 1, "iters", it, "activated", int((act > 0).sum()), "phi max", float(u[:N].max()), flush=True)
    out["phi1000"] = u[:N:13].copy()
    out["h1000"] = u[N:2 * N:13].copy()
    out["iters"] = np.array(orc.iters, dtype=np.int16)
    out["act"] = act
    np.savez_compressed(HERE / "c4_mid_1000.npz", **out)


def c2_small():
    """config 2 (conduction-velocity-benchmark.jl) on 16x16x4 hexes, h = 0.25: PCG2019, corner stimulus."""
    m = O.generate_grid(O.HEX8, (16, 16, 4), (0, 0, 0), (4.0, 4.0, 1.0))
    k1 = 0.17 * 0.62 / (0.17 + 0.62)
    kr = 0.019 * 0.24 / (0.019 + 0.24)
    M = O.assemble_mass(m, 2)
    K = O.assemble_diffusion(m, 2, O.D_TENSOR, np.diag([k1, kr, kr]))
    N = m.ndofs
    u = np.repeat(O.default_initial_state(O.PCG2019), N)
    out = {"u0": u.copy()}
    dt = 0.01
    for name, sub in (("fe", 1), ("adaptive", 10)):
        orc = O.MonodomainOracle(m, O.PCG2019, O.default_params(O.PCG2019), M, K, substeps=sub)
        v = u.copy()
        hist = [v[:N].copy()]
        t = 0.0
        for step in range(300):
            if 0.0 <= t + dt <= 2.1:
                orc.bS = O.assemble_source(m, 2, O.SRC_BOX, [1.5, 2.0, 0.5], t + dt)
            it, rn, conv = orc.step(v, t, dt)
            assert conv
            t += dt
            hist.append(v[:N].copy())
            if step == 0:
                out[f"u1_{name}"] = v.copy()
        out[f"u300_{name}"] = v.copy()
        out[f"iters_{name}"] = np.array(orc.iters)
        out[f"act_{name}"] = activation_steps(np.array(hist), 0.0)
    np.savez_compressed(HERE / "c2_small.npz", **out)


if __name__ == "__main__" and "--c2-full" in sys.argv:
    c2_full()
    print("wrote", HERE / "c2_full.npz")
    sys.exit(0)

if __name__ == "__main__" and "--c2-full-1000" in sys.argv:
    c2_full_1000()
    print("wrote", HERE / "c2_full_1000.npz")
    sys.exit(0)

if __name__ == "__main__" and "--c4-mid-1000" in sys.argv:
    c4_mid_1000()
    print("wrote", HERE / "c4_mid_1000.npz")
    sys.exit(0)

if __name__ == "__main__" and "--c1-full" in sys.argv:
    c1_full()
    print("wrote", HERE / "c1_full.npz")
    sys.exit(0)

if __name__ == "__main__":
    c1_small()
    c2_small()
    for f in sorted(HERE.glob("*.npz")):
        d = np.load(f)
        print(f.name, {k: d[k].shape for k in d.files})
