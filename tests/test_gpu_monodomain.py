"""-m gpu: the LieTrotterGodunov(BackwardEuler + CG, cell solver) step through the C ABI.

north_star parity rules checked here:
  * phi_m within 1e-10 relative L-inf after ONE split step, within 1e-6 after many steps
  * CG iteration counts within +-1 of the CPU path
  * activation times identical at time-step resolution
against (a) the oracle on the same seeded inputs and (b) the committed golden fixtures.
"""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["persistent", "persistent_tma", "multikernel"], autouse=True)
def cg_path(request, dev):
    """Every test of this module runs with all three CG execution models: the one-launch cooperative kernels that small
    (register-resident) and mid-size (TMA sweep) operators get by default (tb_cg_small.cu) and the
    three-kernels-per-iteration path large and multi-GPU ones take (tb_cg.cu)."""
    dev.cg_set_persistent({"persistent": 1, "persistent_tma": 2, "multikernel": 0}[request.param])
    yield request.param
    dev.cg_set_persistent(1)
GOLD = Path(__file__).resolve().parent / "golden"


def rel_linf(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def activation_steps(phi_hist, thr):
    act = np.full(phi_hist.shape[1], -1, dtype=np.int64)
    below = phi_hist[0] < thr
    for n in range(1, phi_hist.shape[0]):
        hit = (phi_hist[n] >= thr) & below & (act < 0)
        act[hit] = n
        below |= phi_hist[n] < thr
    return act


def _c1_model(tb, kappa, ion=None):
    return tb.MonodomainModel(tb.ConstantCoefficient(1.0), tb.ConstantCoefficient(1.0),
                              tb.ConstantCoefficient(tb.SymmetricTensor(2, (kappa[0], 0, kappa[1]))),
                              tb.NoStimulationProtocol(), ion or tb.FHNModel(), "φₘ", "s")


@pytest.mark.parametrize("variant,cell", [("fe", "fe"), ("adaptive", "adaptive")])
def test_c1_small_against_golden(tb, dev, variant, cell):
    """config 1 through the reference-shaped API (ep01_spiral-wave.jl), 200 steps."""
    g = np.load(GOLD / "c1_small.npz")
    mesh = tb.generate_mesh(tb.Quadrilateral, (32, 32), (0.0, 0.0), (2.5, 2.5), device=dev)
    odeform = tb.semidiscretize(tb.ReactionDiffusionSplit(_c1_model(tb, g["kappa"])),
                                tb.FiniteElementDiscretization({"φₘ": tb.LagrangeCollection(1)}), mesh)
    u0 = tb.create_initial_condition(odeform)
    tb.setvariable_(u0, odeform, "φₘ", lambda x: 1.0 if (x[0] <= 1.25 and x[1] <= 1.25) else 0.0)
    tb.setvariable_(u0, odeform, "s", lambda x: 0.1 if x[1] >= 1.25 else 0.0)
    assert np.array_equal(u0, g["u0"])
    cell_solver = tb.ForwardEulerCellSolver() if cell == "fe" else tb.AdaptiveForwardEulerSubstepper(reaction_threshold=0.1)
    stepper = tb.LieTrotterGodunov((tb.BackwardEulerSolver(), cell_solver))
    integ = tb.init(tb.OperatorSplittingProblem(odeform, u0, (0.0, 200.0)), stepper, dt=1.0)
    N = mesh.ndofs
    assert tb.step_(integ)
    u1 = integ.u.to_host()
    assert rel_linf(u1[:N], g[f"u1_{variant}"][:N]) <= 1e-10            # one split step
    assert rel_linf(u1[N:], g[f"u1_{variant}"][N:]) <= 1e-10
    hist = [g["u0"][:N], u1[:N]]
    while integ.t < 200.0 - 1e-9:
        assert tb.step_(integ)
        hist.append(integ.u.column(0))
    u = integ.u.to_host()
    assert rel_linf(u[:N], g[f"u200_{variant}"][:N]) <= 1e-6            # many steps
    it = np.array(integ.cg_iterations)
    assert it.shape == g[f"iters_{variant}"].shape and np.abs(it - g[f"iters_{variant}"]).max() <= 1
    assert np.array_equal(activation_steps(np.array(hist), 0.5), g[f"act_{variant}"])
    assert integ.stats.naccept == 200 and integ.stats.nreject == 0


def test_c1_full_size_1000_steps_against_golden(tb, dev):
    """BASELINE config 1 at its FULL size and length (256x256 quads, 1000 steps of dt = 1, ep01_spiral-wave.jl): the three
    north_star rules -- 1e-10 after one step, 1e-6 after 1000 steps, CG iterations +-1, activation steps identical."""
    g = np.load(GOLD / "c1_full.npz")
    mesh = tb.generate_mesh(tb.Quadrilateral, (256, 256), (0.0, 0.0), (2.5, 2.5), device=dev)
    odeform = tb.semidiscretize(tb.ReactionDiffusionSplit(_c1_model(tb, (4.5e-5, 2.0e-5))),
                                tb.FiniteElementDiscretization({"φₘ": tb.LagrangeCollection(1)}), mesh)
    u0 = tb.create_initial_condition(odeform)
    tb.setvariable_(u0, odeform, "φₘ", lambda x: 1.0 if (x[0] <= 1.25 and x[1] <= 1.25) else 0.0)
    tb.setvariable_(u0, odeform, "s", lambda x: 0.1 if x[1] >= 1.25 else 0.0)
    N = mesh.ndofs
    assert N == 66049
    stepper = tb.LieTrotterGodunov((tb.BackwardEulerSolver(), tb.ForwardEulerCellSolver()))
    integ = tb.init(tb.OperatorSplittingProblem(odeform, u0, (0.0, 1000.0)), stepper, dt=1.0)
    act = np.full(N, -1, dtype=np.int16)
    below = u0[:N] < 0.5
    for step in range(1000):
        assert tb.step_(integ)
        phi = integ.u.column(0)
        hit = (phi >= 0.5) & below & (act < 0)
        act[hit] = step + 1
        below |= phi < 0.5
        if step == 0:
            assert rel_linf(phi, g["phi1"]) <= 1e-10
    u = integ.u.to_host()
    assert rel_linf(u[:N], g["phi1000"]) <= 1e-6 and rel_linf(u[N:], g["s1000"]) <= 1e-6
    it = np.array(integ.cg_iterations)
    assert it.shape == g["iters"].shape and np.abs(it - g["iters"]).max() <= 1
    assert np.array_equal(act, g["act"])
    assert (act > 0).sum() > N // 2                                      # the spiral wave really swept the domain


@pytest.mark.parametrize("variant,sub", [("fe", 1), ("adaptive", 10)])
def test_c2_small_against_golden(tb, dev, variant, sub):
    """config 2 (PCG2019 + corner stimulus with a time window) through the raw C-ABI objects, 300 steps."""
    g = np.load(GOLD / "c2_small.npz")
    md = tb.generate_mesh(tb.Hexahedron, (16, 16, 4), (0, 0, 0), (4.0, 4.0, 1.0), device=dev)
    k1, kr = 0.17 * 0.62 / (0.17 + 0.62), 0.019 * 0.24 / (0.019 + 0.24)
    M = tb.B200CSRMatrix.from_mesh(dev, md)
    K = M.like()
    tb.core.assemble_mass(dev, md, M, 2, 1.0)
    tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_TENSOR, np.diag([k1, kr, kr]), 1.0)
    ion = tb.PCG2019()
    st = tb.MonodomainStepper(dev, M, K, ion.model_id, ion.params())
    st.set_cell_solver(sub, 0.1)
    N = md.ndofs
    u = tb.B200Vector.from_host(dev, g["u0"], 7)
    bS = tb.B200Vector(dev, N, 1)
    dt, t = 0.01, 0.0
    iters, hist = [], [g["u0"][:N]]
    for step in range(300):
        if 0.0 <= t + dt <= 2.1:                                         # needs_update, closed interval
            tb.core.assemble_source(dev, md, bS, 2, tb._lib.SRC_BOX, [1.5, 2.0, 0.5], t + dt)
            st.set_source(bS)
        it, rn, conv = st.step(u, t, dt)
        assert conv
        iters.append(it)
        t += dt
        hist.append(u.column(0))
        if step == 0:
            u1 = u.to_host()
            assert rel_linf(u1[:N], g[f"u1_{variant}"][:N]) <= 1e-10
            for s in range(1, 7):
                assert np.abs(u1[s * N:(s + 1) * N] - g[f"u1_{variant}"][s * N:(s + 1) * N]).max() <= 1e-12
    uf = u.to_host()
    assert rel_linf(uf[:N], g[f"u300_{variant}"][:N]) <= 1e-6
    assert np.abs(np.array(iters) - g[f"iters_{variant}"]).max() <= 1
    assert np.array_equal(activation_steps(np.array(hist), 0.0), g[f"act_{variant}"])
    for h in (st, u, bS, M, K, md):
        h.free()


def test_fused_and_unfused_paths_agree(tb, dev):
    """The fused tb_monodomain_step and the operator-by-operator path (perform_step! per child) are the
    same algorithm; only the block partition of the first dot product differs (the fused init kernel
    reduces per SELL slice), so states agree to rounding and iteration counts within +-1."""
    g = np.load(GOLD / "c1_small.npz")
    res = []
    for fused in (True, False):
        mesh = tb.generate_mesh(tb.Quadrilateral, (32, 32), (0.0, 0.0), (2.5, 2.5), device=dev)
        odeform = tb.semidiscretize(tb.ReactionDiffusionSplit(_c1_model(tb, g["kappa"])),
                                    tb.FiniteElementDiscretization({"φₘ": tb.LagrangeCollection(1)}), mesh)
        integ = tb.init(tb.OperatorSplittingProblem(odeform, g["u0"].copy(), (0.0, 20.0)),
                        tb.LieTrotterGodunov((tb.BackwardEulerSolver(inner_solver=tb.KrylovJL_CG(atol=1e-6, rtol=1e-5)),
                                              tb.AdaptiveForwardEulerSubstepper())), dt=1.0, fused=fused)
        sol = tb.solve_(integ)
        assert sol.retcode == tb.ReturnCode.Success and integ.t == 20.0
        res.append((integ.u.to_host(), list(integ.cg_iterations)))
    assert np.abs(res[0][0] - res[1][0]).max() <= 1e-9 * np.abs(res[0][0]).max()
    assert np.abs(np.array(res[0][1]) - np.array(res[1][1])).max() <= 1


def test_backward_euler_steady_state(tb, dev):
    """test/test_time_integrator.jl:13-41 through the API: u = 1 is a steady state of pure Neumann diffusion."""
    mesh = tb.generate_mesh(tb.Quadrilateral, (4, 4), (0.0, 0.0), (1.0, 1.0), device=dev)
    f = tb.semidiscretize(tb.TransientDiffusionModel(tb.ConstantCoefficient(tb.SymmetricTensor(2, (1.0, 0.0, 1.0))),
                                                     tb.NoStimulationProtocol(), "u"),
                          tb.FiniteElementDiscretization({"u": tb.LagrangeCollection(1)}), mesh)
    integ = tb.init(tb.ODEProblem(f, np.ones(tb.solution_size(f)), (0.0, 1.0)), tb.BackwardEulerSolver(), dt=0.1)
    assert tb.step_(integ)
    assert np.allclose(integ.u.to_host(), 1.0, atol=1e-4)
    sol = tb.solve_(integ)
    assert np.allclose(integ.u.to_host(), 1.0, atol=1e-4)
    assert sol.retcode == tb.ReturnCode.Success and integ.t == pytest.approx(1.0)


def test_linear_solver_failure_rolls_back(tb, dev):
    """euler.jl:95-100 -> force_stepfail -> rollback (type.jl:510-532) -> ConvergenceFailure."""
    g = np.load(GOLD / "c1_small.npz")
    mesh = tb.generate_mesh(tb.Quadrilateral, (32, 32), (0.0, 0.0), (2.5, 2.5), device=dev)
    odeform = tb.semidiscretize(tb.ReactionDiffusionSplit(_c1_model(tb, g["kappa"])),
                                tb.FiniteElementDiscretization({"φₘ": tb.LagrangeCollection(1)}), mesh)
    integ = tb.init(tb.OperatorSplittingProblem(odeform, g["u0"].copy(), (0.0, 5.0)),
                    tb.LieTrotterGodunov((tb.BackwardEulerSolver(inner_solver=tb.KrylovJL_CG(maxiters=2)),
                                          tb.ForwardEulerCellSolver())), dt=1.0)
    assert not tb.step_(integ)
    assert integ.stats.nreject == 1 and integ.t == 0.0
    assert np.array_equal(integ.u.to_host(), g["u0"])
    assert tb.solve_(integ).retcode == tb.ReturnCode.ConvergenceFailure


def test_random_inputs_one_step_vs_oracle(tb, dev, oracle):
    """Seeded random state on a tet mesh with an anisotropic tensor: one LTG step vs the oracle."""
    O = oracle
    mo = O.generate_grid(O.TET4, (6, 5, 4), (0, 0, 0), (1.5, 1.25, 1.0))
    md = tb.DeviceMesh.from_host(dev, O.TET4, mo.conn, mo.coords, mo.celldofs, mo.ndofs)
    D = np.array([[0.02, 0.004, 0.0], [0.004, 0.01, 0.001], [0.0, 0.001, 0.005]])
    Mo, Ko = O.assemble_mass(mo, 2), O.assemble_diffusion(mo, 2, O.D_TENSOR, D)
    M = tb.B200CSRMatrix.from_mesh(dev, md)
    K = M.like()
    tb.core.assemble_mass(dev, md, M, 2, 1.0)
    tb.core.assemble_diffusion(dev, md, K, 2, 1, D, 1.0)
    rng = np.random.default_rng(42)
    N = mo.ndofs
    u = np.concatenate([rng.uniform(0, 1, N), rng.uniform(0, 0.2, N)])
    bSo = O.assemble_source(mo, 2, O.SRC_BALL, [0.7, 2.0, 0.01], 0.5)
    orc = O.MonodomainOracle(mo, O.FHN, O.default_params(O.FHN), Mo, Ko, substeps=10)
    orc.bS = bSo
    uo = u.copy()
    ito, rno, convo = orc.step(uo, 0.0, 0.5)
    st = tb.MonodomainStepper(dev, M, K, 0, O.default_params(O.FHN))
    st.set_cell_solver(10, 0.1)
    bS = tb.B200Vector(dev, N, 1)
    tb.core.assemble_source(dev, md, bS, 2, tb._lib.SRC_BALL, [0.7, 2.0, 0.01], 0.5)
    st.set_source(bS)
    ud = tb.B200Vector.from_host(dev, u, 2)
    it, rn, conv = st.step(ud, 0.0, 0.5)
    assert conv and abs(it - ito) <= 1
    h = ud.to_host()
    assert rel_linf(h[:N], uo[:N]) <= 1e-10 and rel_linf(h[N:], uo[N:]) <= 1e-10
    # host-buffer (end-to-end) entry point gives the same answer
    out = np.empty_like(u)
    it2, rn2, conv2 = st.step_host(ud, u, out, 0.0, 0.5)
    assert (it2, conv2) == (it, conv) and np.array_equal(out, h)
    # pipelined multi-step variant with the state in (pinned) host memory between steps == device-resident stepping
    import torch
    for nsteps in (1, 4, 5):
        ud.upload(u)
        tot, t = 0, 0.0
        for _ in range(nsteps):
            i_, _r, c_ = st.step(ud, t, 0.5)
            tot += i_
            t += 0.5
        want = ud.to_host()
        b0 = torch.empty(u.size, dtype=torch.float64, pin_memory=True).numpy()
        b1 = torch.empty(u.size, dtype=torch.float64, pin_memory=True).numpy()
        b0[:] = u
        b1[:] = np.nan
        st.set_host_chunks({1: 16, 4: 3, 5: 64}[nsteps])         # pieces of the phi round trip: any count, same bits
        tot2, conv3 = st.run_host(ud, b0, b1, 0.0, 0.5, nsteps)
        res = b1 if nsteps % 2 else b0
        assert conv3 and tot2 == tot and np.array_equal(res, want) and np.array_equal(ud.to_host(), want)
    with pytest.raises(tb.TBError):
        st.set_host_chunks(65)
    for x in (st, ud, bS, M, K, md):
        x.free()


@pytest.mark.parametrize("fused", [True, False])
def test_aliev_panfilov_monodomain_vs_oracle(tb, dev, oracle, fused):
    """Aliev-Panfilov through the reference-shaped API: phi_m is the SECOND state column, so the heat solve reads and
    writes column 1 (heat_dofrange = N+1:2N, fem.jl:399-402) while the cell sweep owns both."""
    O = oracle
    nel = (24, 24)
    mesh = tb.generate_mesh(tb.Quadrilateral, nel, (0.0, 0.0), (2.5, 2.5), device=dev)
    ion = tb.AlievPanfilovModel()
    assert tb.state_symbols(ion) == ("s", "φₘ") and tb.transmembranepotential_index(ion) == 2
    model = tb.MonodomainModel(tb.ConstantCoefficient(1.0), tb.ConstantCoefficient(1.0),
                               tb.ConstantCoefficient(tb.SymmetricTensor(2, (4.5e-3, 0, 2.0e-3))),
                               tb.NoStimulationProtocol(), ion, "φₘ", "s")
    odeform = tb.semidiscretize(tb.ReactionDiffusionSplit(model), tb.FiniteElementDiscretization({"φₘ": tb.LagrangeCollection(1)}), mesh)
    u0 = tb.create_initial_condition(odeform)
    tb.setvariable_(u0, odeform, "φₘ", lambda x: 1.0 if (x[0] <= 1.25 and x[1] <= 1.25) else 0.0)
    tb.setvariable_(u0, odeform, "s", lambda x: 0.5 if x[1] >= 1.25 else 0.0)
    N = mesh.ndofs
    assert u0[:N].max() == 0.5 and u0[N:].max() == 1.0                             # layout: s block, then phi block
    integ = tb.init(tb.OperatorSplittingProblem(odeform, u0.copy(), (0.0, 60.0)),
                    tb.LieTrotterGodunov((tb.BackwardEulerSolver(), tb.AdaptiveForwardEulerSubstepper())), dt=0.5, fused=fused)
    mo = O.generate_grid(O.QUAD4, nel, (0.0, 0.0), (2.5, 2.5))
    orc = O.MonodomainOracle(mo, O.ALIEV_PANFILOV, O.default_params(O.ALIEV_PANFILOV), O.assemble_mass(mo, 2),
                             O.assemble_diffusion(mo, 2, O.D_TENSOR, [4.5e-3, 0, 0, 2.0e-3]), phi_idx=1, substeps=10)
    uo = u0.copy()
    for step in range(120):
        ito, rno, convo = orc.step(uo, 0.5 * step, 0.5)
        assert tb.step_(integ) and convo and abs(integ.cg_iterations[-1] - ito) <= 1
        if step == 0:
            assert rel_linf(integ.u.to_host(), uo) <= 1e-10
    h = integ.u.to_host()
    assert rel_linf(h[N:], uo[N:]) <= 1e-6 and rel_linf(h[:N], uo[:N]) <= 1e-6
    assert np.abs(uo - u0).max() > 0.1                                             # something happened


def test_c2_full_size_40_steps_against_golden(tb, dev):
    """BASELINE config 2 at its FULL size (128x128x32 hexes, 549 153 dofs, PCG2019, adaptive substepper, corner stimulus)
    against the oracle's golden: 1e-10 after one step, 1e-8 after 40, identical CG iteration counts."""
    g = np.load(GOLD / "c2_full.npz")
    md = tb.generate_mesh(tb.Hexahedron, (128, 128, 32), (0, 0, 0), (32.0, 32.0, 8.0), device=dev)
    k1, kr = 0.17 * 0.62 / (0.17 + 0.62), 0.019 * 0.24 / (0.019 + 0.24)
    proto = tb.AnalyticalTransmembraneStimulationProtocol(
        tb.AnalyticalCoefficient(tb.BoxStimulus(1.5, 2.0, 0.5), tb.CartesianCoordinateSystem()), [(0.0, 2.1)])
    model = tb.MonodomainModel(tb.ConstantCoefficient(1.0), tb.ConstantCoefficient(1.0),
                               tb.ConstantCoefficient(np.diag([k1, kr, kr])), proto, tb.PCG2019(), "φₘ", "s")
    odeform = tb.semidiscretize(tb.ReactionDiffusionSplit(model), tb.FiniteElementDiscretization({"φₘ": tb.LagrangeCollection(1)}), md)
    u0 = tb.create_initial_condition(odeform)
    N = md.ndofs
    assert N == 549153
    integ = tb.init(tb.OperatorSplittingProblem(odeform, u0, (0.0, 10.0)),
                    tb.LieTrotterGodunov((tb.BackwardEulerSolver(), tb.AdaptiveForwardEulerSubstepper())), dt=0.01)
    for step in range(40):
        assert tb.step_(integ)
        if step == 0:
            assert rel_linf(integ.u.column(0)[::61], g["phi1"]) <= 1e-10
    assert rel_linf(integ.u.column(0)[::61], g["phi40"]) <= 1e-8
    assert rel_linf(integ.u.column(1)[::61], g["h40"]) <= 1e-8
    assert np.abs(np.array(integ.cg_iterations) - g["iters"]).max() <= 1
    assert (integ.u.column(0) > -84.0).sum() == int(g["n_above"][0])


def test_c2_full_size_1000_steps_against_golden(tb, dev):
    """BASELINE config 2 at its FULL size AND full length (1000 steps of dt = 0.01): 1e-10 after one step,
    1e-6 after 1000, CG iterations +-1 on every step, activation steps identical (sampled at every 61st dof)."""
    g = np.load(GOLD / "c2_full_1000.npz")
    md = tb.generate_mesh(tb.Hexahedron, (128, 128, 32), (0, 0, 0), (32.0, 32.0, 8.0), device=dev)
    k1, kr = 0.17 * 0.62 / (0.17 + 0.62), 0.019 * 0.24 / (0.019 + 0.24)
    proto = tb.AnalyticalTransmembraneStimulationProtocol(
        tb.AnalyticalCoefficient(tb.BoxStimulus(1.5, 2.0, 0.5), tb.CartesianCoordinateSystem()), [(0.0, 2.1)])
    model = tb.MonodomainModel(tb.ConstantCoefficient(1.0), tb.ConstantCoefficient(1.0),
                               tb.ConstantCoefficient(np.diag([k1, kr, kr])), proto, tb.PCG2019(), "φₘ", "s")
    odeform = tb.semidiscretize(tb.ReactionDiffusionSplit(model), tb.FiniteElementDiscretization({"φₘ": tb.LagrangeCollection(1)}), md)
    u0 = tb.create_initial_condition(odeform)
    integ = tb.init(tb.OperatorSplittingProblem(odeform, u0, (0.0, 10.0)),
                    tb.LieTrotterGodunov((tb.BackwardEulerSolver(), tb.AdaptiveForwardEulerSubstepper())), dt=0.01)
    act = np.full(g["act"].shape, -1, dtype=np.int16)
    for step in range(1000):
        assert tb.step_(integ)
        if step == 0:
            assert rel_linf(integ.u.column(0)[::61], g["phi1"]) <= 1e-10
        ph = integ.u.column(0)[::61]
        act[(act < 0) & (ph >= 0.0)] = step + 1
    assert rel_linf(integ.u.column(0)[::61], g["phi1000"]) <= 1e-6
    assert rel_linf(integ.u.column(1)[::61], g["h1000"]) <= 1e-6
    assert np.abs(np.array(integ.cg_iterations) - g["iters"]).max() <= 1
    assert np.array_equal(act, g["act"]) and (act > 0).sum() > 50                # activation steps identical at time-step resolution
    assert (integ.u.column(0) > -84.0).sum() == int(g["n_above"][0])


def test_c2_full_size_properties(tb, dev):
    """BASELINE config 2 at FULL size (128x128x32, PCG2019): properties that need no oracle.
    Pure-Neumann diffusion conserves 1^T M phi; the CG residual really is below tolerance; resting
    tissue stays at rest; the stimulated corner activates and the wave moves away from it."""
    md = tb.generate_mesh(tb.Hexahedron, (128, 128, 32), (0, 0, 0), (32.0, 32.0, 8.0), device=dev)
    k1, kr = 0.17 * 0.62 / (0.17 + 0.62), 0.019 * 0.24 / (0.019 + 0.24)
    M = tb.B200CSRMatrix.from_mesh(dev, md)
    K = M.like()
    tb.core.assemble_mass(dev, md, M, 2, 1.0)
    tb.core.assemble_diffusion(dev, md, K, 2, 1, np.diag([k1, kr, kr]), 1.0)
    N = md.ndofs
    ion = tb.PCG2019()
    x = md.dof_coords()
    # (1) BE + CG alone: mass conservation and true residual
    A = M.like()
    A.axpby_values(M, K, 0.01)
    phi = tb.B200Vector.from_host(dev, -85.0 + 100.0 * np.exp(-((x - [16, 16, 4]) ** 2).sum(1) / 8.0))
    b, y, xs = tb.B200Vector(dev, N), tb.B200Vector(dev, N), tb.B200Vector(dev, N)
    M.mul(b, phi)
    it, rn, conv = tb.core.cg_solve(dev, A, b, xs)
    assert conv and 10 < it < 200
    A.mul(y, xs)
    r = y.to_host() - b.to_host()
    assert np.linalg.norm(r) <= 1.01 * (tb.SQRT_EPS + tb.SQRT_EPS * np.linalg.norm(b.to_host())) + 1e-12
    M.mul(y, xs)
    assert y.to_host().sum() == pytest.approx(b.to_host().sum(), rel=1e-9)
    # (2) full LTG steps with the corner stimulus
    st = tb.MonodomainStepper(dev, M, K, ion.model_id, ion.params())
    u = tb.B200Vector.from_host(dev, np.repeat(tb.default_initial_state(ion), N), 7)
    bS = tb.B200Vector(dev, N, 1)
    t, dt = 0.0, 0.01
    for step in range(150):
        tb.core.assemble_source(dev, md, bS, 2, tb._lib.SRC_BOX, [1.5, 2.0, 0.5], t + dt)
        st.set_source(bS)
        it, rn, conv = st.step(u, t, dt)
        assert conv
        t += dt
    ph = u.column(0)
    corner = np.all(x < 1.0, axis=1)
    far = (x[:, 0] > 16.0) & (x[:, 1] > 16.0)
    assert ph[corner].min() > -40.0                      # stimulated corner has depolarised
    assert np.abs(ph[far] + 85.0).max() < 0.1            # far field still at rest
    assert np.isfinite(u.to_host()).all()
    for h in (st, u, bS, A, phi, b, y, xs, M, K, md):
        h.free()


@pytest.mark.parametrize("mode,nel", [(1, (48, 40)), (2, (48, 40)), (0, (48, 40)), (1, (40, 36, 30))])
def test_run_without_readbacks_matches_stepping(tb, dev, mode, nel):
    """tb_monodomain_run: with a persistent CG kernel (mode 1: register-resident, mode 2: TMA sweep) the whole run is enqueued
    without a read-back, iteration counts and convergence flags folded on the device (k_pcg_fold); with the multi-kernel path
    (mode 0) it is the per-step loop.  Either way: the state and the total iteration count of stepping one by one, to the bit."""
    ct = tb.Quadrilateral if len(nel) == 2 else tb.Hexahedron
    md = tb.generate_mesh(ct, nel, (0.0,) * len(nel), tuple(0.25 * n for n in nel), device=dev)
    M = tb.B200CSRMatrix.from_mesh(dev, md)
    K = M.like()
    tb.core.assemble_mass(dev, md, M, 2, 1.0)
    tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_TENSOR, np.diag([0.13, 0.02, 0.02][:len(nel)]), 1.0)
    ion = tb.ParametrizedFHNModel()
    st = tb.MonodomainStepper(dev, M, K, ion.model_id, ion.params())
    n = md.ndofs
    rng = np.random.default_rng(8)
    u0 = np.concatenate([rng.uniform(0, 1, n), 0.1 * rng.uniform(0, 1, n)])
    dev.cg_set_persistent(mode)
    try:
        for precond in (tb._lib.PRECOND_NONE, tb._lib.PRECOND_JACOBI):
            st.set_preconditioner(precond)
            ua, ub = tb.B200Vector.from_host(dev, u0, 2), tb.B200Vector.from_host(dev, u0, 2)
            tot, t = 0, 0.0
            for _ in range(7):
                it, _rn, conv = st.step(ua, t, 0.7)
                assert conv and dev.cg_last_path() == mode
                tot += it
                t += 0.7
            tot2, conv2 = st.run(ub, 0.0, 0.7, 7)
            assert conv2 and tot2 == tot and np.array_equal(ua.to_host(), ub.to_host())
            it, _rn, conv = st.step(ub, t, 0.7)                       # and the per-step entry point still reports afterwards
            assert conv and it > 0
            ua.free(); ub.free()
    finally:
        dev.cg_set_persistent(1)
    for h in (st, M, K, md):
        h.free()
