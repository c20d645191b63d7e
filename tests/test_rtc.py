"""ReactionTangentController (src/solver/time/rtc.jl): the sigma(R) map against the reference's own
known answers (test/test_os_gearing.jl:250-296), and -- on the GPU -- the adaptive integrator against the oracle's
restatement and against the reference's integration test (test/integration/test_electrophysiology.jl:66-100)."""
import math

import numpy as np
import pytest


def _expected(R, s, c, lo, hi):
    return (1 - 1 / (1 + math.exp((c - R) * s))) * (hi - lo) + lo


@pytest.mark.parametrize("bounds", [(0.5, 2.0), (0.01, 0.1)])
def test_sigma_known_answers(oracle, bounds):
    import thunderbolt_jl_b200 as tb
    lo, hi = bounds
    ltg = tb.LieTrotterGodunov((tb.BackwardEulerSolver(), tb.ForwardEulerCellSolver()))
    # "Sigmoid formula at R = 0.5" (test_os_gearing.jl:253-270)
    rtc = tb.ReactionTangentController(ltg, 0.5, 1.0, bounds)
    assert rtc.inner_algs is ltg.inner_algs                               # LTG unwrapped (rtc.jl:37-38)
    assert rtc.next_dt(0.5) == pytest.approx(_expected(0.5, 0.5, 1.0, lo, hi), rel=1e-15)
    assert oracle.rtc_next_dt(0.5, 0.5, 1.0, bounds) == pytest.approx(rtc.next_dt(0.5), rel=1e-15)
    # "sigma_s = Inf": step function, the boundary R == sigma_c goes to dt_max (test_os_gearing.jl:275-296)
    for c, want in ((0.75, hi), (0.5, hi), (0.25, lo)):
        assert tb.ReactionTangentController(ltg, math.inf, c, bounds).next_dt(0.5) == want
        assert oracle.rtc_next_dt(0.5, math.inf, c, bounds) == want
    # monotone: larger tangents give smaller steps, always inside the bounds
    Rs = np.linspace(-5, 20, 101)
    d = np.array([rtc.next_dt(r) for r in Rs])
    assert np.all(np.diff(d) <= 0) and d.min() >= lo and d.max() <= hi


def _problem(tb, dev, O, nel=(8, 8)):
    """test_electrophysiology.jl:72-92 ("Single subdomain"): 8x8 quads on [-2.5,2.5]^2, FHN, apex stimulus."""
    mesh = tb.generate_mesh(tb.Quadrilateral, nel, (-2.5, -2.5), (2.5, 2.5), device=dev)
    proto = tb.AnalyticalTransmembraneStimulationProtocol(
        tb.AnalyticalCoefficient(tb.BallStimulus(0.1, 2.0, 0.01), tb.CartesianCoordinateSystem()), [(0.0, 2.1)])
    model = tb.MonodomainModel(tb.ConstantCoefficient(1.0), tb.ConstantCoefficient(1.0),
                               tb.ConstantCoefficient(np.array([[4.5e-4, 0.0], [0.0, 2.0e-4]])), proto, tb.FHNModel(), "φₘ", "s1")
    odeform = tb.semidiscretize(tb.ReactionDiffusionSplit(model),
                                tb.FiniteElementDiscretization({"φₘ": tb.LagrangeCollection(1)}), mesh)
    return mesh, odeform


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
def test_rtc_integrator_vs_oracle_and_fixed_dt(tb, dev, oracle, fused):
    O = oracle
    mesh, odeform = _problem(tb, dev, O)
    N = mesh.ndofs
    rng = np.random.default_rng(0)
    # a travelling front so that the tangent really moves: phi = 1 on the left third, recovery variable 0
    x = mesh.dof_coords()
    u0 = np.concatenate([np.where(x[:, 0] < -0.8, 1.0, 0.0) + 0.01 * rng.standard_normal(N), np.zeros(N)])
    ltg = tb.LieTrotterGodunov((tb.BackwardEulerSolver(), tb.ForwardEulerCellSolver()))
    rtc = tb.ReactionTangentController(ltg, 0.5, 1.0, (0.5, 2.0))
    tspan = (0.0, 50.0)
    integ = tb.init(tb.OperatorSplittingProblem(odeform, u0.copy(), tspan), ltg, dt=1.0, fused=fused)
    tb.solve_(integ)
    integ_rtc = tb.init(tb.OperatorSplittingProblem(odeform, u0.copy(), tspan), rtc, dt=1.0, fused=fused)
    tb.solve_(integ_rtc)
    assert integ_rtc.sol.retcode == tb.ReturnCode.Success and integ_rtc.t == tspan[1]
    a, b = integ.u.to_host(), integ_rtc.u.to_host()
    # same solution up to the first-order time error of dt in [0.5, 2] vs dt = 1 on a travelling front (the reference's
    # own scenario, a 0.01 stimulus, passes rtol = 1e-2; this one is deliberately more dynamic so that R really moves)
    assert np.linalg.norm(a - b) <= 0.15 * max(np.linalg.norm(a), np.linalg.norm(b))
    assert integ_rtc.stats.naccept != integ.stats.naccept                                     # dt moved away from 1.0
    assert min(integ_rtc.dts[1:-1]) >= 0.5 and max(integ_rtc.dts) <= 2.0
    # oracle: same controller around the oracle's LTG step
    mo = O.generate_grid(O.QUAD4, (8, 8), (-2.5, -2.5), (2.5, 2.5))
    D = np.array([[4.5e-4, 0.0], [0.0, 2.0e-4]])
    orc = O.MonodomainOracle(mo, O.FHN, O.default_params(O.FHN), O.assemble_mass(mo, 2), O.assemble_diffusion(mo, 2, O.D_TENSOR, D))
    uo, t, dt, dts = u0.copy(), 0.0, 1.0, []
    while t < tspan[1] * (1 - 1e-15):
        step = min(dt, tspan[1] - t) if tspan[1] - t < dt * (1 - 1e-12) else dt
        # stimulus only re-assembled inside [0, 2.1]; afterwards the last vector keeps being added (euler.jl:88-91)
        if 0.0 <= t + step <= 2.1:
            orc.bS = O.assemble_source(mo, 2, O.SRC_BALL, [0.1, 2.0, 0.01], t + step)
        elif orc.bS is None:
            orc.bS = np.zeros(N)
        it, rn, conv = orc.step(uo, t, step)
        assert conv
        t += step
        dts.append(step)
        dt = O.rtc_next_dt(orc.reaction_tangent(), 0.5, 1.0, (0.5, 2.0))
    assert len(dts) == len(integ_rtc.dts)
    assert np.allclose(dts, integ_rtc.dts, rtol=1e-9, atol=0)
    assert np.abs(b[:N] - uo[:N]).max() <= 1e-8 * np.abs(uo[:N]).max()
