"""Host-side logic of the API mirror that needs no GPU: numbering, index sets, defaults, gating."""
import numpy as np
import pytest


def test_close_dofs_matches_oracle(tb, oracle):
    O = oracle
    for ct, nel in ((O.QUAD4, (5, 3)), (O.HEX8, (3, 2, 4)), (O.TET4, (2, 3, 2)), (O.TRI3, (4, 4))):
        m = O.generate_grid(ct, nel, (0,) * len(nel), (1,) * len(nel))
        celldofs, ndofs = tb.api.close_dofs(m.conn)
        assert ndofs == m.ndofs and np.array_equal(celldofs, m.celldofs)


def test_symmetric_tensor_helper(tb):
    T = tb.SymmetricTensor(2, (4.5e-5, 0, 2.0e-5))
    assert np.array_equal(T, [[4.5e-5, 0], [0, 2.0e-5]])
    T = tb.SymmetricTensor(3, (1, 2, 3, 4, 5, 6))
    assert np.array_equal(T, [[1, 2, 3], [2, 4, 5], [3, 5, 6]])


def test_default_initial_states(tb, oracle):
    assert np.array_equal(tb.default_initial_state(tb.FHNModel()), [0.0, 0.0])
    assert np.allclose(tb.default_initial_state(tb.PCG2019()), oracle.default_initial_state(oracle.PCG2019), rtol=1e-15)
    assert np.array_equal(tb.PCG2019().params(), oracle.default_params(oracle.PCG2019))
    assert np.array_equal(tb.FHNModel().params(), oracle.default_params(oracle.FHN))
    assert tb.num_states(tb.PCG2019()) == 7 and tb.state_symbols(tb.PCG2019())[0] == "φₘ"
    with pytest.raises(TypeError):
        tb.PCG2019(g_Nope=1.0)


class _FakeMesh:
    def __init__(self, n):
        self.ndofs, self.dim = n, 2


def test_semidiscretize_index_sets(tb):
    """test/test_solution_variables.jl:76-88: phi index set = 1:N and equals solution_indices[1]."""
    model = tb.MonodomainModel(tb.ConstantCoefficient(1.0), tb.ConstantCoefficient(1.0),
                               tb.ConstantCoefficient(tb.SymmetricTensor(2, (1.0, 0, 1.0))), tb.NoStimulationProtocol(),
                               tb.FHNModel(), "φₘ", "s")
    f = tb.semidiscretize(tb.ReactionDiffusionSplit(model),
                          tb.FiniteElementDiscretization({"φₘ": tb.LagrangeCollection(1)}), _FakeMesh(25))
    assert np.array_equal(f.solution_indices[0], np.arange(1, 26))
    assert f.solution_indices[1] == range(1, 51)
    assert tb.solution_size(f) == 50
    heat, ode = f.functions
    assert heat.mass_term.qrc.order == 2 and heat.bilinear_term.qrc.order == 2    # fem.jl:52-55
    u0 = tb.create_initial_condition(f)
    assert u0.shape == (50,) and not u0.any()
    model.ion = tb.PCG2019()
    f = tb.semidiscretize(tb.ReactionDiffusionSplit(model),
                          tb.FiniteElementDiscretization({"φₘ": tb.LagrangeCollection(1)}), _FakeMesh(4))
    u0 = tb.create_initial_condition(f)
    assert np.array_equal(u0[:4], [-85.0] * 4) and np.allclose(u0[4:8], tb.default_initial_state(tb.PCG2019())[1])
    with pytest.raises(KeyError):
        tb.semidiscretize(tb.ReactionDiffusionSplit(model), tb.FiniteElementDiscretization({"u": tb.LagrangeCollection(1)}),
                          _FakeMesh(4))
    with pytest.raises(NotImplementedError):
        tb.semidiscretize(tb.ReactionDiffusionSplit(model), tb.FiniteElementDiscretization({"φₘ": tb.LagrangeCollection(2)}),
                          _FakeMesh(4))


def test_needs_update_closed_intervals(tb):
    """src/discretization/operator.jl:17-26"""
    proto = tb.AnalyticalTransmembraneStimulationProtocol(tb.AnalyticalCoefficient(tb.BoxStimulus(1.5, 2.0, 0.5)),
                                                          [(0.0, 2.1), (5.0, 6.0)])
    op = tb.api.LinearOperator(tb.api.LinearIntegrator(proto, tb.QuadratureRuleCollection(2)), None, None)
    assert [tb.needs_update(op, t) for t in (-0.1, 0.0, 2.1, 2.2, 5.0, 6.0, 6.01)] == [False, True, True, False, True, True, False]
    assert not tb.needs_update(tb.api.LinearNullOperator(), 1.0)
    assert isinstance(tb.setup_operator(None, tb.api.LinearIntegrator(tb.NoStimulationProtocol(), None), None, _FakeMesh(3)),
                      tb.api.LinearNullOperator)


def test_diffusion_coefficient_lowering(tb):
    class M:
        dim, ncells, nv = 3, 2, 4
    k = tb.SymmetricTensor(3, (1.0, 0, 0, 2.0, 0, 3.0))
    D = tb.api.ConductivityToDiffusivityCoefficient(tb.ConstantCoefficient(k), tb.ConstantCoefficient(2.0), tb.ConstantCoefficient(0.5))
    kind, data, cmchi = tb.api._diffusion_data(D, M)
    assert kind == 1 and cmchi == 1.0 and np.array_equal(data.reshape(3, 3), k)
    ms = tb.OrthotropicMicrostructureModel(tb.ConstantCoefficient((0, 0, 1.0)), tb.ConstantCoefficient((0, 1.0, 0)),
                                           tb.ConstantCoefficient((1.0, 0, 0)))
    kind, data, cmchi = tb.api._diffusion_data(tb.SpectralTensorCoefficient(ms, tb.ConstantCoefficient((0.3, 0.1, 0.1))), M)
    assert kind == 2 and data.size == 3 + 2 * 4 * 9
    assert np.array_equal(data[3:12], [0, 0, 1, 0, 1, 0, 1, 0, 0])
    kind, data, _ = tb.api._diffusion_data(tb.ConstantCoefficient(0.7), M)
    assert kind == 0 and data[0] == 0.7


def test_unsupported_kwargs_are_refused(tb):
    """test/test_time_integrator.jl:43-273 ('unsupported kwargs refused')"""
    with pytest.raises(TypeError):
        tb.init(None, None, dt=0.1, callback=None)


def test_cell_model_tables_and_solver_options(tb):
    """Host-side facts of the API mirror that need no GPU: state layouts of the three ionic models
    (fhn.jl:14-19, pcg2019.jl:136, aliev-panfilov.jl:11-14), the preconditioner switch, the RTC unwrapping."""
    import numpy as np
    assert tb.state_symbols(tb.FHNModel()) == ("φₘ", "s") and tb.transmembranepotential_index(tb.FHNModel()) == 1
    assert tb.num_states(tb.PCG2019()) == 7 and tb.transmembranepotential_index(tb.PCG2019()) == 1
    ap = tb.AlievPanfilovModel()
    assert tb.state_symbols(ap) == ("s", "φₘ") and tb.transmembranepotential_index(ap) == 2 and tb.num_states(ap) == 2
    assert np.array_equal(ap.params(), [1.0 / 12.9, 8.0, 0.05, 0.002, 0.2, 0.3])
    assert np.array_equal(tb.default_initial_state(ap), [0.0, 0.0])
    assert tb.B200CG().precond == tb._lib.PRECOND_NONE
    assert tb.B200CG(precs=tb.JacobiPreconditioner()).precond == tb._lib.PRECOND_JACOBI
    ltg = tb.LieTrotterGodunov((tb.BackwardEulerSolver(), tb.AdaptiveForwardEulerSubstepper()))
    rtc = tb.ReactionTangentController(ltg, 0.5, 1.0, (0.5, 2.0))
    assert rtc.inner_algs is ltg.inner_algs and 0.5 <= rtc.next_dt(3.0) < rtc.next_dt(0.0) <= 2.0
