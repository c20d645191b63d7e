"""Traced stimulus closures (trace.py -> tb_assemble_source_program): the closure of an AnalyticalCoefficient
(analytical_coefficient.jl:80-101 evaluates it at every quadrature point) crosses the C ABI as a postfix program.

CPU: the tracer, the Python interpreter of the program and the product's own evaluator (tb_program_eval compiled for the
host) against the closures themselves -- the reference's stimulus closures among them.  GPU: the assembled vector against
the host-evaluated path (tb_assemble_source_qp) and against the built-in families."""
import math

import numpy as np
import pytest

import thunderbolt_jl_b200 as tb
from thunderbolt_jl_b200 import trace as T

# the reference's own stimulus closures, written branch-free
CLOSURES = {
    # bak/examples/conduction-velocity-benchmark.jl:47-50: maximum(x) < 1.5 && t < 2.0 ? 0.5 : 0.0
    "cv_box": lambda x, t: T.where((T.maximum(T.maximum(x[0], x[1]), x[2]) < 1.5) & (t < 2.0), 0.5, 0.0),
    # test/integration/test_electrophysiology.jl:83: norm(x) < 0.25 && t < 2.0 ? 0.5 : 0.0
    "ball": lambda x, t: T.where((T.norm(x) < 0.25) & (t < 2.0), 0.5, 0.0),
    # benchmarks/benchmarks-cuda-linear-form.jl:4-18: cos(2 pi t) exp(-|x|^2)
    "cosexp": lambda x, t: np.cos(2.0 * math.pi * t) * np.exp(-(x[0] * x[0] + x[1] * x[1] + x[2] * x[2])),
    # benchmarks/benchmarks-linear-form.jl:16-21: norm(x) + t
    "normt": lambda x, t: T.norm(x) + t,
    # ep04_geselowitz-ecg.jl:15-26: t <= 2 && x[0] < 0.15 ? 0.5 / 0.25 * exp(t / 0.25) : 0
    "endo": lambda x, t: T.where((t <= 2.0) & (x[0] < 0.15), 0.5 / 0.25 * T.exp(t / 0.25), 0.0),
    "poly": lambda x, t: (x[0] - 0.3) ** 2 * x[1] - x[2] ** 3 / (1.0 + t) + abs(x[0] - x[1]),
    "mixed": lambda x, t: T.minimum(x[0] * t, 2.0) + np.tanh(x[1]) - np.sqrt(abs(x[2]) + 1.0) + (x[0] > x[1]) * 0.25,
    # closures written with plain numpy: ufuncs on tracing numbers dispatch through __array_ufunc__, reductions over the
    # coordinate vector (np.linalg.norm, x @ x) through the object loops
    "numpy_minmax": lambda x, t: np.maximum(x[0], 1.0) + np.minimum(x[1], t) * np.where(False, 0.0, 1.0),
    "numpy_norm": lambda x, t: np.linalg.norm(x) + t,
    "numpy_dot": lambda x, t: np.exp(-(x @ x)) * np.cos(2 * np.pi * t),
    "numpy_logic": lambda x, t: (np.less(x[0], 0.5) & np.greater_equal(t, 0.1)) * 0.5 + np.power(x[1], 2) + np.square(x[2]),
}
EXACT = {"cv_box", "ball", "normt"}     # +,-,*,/,sqrt,abs,min,max,compare only: bitwise on every side (poly: Python's
                                        # float ** 2 is libm pow, the trace lowers it to x*x like Julia's literal_pow)


def _pts(n=300, dim=3):
    rng = np.random.default_rng(11)
    x = rng.uniform(-2.0, 2.0, (n, dim))
    x[:5] = [[1.5, 0.0, 0.0][:dim], [0.25, 0.0, 0.0][:dim], [0.0] * dim, [0.15, 1.0, 1.0][:dim], [1.4999999999999998, 1.5, 0][:dim]]
    return x


@pytest.mark.parametrize("name", list(CLOSURES))
def test_trace_reproduces_closure(name, hostmath):
    f = CLOSURES[name]
    prog = T.trace_source(f, 3)
    assert prog is not None and 0 < len(prog) <= T.MAXCODE
    x = _pts()
    for t in (0.0, 0.3, 1.999, 2.0, 2.5):
        want = np.array([float(f(xi, t)) for xi in x])
        got_py = np.array([prog(xi, t) for xi in x])
        got_c = np.zeros(len(x))
        rc = hostmath.hm_program_eval(3, prog.code, len(prog), np.ascontiguousarray(prog.consts) if prog.consts.size else np.zeros(1),
                                      prog.consts.size, np.ascontiguousarray(x), len(x), t, got_c)
        assert rc == 0
        if name in EXACT:
            assert np.array_equal(got_py, want) and np.array_equal(got_c, want)
        else:
            assert np.allclose(got_py, want, rtol=1e-12, atol=1e-14) and np.allclose(got_c, want, rtol=1e-12, atol=1e-14)


def test_untraceable_closures_fall_back():
    assert T.trace_source(lambda x, t: 0.5 if (max(x) < 1.5 and t < 2.0) else 0.0, 3) is None       # Python branch
    assert T.trace_source(lambda x, t: math.exp(-x[0]), 3) is None                                  # math.* wants a float
    state = {"n": 0}

    def counting(x, t):
        state["n"] += 1
        return x[0] * state["n"]
    assert T.trace_source(counting, 3) is None                                                      # not a function of (x, t)
    deep = lambda x, t: sum((x[0] + float(k)) * (x[1] - float(k)) for k in range(40))              # > 96 instructions
    assert T.trace_source(deep, 3) is None
    assert T.trace_source(lambda x, t: x[0] + 1.0, 2) is not None


def test_program_validation(hostmath):
    out, x, c = np.zeros(1), np.zeros(3), np.zeros(1)
    bad = {-2: [99], -3: [T.OP["X"] | (3 << 8)], -4: [T.OP["ADD"]], -5: [T.OP["T"], T.OP["T"]], -1: [T.OP["T"]] * 97}
    for rc, code in bad.items():
        assert hostmath.hm_program_eval(3, np.array(code, dtype=np.int32), len(code), c, 0, x, 1, 0.0, out) == rc
    assert hostmath.hm_program_eval(3, np.array([T.OP["CONST"]], dtype=np.int32), 1, c, 0, x, 1, 0.0, out) == -3


@pytest.mark.gpu
@pytest.mark.parametrize("ct,nel", [(tb.Hexahedron, (9, 7, 5)), (tb.Tetrahedron, (5, 4, 3)), (tb.Quadrilateral, (17, 13)),
                                    (tb.Triangle, (11, 9))])
def test_program_matches_host_evaluated_path(dev, ct, nel):
    dim = len(nel)
    mesh = tb.generate_mesh(ct, nel, (-1.0,) * dim, (2.0, 1.7, 1.2)[:dim], device=dev)
    conn, coords, _ = mesh.download()
    pts, _w = tb.core.quadrature(mesh.celltype, 2)
    xq = np.einsum("qa,cad->cqd", tb.api._shape_values(mesh.celltype, pts), coords[conn])
    b1, b2 = tb.B200Vector(dev, mesh.ndofs, 1), tb.B200Vector(dev, mesh.ndofs, 1)
    for name, f3 in CLOSURES.items():
        f = f3 if dim == 3 else (lambda x, t, f3=f3: f3(np.array([x[0], x[1], 0.5 * x[0]], dtype=object if isinstance(x[0], T.Sym) else float), t))
        prog = T.trace_source(f, dim)
        assert prog is not None
        for t in (0.3, 2.5):
            tb.core.assemble_source_program(dev, mesh, b1, 2, prog.code, prog.consts, t)
            fq = np.array([[float(f(xq[c, k], t)) for k in range(xq.shape[1])] for c in range(xq.shape[0])])
            tb.core.assemble_source_qp(dev, mesh, b2, 2, fq)
            got, want = b1.to_host(), b2.to_host()
            if name in EXACT:
                # same quadrature points? the device forms x_q = sum_a N_a x_a itself; the host copy does it with numpy
                assert np.allclose(got, want, rtol=1e-13, atol=1e-15 * np.abs(want).max() if np.abs(want).max() > 0 else 0.0)
            else:
                assert np.allclose(got, want, rtol=1e-12, atol=1e-14 * max(np.abs(want).max(), 1e-300))
    # the built-in family and the traced reference closure are the same function: bitwise equal vectors
    if dim == 3:
        tb.core.assemble_source(dev, mesh, b2, 2, tb._lib.SRC_BOX, [1.5, 2.0, 0.5], 0.3)
        prog = T.trace_source(CLOSURES["cv_box"], 3)
        tb.core.assemble_source_program(dev, mesh, b1, 2, prog.code, prog.consts, 0.3)
        assert np.array_equal(b1.to_host(), b2.to_host())
        tb.core.assemble_source(dev, mesh, b2, 2, tb._lib.SRC_NORMT, [0.0], 0.3)
        prog = T.trace_source(CLOSURES["normt"], 3)
        tb.core.assemble_source_program(dev, mesh, b1, 2, prog.code, prog.consts, 0.3)
        assert np.array_equal(b1.to_host(), b2.to_host())
    for h in (b1, b2, mesh):
        h.free()


@pytest.mark.gpu
def test_integrator_traces_closure_stimulus(dev):
    """update_operator! on a closure stimulus: traced and host-evaluated paths give the same trajectory"""
    res = {}
    for traced in (True, False):
        tb.api.TRACE_CLOSURES = traced
        try:
            mesh = tb.generate_mesh(tb.Hexahedron, (8, 8, 4), (0, 0, 0), (2.0, 2.0, 1.0), device=dev)
            f = tb.AnalyticalCoefficient(lambda x, t: T.where((T.norm(x) < 0.8) & (t < 0.5), 0.5, 0.0), tb.CartesianCoordinateSystem(mesh))
            proto = tb.AnalyticalTransmembraneStimulationProtocol(f, [(0.0, 0.6)])
            model = tb.MonodomainModel(tb.ConstantCoefficient(1.0), tb.ConstantCoefficient(1.0),
                                       tb.ConstantCoefficient(tb.SymmetricTensor(3, [0.1, 0, 0, 0.05, 0, 0.05])), proto,
                                       tb.ParametrizedFHNModel())
            odeform = tb.semidiscretize(tb.ReactionDiffusionSplit(model),
                                        tb.FiniteElementDiscretization({"φₘ": tb.LagrangeCollection(1)}), mesh)
            u0 = tb.create_initial_condition(odeform)
            prob = tb.OperatorSplittingProblem(odeform, u0, (0.0, 1.0))
            integ = tb.init(prob, tb.LieTrotterGodunov((tb.BackwardEulerSolver(), tb.ForwardEulerCellSolver())), dt=0.1)
            for _ in range(8):
                assert tb.step_(integ)
            res[traced] = integ.u.to_host().copy()
        finally:
            tb.api.TRACE_CLOSURES = True
    assert np.abs(res[True]).max() > 1e-3
    assert np.allclose(res[True], res[False], rtol=1e-11, atol=1e-14)
