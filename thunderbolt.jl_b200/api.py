"""Host-side mirror of Thunderbolt.jl's public API for the monodomain path.

Names, argument meaning and error behaviour follow the reference so that tests read like the
reference's own (`!` becomes a trailing underscore: `step!` -> `step_`).  Everything numerical is a
call into libtbolt_b200.so; this module only carries the plumbing the reference does in Julia:

  models / protocols            src/modeling/electrophysiology.jl:240-381, cells/fhn.jl, cells/pcg2019.jl
  coefficients                  src/modeling/core/coefficients.jl, microstructure.jl
  semidiscretize + index sets   src/discretization/fem.jl:170-196,371-419; solution_variables.jl:53-68
  solver structs + caches       src/solver/time/euler.jl:4-179, partitioned_solver.jl:57-269
  operators                     src/solver/interface.jl:17-94, src/discretization/operator.jl:2-32
  integrator                    src/solver/time/integrator/type.jl:79-498, operatorsplitting-interface.jl:23-232
"""
from __future__ import annotations

import enum
import math
from dataclasses import dataclass, field
from typing import Callable, Sequence

import numpy as np

from . import _lib as L
from . import core
from . import trace
from .core import B200CSRMatrix, B200Device, B200Vector, DeviceMesh, SQRT_EPS

TRACE_CLOSURES = True      # False: every closure takes the host-evaluated path (tests compare the two)

# ---------------------------------------------------------------------------------------------
# cells
# ---------------------------------------------------------------------------------------------
Quadrilateral, Hexahedron, Triangle, Tetrahedron = L.QUAD4, L.HEX8, L.TRI3, L.TET4


@dataclass
class ParametrizedFHNModel:
    """src/modeling/cells/fhn.jl:6-13"""
    a: float = 0.1
    b: float = 0.5
    c: float = 1.0
    d: float = 0.0
    e: float = 0.01
    f: float = 1.0
    model_id = L.FHN

    def params(self):
        return np.array([self.a, self.b, self.c, self.d, self.e, self.f])


FHNModel = ParametrizedFHNModel


@dataclass
class ParametrizedAlievPanfilovModel:
    """src/modeling/cells/aliev-panfilov.jl:1-34.  The recovery variable comes first: state_symbols = (s, φₘ), so the
    transmembrane potential is state 2 (1-based) -- the one model on the path with φₘ not in front."""
    c_t: float = 1.0 / 12.9     # cₜ
    k: float = 8.0
    a: float = 0.05
    eps0: float = 0.002         # ϵ₀
    mu1: float = 0.2            # μ₁
    mu2: float = 0.3            # μ₂
    model_id = L.ALIEV_PANFILOV

    def params(self):
        return np.array([self.c_t, self.k, self.a, self.eps0, self.mu1, self.mu2])


AlievPanfilovModel = ParametrizedAlievPanfilovModel

_PCG_FIELDS = ("g_Na E_m k_m tau_m E_h k_h delta_h tau_h0 g_K1 E_z k_z g_to E_r k_r E_s k_s tau_s g_CaL E_d k_d E_f k_f "
               "tau_f g_Kr E_xr k_xr tau_xr E_y k_y g_Ks E_xs k_xs tau_xs E_Na E_K E_Ca").split()
_PCG_DEFAULTS = (12.0, -52.244, 6.5472, 0.12, -78.7, 5.93, 0.799163, 6.80738, 0.73893, -91.9655, 12.4997, 0.1688, 14.3116,
                 11.462, -47.9286, 4.9314, 9.90669, 0.11503, 0.7, 4.3, -15.7, 4.6, 30.0, 0.056, -26.6, 6.5, 334.0, -49.6,
                 23.5, 0.008, 24.6, 12.1, 628.0, 65.0, -85.0, 50.0)


class ParametrizedPCG2019Model:
    """src/modeling/cells/pcg2019.jl:4-48 (36 parameters, declaration order)."""
    model_id = L.PCG2019

    def __init__(self, **kw):
        vals = dict(zip(_PCG_FIELDS, _PCG_DEFAULTS))
        for k, v in kw.items():
            if k not in vals:
                raise TypeError(f"unknown PCG2019 parameter {k}")
            vals[k] = float(v)
        self.__dict__.update(vals)

    def params(self):
        return np.array([getattr(self, k) for k in _PCG_FIELDS])


PCG2019 = ParametrizedPCG2019Model


def num_states(ion) -> int:
    return 7 if ion.model_id == L.PCG2019 else 2


def state_symbols(ion):
    if ion.model_id == L.ALIEV_PANFILOV:
        return ("s", "φₘ")                                      # aliev-panfilov.jl:13-14
    return ("φₘ", "s") if ion.model_id == L.FHN else ("φₘ", "h", "m", "f", "s", "xs", "xr")


def transmembranepotential_index(ion) -> int:
    """1-based like the reference (electrophysiology.jl:107-153): position of :φₘ in state_symbols."""
    return state_symbols(ion).index("φₘ") + 1


def default_initial_state(ion) -> np.ndarray:
    """fhn.jl:19, pcg2019.jl:137-152"""
    if ion.model_id in (L.FHN, L.ALIEV_PANFILOV):
        return np.zeros(2)
    p = ion

    def sig(phi, E, k, sign):
        return 1.0 / (1.0 + math.exp(sign * (phi - E) / k))

    u0 = np.zeros(7)
    u0[0] = p.E_K
    u0[1] = sig(u0[0], p.E_h, p.k_h, 1.0)
    u0[2] = sig(u0[0], p.E_m, p.k_m, -1.0)
    u0[3] = sig(u0[0], p.E_f, p.k_f, 1.0)
    u0[4] = sig(u0[0], p.E_s, p.k_s, 1.0)
    u0[5] = sig(u0[0], p.E_xs, p.k_xs, -1.0)
    u0[6] = sig(u0[0], p.E_xr, p.k_xr, -1.0)
    return u0


# ---------------------------------------------------------------------------------------------
# coefficients, protocols, models
# ---------------------------------------------------------------------------------------------
def SymmetricTensor(dim: int, data: Sequence[float]) -> np.ndarray:
    """Tensors.jl SymmetricTensor{2,dim}(data): lower triangle, column major ((11,21,22) / (11,21,31,22,32,33))."""
    T = np.zeros((dim, dim))
    k = 0
    for j in range(dim):
        for i in range(j, dim):
            T[i, j] = T[j, i] = data[k]
            k += 1
    return T


@dataclass
class ConstantCoefficient:
    """coefficients.jl:106-120"""
    val: object


@dataclass
class FieldCoefficient:
    """coefficients.jl:36-99: elementwise_data[cell, local node, component]."""
    elementwise_data: np.ndarray


@dataclass
class OrthotropicMicrostructureModel:
    """microstructure.jl:140-187"""
    fiber_coefficient: object
    sheetlet_coefficient: object
    normal_coefficient: object


@dataclass
class SpectralTensorCoefficient:
    """coefficients.jl:451-488"""
    eigenvectors: OrthotropicMicrostructureModel
    eigenvalues: ConstantCoefficient


class CartesianCoordinateSystem:
    def __init__(self, mesh=None):
        self.mesh = mesh


@dataclass
class AnalyticalCoefficient:
    """analytical_coefficient.jl:7-39: f(x, t) evaluated in a coordinate system.

    `f` is either a Python callable (evaluated on the host at the quadrature points) or one of the
    built-in families below (evaluated inside the assembly kernel)."""
    f: object
    coordinate_system: object = None


@dataclass
class BoxStimulus:
    """maximum(x) < xmax && t < tmax ? amplitude : 0   (bak/examples/conduction-velocity-benchmark.jl:47-50)"""
    xmax: float
    tmax: float
    amplitude: float
    kind = L.SRC_BOX

    def prm(self):
        return [self.xmax, self.tmax, self.amplitude]

    def __call__(self, x, t):
        return self.amplitude if (max(x) < self.xmax and t < self.tmax) else 0.0


@dataclass
class BallStimulus:
    """norm(x) < radius && t < tmax ? amplitude : 0   (test/integration/test_electrophysiology.jl:83)"""
    radius: float
    tmax: float
    amplitude: float
    kind = L.SRC_BALL

    def prm(self):
        return [self.radius, self.tmax, self.amplitude]

    def __call__(self, x, t):
        return self.amplitude if (math.sqrt(sum(v * v for v in x)) < self.radius and t < self.tmax) else 0.0


@dataclass
class UniformEndocardialActivation:
    """docs/src/literate-tutorials/ep04_geselowitz-ecg.jl:15-26"""
    transmural_depth: float = 0.15
    tmax: float = 2.0
    amplitude: float = 0.5
    tau: float = 0.25
    kind = L.SRC_ENDO

    def prm(self):
        return [self.transmural_depth, self.tmax, self.amplitude, self.tau]


class NoStimulationProtocol:
    """electrophysiology.jl:251-255"""


@dataclass
class AnalyticalTransmembraneStimulationProtocol:
    """electrophysiology.jl:260-283"""
    f: AnalyticalCoefficient
    nonzero_intervals: Sequence[tuple]


@dataclass
class MonodomainModel:
    """electrophysiology.jl:338-368"""
    χ: ConstantCoefficient
    Cₘ: ConstantCoefficient
    κ: object
    stim: object
    ion: object
    transmembrane_solution_symbol: str = "φₘ"
    internal_state_symbol: str = "s"
    cell_coordinates: object = None


@dataclass
class ReactionDiffusionSplit:
    """electrophysiology.jl:379-381"""
    model: MonodomainModel


@dataclass
class TransientDiffusionModel:
    """core/diffusion.jl:62-70"""
    κ: object
    source: object
    solution_variable_symbol: str


@dataclass
class ConductivityToDiffusivityCoefficient:
    """coefficients.jl:122-162: κ/(Cₘ χ)"""
    conductivity_tensor_coefficient: object
    capacitance_coefficient: ConstantCoefficient
    χ_coefficient: ConstantCoefficient


# ---------------------------------------------------------------------------------------------
# discretization
# ---------------------------------------------------------------------------------------------
@dataclass
class LagrangeCollection:
    order: int = 1


@dataclass
class QuadratureRuleCollection:
    order: int = 2


@dataclass
class ElementAssemblyStrategy:
    device: object = None


SequentialAssemblyStrategy = PerColorAssemblyStrategy = ElementAssemblyStrategy


class FiniteElementDiscretization:
    """fem.jl:19-47"""

    def __init__(self, interpolations: dict, dbcs=(), qrcs=None, fqrcs=None, assembly_strategy=None):
        self.interpolations = interpolations
        self.dbcs = list(dbcs)
        self.qrcs = qrcs or {}
        self.fqrcs = fqrcs or {}
        self.assembly_strategy = assembly_strategy or ElementAssemblyStrategy()


def _extract_qrc(ipc):
    """fem.jl:52-55: QuadratureRuleCollection(max(2*order-1, 2))"""
    if isinstance(ipc, tuple):
        return ipc[1]
    return QuadratureRuleCollection(max(2 * ipc.order - 1, 2))


_default_device: B200Device | None = None


def default_device() -> B200Device:
    global _default_device
    if _default_device is None:
        _default_device = B200Device(0)
    return _default_device


def set_default_device(dev: B200Device):
    global _default_device
    _default_device = dev


def generate_mesh(celltype, nel, left, right, device: B200Device | None = None) -> DeviceMesh:
    """generate_mesh = to_mesh ∘ generate_grid (src/mesh/generators.jl:942), built in HBM."""
    return DeviceMesh.generate_grid(device or default_device(), celltype, nel, left, right)


def close_dofs(conn: np.ndarray) -> tuple[np.ndarray, int]:
    """Ferrite DofHandler close! for one Lagrange-1 field on a host grid: first-touch numbering."""
    flat = np.asarray(conn, dtype=np.int64).ravel()
    uniq, first = np.unique(flat, return_index=True)
    order = np.argsort(first, kind="stable")
    node2dof = np.full(flat.max() + 1, -1, dtype=np.int64)
    node2dof[uniq[order]] = np.arange(uniq.size)
    return node2dof[flat].reshape(np.asarray(conn).shape), int(uniq.size)


def to_mesh(celltype, cells, nodes, device: B200Device | None = None, index_base=0) -> DeviceMesh:
    """to_mesh(Grid(cells, nodes)) for a host grid (src/mesh/simple_meshes.jl:181-247)."""
    cells = np.asarray(cells, dtype=np.int64) - index_base
    celldofs, ndofs = close_dofs(cells)
    return DeviceMesh.from_host(device or default_device(), celltype, cells, nodes, celldofs, ndofs)


@dataclass
class BilinearMassIntegrator:
    ρ: ConstantCoefficient
    qrc: QuadratureRuleCollection
    sym: str


@dataclass
class BilinearDiffusionIntegrator:
    D: object
    qrc: QuadratureRuleCollection
    sym: str


@dataclass
class LinearIntegrator:
    integrand: object
    qrc: QuadratureRuleCollection


@dataclass
class AffineODEFunction:
    """functions.jl:79-88"""
    mass_term: BilinearMassIntegrator
    bilinear_term: BilinearDiffusionIntegrator
    source_term: LinearIntegrator
    dh: DeviceMesh
    strategy: object = None


@dataclass
class PointwiseODEFunction:
    """functions.jl:46-65"""
    ode: object
    x: object
    associated_states: range
    state_symbol: str = "s"


@dataclass
class GenericSplitFunction:
    functions: tuple
    solution_indices: tuple


def solution_size(f) -> int:
    if isinstance(f, GenericSplitFunction):
        return max(int(np.max(ix)) if not isinstance(ix, range) else ix.stop - 1 for ix in f.solution_indices)
    if isinstance(f, AffineODEFunction):
        return f.dh.ndofs
    return len(f.associated_states)


def semidiscretize(model, discretization: FiniteElementDiscretization, mesh: DeviceMesh):
    """fem.jl:170-196 (TransientDiffusionModel) and :371-411 (ReactionDiffusionSplit{MonodomainModel})."""
    if isinstance(model, TransientDiffusionModel):
        if discretization.dbcs:
            raise AssertionError("Dirichlet conditions not supported yet for TransientDiffusionProblem")
        sym = model.solution_variable_symbol
        if sym not in discretization.interpolations:
            raise KeyError(f"no interpolation for field {sym} in the discretization")
        ipc = discretization.interpolations[sym]
        qrc = _extract_qrc(ipc)
        ipc = ipc[0] if isinstance(ipc, tuple) else ipc
        if ipc.order != 1:
            raise NotImplementedError("only LagrangeCollection{1} is on the B200 path")
        return AffineODEFunction(
            BilinearMassIntegrator(ConstantCoefficient(1.0), discretization.qrcs.get("mass", qrc), sym),
            BilinearDiffusionIntegrator(model.κ, qrc, sym), LinearIntegrator(model.source, qrc), mesh,
            discretization.assembly_strategy)
    if isinstance(model, ReactionDiffusionSplit) and isinstance(model.model, dict):
        from . import multidomain                                   # fem.jl:434-542: one model per subdomain (+ interfaces)
        import sys as _sys
        return multidomain.semidiscretize_multidomain(model.model, discretization, mesh, default_device(), _sys.modules[__name__])
    if isinstance(model, ReactionDiffusionSplit):
        ep = model.model
        heatfun = semidiscretize(
            TransientDiffusionModel(ConductivityToDiffusivityCoefficient(ep.κ, ep.Cₘ, ep.χ), ep.stim,
                                    ep.transmembrane_solution_symbol), discretization, mesh)
        n = mesh.ndofs
        ns = num_states(ep.ion)
        odefun = PointwiseODEFunction(ep.ion, None, range(1, ns * n + 1), ep.internal_state_symbol)
        φidx = transmembranepotential_index(ep.ion)
        heat_dofrange = (φidx - 1) * n + np.arange(1, n + 1)      # 1-based, as fem.jl:399-402
        return GenericSplitFunction((heatfun, odefun), (heat_dofrange, range(1, ns * n + 1)))
    raise TypeError(f"no semidiscretize method for {type(model).__name__}")


def create_initial_condition(f) -> np.ndarray:
    """functions.jl:312-339: zeros + every model's default_initial_state, state blocked."""
    if isinstance(f, GenericSplitFunction) and hasattr(f.functions[1], "functions"):
        from . import multidomain
        import sys as _sys
        return multidomain.create_initial_condition_multidomain(f, _sys.modules[__name__])
    if isinstance(f, GenericSplitFunction):
        odefun = f.functions[1]
        n = len(odefun.associated_states) // num_states(odefun.ode)
        u = np.zeros(solution_size(f))
        for s, v in enumerate(default_initial_state(odefun.ode)):
            u[s * n:(s + 1) * n] = v
        return u
    return np.zeros(solution_size(f))


def setvariable_(u0: np.ndarray, f: GenericSplitFunction, sym: str, fn: Callable):
    """setvariable!(u0, odeform, :sym) do x ... end  (ep01_spiral-wave.jl:113-118)."""
    heatfun, odefun = f.functions
    n = heatfun.dh.ndofs
    syms = state_symbols(odefun.ode)
    φcol = transmembranepotential_index(odefun.ode) - 1
    names = {s: i for i, s in enumerate(syms)}
    names[heatfun.mass_term.sym] = φcol                          # the user's name for the transmembrane potential
    names.setdefault(odefun.state_symbol, next(i for i in range(len(syms)) if i != φcol))   # ... and for the first internal state
    if sym not in names:
        raise KeyError(f"unknown solution variable {sym}")
    s = names[sym]
    x = heatfun.dh.dof_coords()
    u0[s * n:(s + 1) * n] = [fn(xi) for xi in x]
    return u0


# ---------------------------------------------------------------------------------------------
# operators (setup_operator / update_operator! / needs_update)
# ---------------------------------------------------------------------------------------------
def _diffusion_data(D, mesh: DeviceMesh):
    """Lower a coefficient tree to (kind, data, cm_chi) of the C ABI."""
    cmchi = 1.0
    if isinstance(D, ConductivityToDiffusivityCoefficient):
        cmchi = float(D.capacitance_coefficient.val) * float(D.χ_coefficient.val)
        D = D.conductivity_tensor_coefficient
    if isinstance(D, ConstantCoefficient):
        v = np.asarray(D.val, dtype=np.float64)
        if v.ndim == 0:
            return L.D_SCALAR, v.reshape(1), cmchi
        if v.shape != (mesh.dim, mesh.dim):
            raise ValueError(f"conductivity tensor must be {mesh.dim}x{mesh.dim}")
        return L.D_TENSOR, v.ravel(), cmchi
    if isinstance(D, SpectralTensorCoefficient):
        ms = D.eigenvectors
        lam = np.asarray(D.eigenvalues.val, dtype=np.float64).ravel()
        if lam.size != 3:
            raise ValueError("orthotropic spectral coefficient needs three eigenvalues")

        def field(c):
            if isinstance(c, ConstantCoefficient):
                return np.broadcast_to(np.asarray(c.val, dtype=np.float64), (mesh.ncells, mesh.nv, 3))
            if isinstance(c, FieldCoefficient):
                return np.asarray(c.elementwise_data, dtype=np.float64).reshape(mesh.ncells, mesh.nv, 3)
            raise TypeError("fibre coefficient must be ConstantCoefficient or FieldCoefficient")

        fsn = np.stack([field(ms.fiber_coefficient), field(ms.sheetlet_coefficient), field(ms.normal_coefficient)], axis=2)
        return L.D_SPECTRAL, np.concatenate([lam, fsn.ravel()]), cmchi
    raise TypeError(f"unsupported diffusion coefficient {type(D).__name__}")


class BilinearOperator:
    """FerriteOperators BilinearFerriteOperator: fields .A, .integrator, .dh"""

    def __init__(self, integrator, dh: DeviceMesh, A: B200CSRMatrix):
        self.integrator, self.dh, self.A = integrator, dh, A


class LinearOperator:
    """FerriteOperators LinearFerriteOperator: fields .b, .integrator, .dh"""

    def __init__(self, integrator, dh: DeviceMesh, b: B200Vector):
        self.integrator, self.dh, self.b = integrator, dh, b


class LinearNullOperator:
    """interface.jl:17-64: source operator of a NoStimulationProtocol"""
    b = None


def setup_operator(strategy, integrator, solver, dh: DeviceMesh, pattern_of: B200CSRMatrix | None = None):
    if isinstance(integrator, LinearIntegrator) and isinstance(integrator.integrand, NoStimulationProtocol):
        return LinearNullOperator()
    dev = dh.dev
    if isinstance(integrator, LinearIntegrator):
        return LinearOperator(integrator, dh, B200Vector(dev, dh.ndofs, 1))
    A = pattern_of.like() if pattern_of is not None else B200CSRMatrix.from_mesh(dev, dh)
    return BilinearOperator(integrator, dh, A)


# ---------------------------------------------------------------------------------------------
# ECG post-processing
# ---------------------------------------------------------------------------------------------
class Plonsey1964ECGGaussCache:
    """src/modeling/electrophysiology/ecg.jl:55-75: built on the diffusion operator `op` (its integrator's D and
    quadrature rule, its dof handler) and a transmembrane potential vector.  The reference stores kappa*grad(phi) at
    every quadrature point between update_ecg! and evaluate_ecg; here the cache keeps a device copy of phi and the
    flux is formed in registers inside the one element sweep evaluate_ecg launches (tb_ecg_plonsey)."""

    def __init__(self, op: "BilinearOperator", φₘ):
        if not isinstance(op.integrator, BilinearDiffusionIntegrator):
            raise TypeError("Plonsey1964ECGGaussCache needs the diffusion operator")
        self.op = op
        self.φₘ = B200Vector(op.dh.dev, op.dh.ndofs, 1)
        update_ecg_(self, φₘ)


def update_ecg_(cache: Plonsey1964ECGGaussCache, φₘ, col: int = 0):
    """update_ecg!(cache, φₘ), ecg.jl:150-158.  φₘ: host array of ndofs values, or a B200Vector (state column `col`)."""
    if isinstance(φₘ, B200Vector):
        cache.φₘ.copy_from(φₘ, scol=col, dcol=0)
    else:
        φ = np.ascontiguousarray(φₘ, dtype=np.float64).ravel()
        if φ.size != cache.op.dh.ndofs:
            raise ValueError("φₘ must have one value per dof")
        cache.φₘ.upload(φ)


def evaluate_ecg(cache: Plonsey1964ECGGaussCache, x, κₜ: float):
    """evaluate_ecg(cache, x, κₜ), ecg.jl:86-148: x one point (returns a float) or a sequence of points (an array)."""
    dh = cache.op.dh
    kind, data, cmchi = _diffusion_data(cache.op.integrator.D, dh)
    pts = np.asarray(x, dtype=np.float64)
    single = pts.ndim == 1
    out = core.ecg_plonsey(dh.dev, dh, cache.op.integrator.qrc.order, kind, data, cache.φₘ, np.atleast_2d(pts), κₜ, cmchi)
    return float(out[0]) if single else out


def needs_update(op, t) -> bool:
    """src/discretization/operator.jl:2-32: closed intervals."""
    if isinstance(op, LinearNullOperator):
        return False
    proto = op.integrator.integrand
    if isinstance(proto, NoStimulationProtocol):
        return False
    return any(a <= t <= b for a, b in proto.nonzero_intervals)


def update_operator_(op, t):
    """update_operator!(op, t): THE assembly."""
    if isinstance(op, LinearNullOperator):
        return
    dh, dev = op.dh, op.dh.dev
    q = op.integrator.qrc.order
    if isinstance(op, BilinearOperator):
        if isinstance(op.integrator, BilinearMassIntegrator):
            core.assemble_mass(dev, dh, op.A, q, float(op.integrator.ρ.val))
        else:
            kind, data, cmchi = _diffusion_data(op.integrator.D, dh)
            core.assemble_diffusion(dev, dh, op.A, q, kind, data, cmchi)
        return
    f = op.integrator.integrand.f.f
    if hasattr(f, "kind"):
        core.assemble_source(dev, dh, op.b, q, f.kind, f.prm(), t)
        return
    # a closure cannot cross the C ABI, its expression can: trace it once into a postfix program the element kernel
    # evaluates at every quadrature point (trace.py); nothing but the program travels per update
    if not hasattr(op, "_program"):
        op._program = trace.trace_source(f, dh.dim) if TRACE_CLOSURES else None
    if op._program is not None:
        core.assemble_source_program(dev, dh, op.b, q, op._program.code, op._program.consts, t)
        return
    # untraceable closure (Python control flow on x or t): evaluate at the quadrature points on the host; the points are
    # computed once per operator
    if getattr(op, "_xq", None) is None:
        conn, coords, _ = dh.download()
        pts, _w = core.quadrature(dh.celltype, q)
        N = _shape_values(dh.celltype, pts)                       # nq x nv
        op._xq = np.einsum("qa,cad->cqd", N, coords[conn])        # ncells x nq x dim
    xq = op._xq
    fq = np.array([[f(xq[c, k], t) for k in range(xq.shape[1])] for c in range(xq.shape[0])])
    core.assemble_source_qp(dev, dh, op.b, q, fq)


def _shape_values(celltype, pts):
    pts = np.asarray(pts)
    if celltype == L.QUAD4:
        sx, sy = np.array([-1, 1, 1, -1.0]), np.array([-1, -1, 1, 1.0])
        return 0.25 * (1 + pts[:, :1] * sx) * (1 + pts[:, 1:2] * sy)
    if celltype == L.HEX8:
        sx, sy, sz = (np.array([-1, 1, 1, -1, -1, 1, 1, -1.0]), np.array([-1, -1, 1, 1, -1, -1, 1, 1.0]),
                      np.array([-1, -1, -1, -1, 1, 1, 1, 1.0]))
        return 0.125 * (1 + pts[:, :1] * sx) * (1 + pts[:, 1:2] * sy) * (1 + pts[:, 2:3] * sz)
    if celltype == L.TET4:
        return np.stack([1 - pts.sum(1), pts[:, 0], pts[:, 1], pts[:, 2]], axis=1)
    return np.stack([pts[:, 0], pts[:, 1], 1 - pts.sum(1)], axis=1)


# ---------------------------------------------------------------------------------------------
# solvers
# ---------------------------------------------------------------------------------------------
@dataclass
class B200CG:
    """Stands where LinearSolve.KrylovJL_CG(atol=…, rtol=…) stands (euler.jl:10); defaults are LinearSolve's."""
    atol: float = SQRT_EPS
    rtol: float = SQRT_EPS
    maxiters: int | None = None
    precs: object = None          # None, JacobiPreconditioner(), BlockJacobiPreconditioner(nblocks) or ChebyshevPreconditioner(...)
                                  # (LinearSolve's `precs`, used with ldiv = false)

    @property
    def precond(self) -> int:
        if isinstance(self.precs, JacobiPreconditioner):
            return L.PRECOND_JACOBI
        if isinstance(self.precs, BlockJacobiPreconditioner):
            return L.PRECOND_BLOCK_JACOBI
        if isinstance(self.precs, ChebyshevPreconditioner):
            return L.PRECOND_CHEBYSHEV
        return L.PRECOND_NONE

    def configure(self, dev, nrows: int):
        """what `precs(A, p)` does at init_cacheval time: hand the preconditioner's parameters to the device"""
        if isinstance(self.precs, BlockJacobiPreconditioner):
            dev.cg_set_block_jacobi(nrows, self.precs.nblocks, self.precs.row_block)
        elif isinstance(self.precs, ChebyshevPreconditioner):
            dev.cg_set_chebyshev(self.precs.degree, self.precs.ratio)


@dataclass
class JacobiPreconditioner:
    """M = diag(A)^-1, rebuilt from the operator at every solve; stands where the reference's examples put
    KrylovPreconditioners.BlockJacobiPreconditioner (bak/examples-gpu/spiral-wave.jl:95-105) with one-row blocks."""


@dataclass
class BlockJacobiPreconditioner:
    """KrylovPreconditioners.BlockJacobiPreconditioner(A, nblocks, backend) as used at bak/examples-gpu/spiral-wave.jl:95-105:
    dense inverses of `nblocks` diagonal blocks, updated from A at every solve.  row_block = block id per row (the reference
    partitions with Metis); None = equal contiguous ranges of the dof numbering."""
    nblocks: int = 1000
    row_block: object = None


@dataclass
class ChebyshevPreconditioner:
    """z = q_d(D^-1 A) D^-1 r over [lmax/ratio, lmax] (lmax: Gershgorin) -- the polynomial smoother the reference's multigrid
    configuration defaults to ("damped Jacobi with Chebyshev-optimal omega", src/solver/linear/multigrid.jl:29-33), used as a
    stand-alone preconditioner."""
    degree: int = 8
    ratio: float = 30.0


KrylovJL_CG = B200CG


@dataclass
class BackwardEulerSolver:
    """euler.jl:4-15"""
    inner_solver: B200CG = field(default_factory=B200CG)
    solution_vector_type: type = B200Vector
    system_matrix_type: type = B200CSRMatrix
    monitor: object = None


@dataclass
class ForwardEulerCellSolver:
    """partitioned_solver.jl:57-60"""
    solution_vector_type: type = B200Vector
    batch_size_hint: int = 32


@dataclass
class AdaptiveForwardEulerSubstepper:
    """partitioned_solver.jl:169-175"""
    substeps: int = 10
    reaction_threshold: float = 0.1
    solution_vector_type: type = B200Vector
    batch_size_hint: int = 32


@dataclass
class LieTrotterGodunov:
    """OS.LieTrotterGodunov((heat, cell))"""
    inner_algs: tuple


@dataclass
class ReactionTangentController:
    """rtc.jl:23-40: LieTrotterGodunov whose next step length is sigma(R_max), R_max = max_i dphi_m/dt of the cell
    sweep (rtc.jl:51-78).  A heuristic map: no error estimator, a step is rejected only when an inner solve fails."""
    inner_algs: object
    σ_s: float
    σ_c: float
    Δt_bounds: tuple

    def __post_init__(self):
        if isinstance(self.inner_algs, LieTrotterGodunov):          # rtc.jl:37-38
            self.inner_algs = self.inner_algs.inner_algs

    def next_dt(self, R: float) -> float:
        """step_accept_controller!, rtc.jl:121-133"""
        lo, hi = self.Δt_bounds
        if math.isinf(self.σ_s):
            return lo if R > self.σ_c else hi
        return (1 - 1 / (1 + math.exp((self.σ_c - R) * self.σ_s))) * (hi - lo) + lo


class ReturnCode(enum.Enum):
    Default = 0
    Success = 1
    MaxIters = 2
    DtNaN = 3
    Unstable = 4
    ConvergenceFailure = 5
    Failure = 6


@dataclass
class OperatorSplittingProblem:
    f: GenericSplitFunction
    u0: object
    tspan: tuple


@dataclass
class ODEProblem:
    f: object
    u0: object
    tspan: tuple


PointwiseODEProblem = ODEProblem


@dataclass
class IntegratorStats:
    """type.jl:1-7"""
    naccept: int = 0
    nreject: int = 0


@dataclass
class Solution:
    retcode: ReturnCode = ReturnCode.Default


class BackwardEulerSolverCache:
    """euler.jl:21-69,122-179: M, K, source, A, Δt_last + the linear solver cache."""

    def __init__(self, f: AffineODEFunction, solver: BackwardEulerSolver, t0, u: B200Vector, ucol=0):
        dh, dev = f.dh, f.dh.dev
        self.f, self.solver, self.dev, self.uₙ, self.ucol = f, solver, dev, u, ucol
        self.M = setup_operator(f.strategy, f.mass_term, solver, dh)
        self.K = setup_operator(f.strategy, f.bilinear_term, solver, dh, pattern_of=self.M.A)
        self.source_term = setup_operator(ElementAssemblyStrategy(dev), f.source_term, solver, dh)
        self.A = self.M.A.like()
        self.b = B200Vector(dev, dh.ndofs, 1)
        self.uprev = B200Vector(dev, dh.ndofs, 1)
        self.Δt_last = 0.0
        self.iters, self.resid = [], []
        update_operator_(self.M, t0)          # "initial assembly", euler.jl:172-176
        update_operator_(self.K, t0)
        update_operator_(self.source_term, t0)


class PointwiseSolverCache:
    def __init__(self, f: PointwiseODEFunction, solver, u: B200Vector):
        self.f, self.solver, self.uₙ = f, solver, u
        self.substeps = getattr(solver, "substeps", 1) if isinstance(solver, AdaptiveForwardEulerSubstepper) else 1
        self.threshold = getattr(solver, "reaction_threshold", 0.1)


def setup_solver_cache(f, solver, t0, u: B200Vector, ucol=0):
    if isinstance(f, AffineODEFunction):
        return BackwardEulerSolverCache(f, solver, t0, u, ucol)
    return PointwiseSolverCache(f, solver, u)


def _isapprox(a, b):
    return a == b or abs(a - b) <= SQRT_EPS * max(abs(a), abs(b))


def perform_step_(f, cache, t, Δt, want_tangent=False) -> bool:
    """perform_step!(f, cache, t, Δt) -> Bool for the two (function, solver) pairs on the path."""
    if isinstance(cache, BackwardEulerSolverCache):
        dev = cache.dev
        if not _isapprox(Δt, cache.Δt_last):                         # euler.jl:82
            cache.A.axpby_values(cache.M.A, cache.K.A, Δt)
            cache.Δt_last = Δt
        cache.uprev.copy_from(cache.uₙ, scol=cache.ucol)              # forward_sync of the OS child
        cache.M.A.mul(cache.b, cache.uprev)                           # b = M uprev, euler.jl:85
        if needs_update(cache.source_term, t + Δt):                  # euler.jl:118-120
            update_operator_(cache.source_term, t + Δt)
        if not isinstance(cache.source_term, LinearNullOperator):    # add!(b, S) is unconditional, euler.jl:88-91
            _add(dev, cache.b, cache.source_term.b)
        s = cache.solver.inner_solver
        s.configure(dev, cache.A.nrows)
        it, rn, conv = core.cg_solve(dev, cache.A, cache.b, cache.uₙ, s.atol, s.rtol, s.maxiters, xcol=cache.ucol,
                                     precond=s.precond)
        cache.iters.append(it)
        cache.resid.append(rn)
        return conv
    ion = f.ode
    R = core.cell_step(cache.uₙ.dev, ion.model_id, ion.params(), cache.uₙ, t, Δt, cache.substeps, cache.threshold,
                       phi_idx=transmembranepotential_index(ion) - 1, want_max=want_tangent)
    if want_tangent:
        cache.R = R                                   # max over the dumat column of phi_m (rtc.jl:64-67)
    return True


def _add(dev, b: B200Vector, s: B200Vector):
    """add!(b, S): b .+= S.b"""
    b.axpy(1.0, s)


class ThunderboltTimeIntegrator:
    """type.jl:79-126 + the OS outer integrator for LieTrotterGodunov: fields u, t, dt, stats, sol."""

    def __init__(self, prob, alg, dt, fused=True, maxiters=10**9):
        if dt is None or not (dt == dt):
            self.sol = Solution(ReturnCode.DtNaN)
            raise ValueError("dt must be given and finite")
        self.prob, self.alg = prob, alg
        self.t, self.dt = float(prob.tspan[0]), float(dt)
        self.tstop = float(prob.tspan[1])
        self.stats = IntegratorStats()
        self.sol = Solution()
        self.maxiters = maxiters
        self.iter = 0
        self.controller = alg if isinstance(alg, ReactionTangentController) else None
        self.R = 0.0                                  # ReactionTangentControllerCache.R (rtc.jl:85-88)
        self.dts = []                                 # accepted step lengths
        f = prob.f
        dev = self._device_of(f)
        self.dev = dev
        if isinstance(f, GenericSplitFunction):
            heatfun, odefun = f.functions
            ns = num_states(odefun.ode)
            self.u = prob.u0 if isinstance(prob.u0, B200Vector) else B200Vector.from_host(dev, prob.u0, ns)
            self.uprev = B200Vector(dev, self.u.n, self.u.ncols)
            heat_alg, cell_alg = alg.inner_algs
            φcol = transmembranepotential_index(odefun.ode) - 1
            self.caches = (setup_solver_cache(heatfun, heat_alg, self.t, self.u, φcol),
                           setup_solver_cache(odefun, cell_alg, self.t, self.u))
            self.fused = None
            if fused:
                hc, cc = self.caches
                st = core.MonodomainStepper(dev, hc.M.A, hc.K.A, odefun.ode.model_id, odefun.ode.params(), φcol)
                s = heat_alg.inner_solver
                st.set_cg(s.atol, s.rtol, s.maxiters)
                s.configure(dev, hc.M.A.nrows)
                st.set_preconditioner(s.precond)
                st.set_cell_solver(cc.substeps, cc.threshold)
                self.fused = st
        elif isinstance(f, AffineODEFunction):
            self.u = prob.u0 if isinstance(prob.u0, B200Vector) else B200Vector.from_host(dev, prob.u0, 1)
            self.uprev = B200Vector(dev, self.u.n, 1)
            self.caches = (setup_solver_cache(f, alg, self.t, self.u, 0),)
            self.fused = None
        else:
            ns = num_states(f.ode)
            self.u = prob.u0 if isinstance(prob.u0, B200Vector) else B200Vector.from_host(dev, prob.u0, ns)
            self.uprev = B200Vector(dev, self.u.n, ns)
            self.caches = (setup_solver_cache(f, alg, self.t, self.u),)
            self.fused = None

    @staticmethod
    def _device_of(f):
        if isinstance(f, GenericSplitFunction):
            return f.functions[0].dh.dev
        if isinstance(f, AffineODEFunction):
            return f.dh.dev
        return default_device()

    @property
    def dtcache(self):
        return self.dt

    @property
    def cg_iterations(self):
        hc = self.caches[0]
        return hc.iters if isinstance(hc, BackwardEulerSolverCache) else []

    def _functions(self):
        f = self.prob.f
        return f.functions if isinstance(f, GenericSplitFunction) else (f,)

    def _step_once(self) -> bool:
        """One accepted-or-rejected step (type.jl:189-218; rollback type.jl:510-532)."""
        t, dt = self.t, self.dt
        if self.tstop - t < dt * (1 - 1e-12):
            dt = self.tstop - t                      # land exactly on the tstop
        for c in range(self.u.ncols):
            self.uprev.copy_from(self.u, scol=c, dcol=c)
        ok = True
        if self.fused is not None:
            hc = self.caches[0]
            src = hc.source_term
            if needs_update(src, t + dt):
                update_operator_(src, t + dt)
            self.fused.set_source(None if isinstance(src, LinearNullOperator) else src.b)
            if self.controller is not None:
                it, rn, ok, self.R = self.fused.step_rt(self.u, t, dt)
            else:
                it, rn, ok = self.fused.step(self.u, t, dt)
            hc.iters.append(it)
            hc.resid.append(rn)
        else:
            for f, cache in zip(self._functions(), self.caches):     # children in tuple order: heat, then cells
                if not perform_step_(f, cache, t, dt, want_tangent=self.controller is not None):
                    ok = False
                    break
            if self.controller is not None and ok:
                self.R = max(0.0, self.dev.allreduce_max(self.caches[-1].R))   # rtc.jl:57-66: R starts at 0.0
        self.iter += 1
        if ok:
            self.stats.naccept += 1
            self.t = t + dt
            self.dts.append(dt)
            if self.controller is not None:           # stepsize_controller! + step_accept_controller!, rtc.jl:103-133
                self.dt = self.controller.next_dt(self.R)
        else:
            self.stats.nreject += 1
            for c in range(self.u.ncols):
                self.u.copy_from(self.uprev, scol=c, dcol=c)
            self.sol.retcode = ReturnCode.ConvergenceFailure       # non-adaptive: diffeq-interface.jl:347-354
        return ok


def init(prob, alg, dt=None, **kw) -> ThunderboltTimeIntegrator:
    unsupported = set(kw) - {"verbose", "maxiters", "fused"}
    if unsupported:
        raise TypeError(f"unsupported keyword arguments: {sorted(unsupported)}")
    if isinstance(prob.f, GenericSplitFunction) and hasattr(prob.f.functions[1], "functions"):
        from . import multidomain                                   # PointwiseMultiODEFunction: multi-subdomain split
        import sys as _sys
        if dt is None or not (dt == dt):
            raise ValueError("dt must be given and finite")
        return multidomain.MultiDomainIntegrator(prob, alg, dt, _sys.modules[__name__], maxiters=kw.get("maxiters", 10**9))
    return ThunderboltTimeIntegrator(prob, alg, dt, fused=kw.get("fused", True), maxiters=kw.get("maxiters", 10**9))


def step_(integ: ThunderboltTimeIntegrator) -> bool:
    """SciMLBase.step!(integrator)"""
    return integ._step_once()


def solve_(integ: ThunderboltTimeIntegrator) -> Solution:
    """SciMLBase.solve!(integrator)"""
    while integ.t < integ.tstop * (1 - 1e-15) - 1e-300 or (integ.tstop == 0 and integ.t < 0):
        if integ.iter >= integ.maxiters:
            integ.sol.retcode = ReturnCode.MaxIters
            return integ.sol
        if not integ._step_once():
            return integ.sol
    integ.sol.retcode = ReturnCode.Success
    return integ.sol
