"""Multi-subdomain reaction-diffusion splits: the host side of

    semidiscretize(ReactionDiffusionSplit(Dict(name => MonodomainModel | InterfaceDiffusionModel)), discretization, mesh)

(src/discretization/fem.jl:245-296, 434-542), `PointwiseMultiODEFunction` with `PointBlockedLayout` blocks
(src/modeling/functions.jl:72, src/modeling/solution_variables.jl:41-68, src/solver/time/partitioned_solver.jl:23-35,126-155)
and `BilinearInterfaceDiffusionIntegrator` (src/modeling/core/diffusion.jl:81-140), over the C ABI entry points
tb_cell_step_blocks / tb_vec_gather / tb_vec_scatter / tb_assemble_interface_diffusion.

The reference takes the interface grid from FerriteInterfaceElements' `insert_interfaces` (an un-vendored package;
test/integration/test_electrophysiology.jl:131); `insert_interfaces` below is this project's restatement of what that call
has to produce for two cell sets: the nodes on their common boundary duplicated, the second set re-pointed at the copies,
one interface cell per shared facet ("here" = first set's side).  Subdomains are taken in the ORDER GIVEN (the reference
iterates a Julia Dict, whose order is unspecified): dofs are numbered first-touch over the first subdomain's cells, then the
second's, ...; interface subdomains come last and introduce no dofs (fem.jl:277-291).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib as L
from . import core
from .core import B200CSRMatrix, B200Vector, DeviceMesh

SQRT_EPS = core.SQRT_EPS


class StateBlockedLayout:
    """all points of a state consecutively (structure of arrays), solution_variables.jl:41-50"""


class PointBlockedLayout:
    """all states of a point consecutively (array of structs)"""


@dataclass
class StateBlock:
    """StateBlock(offset, npoints, nstates, layout), solution_variables.jl:53-58; offset 0-based"""
    offset: int
    npoints: int
    nstates: int
    layout: object = field(default_factory=PointBlockedLayout)


def state_range(b: StateBlock, k: int) -> range:
    """1-based slots of point k (1-based) in the solution vector, solution_variables.jl:60-68"""
    if isinstance(b.layout, StateBlockedLayout):
        first = b.offset + k
        return range(first, first + (b.nstates - 1) * b.npoints + 1, b.npoints)
    first = b.offset + (k - 1) * b.nstates + 1
    return range(first, first + b.nstates)


@dataclass
class InterfaceDiffusionModel:
    """InterfaceDiffusionModel(G, solution_variable_symbol, interface_interpolation_symbol), fem.jl:224-243"""
    G: object
    solution_variable_symbol: str = "φₘ"
    interface_interpolation_symbol: str = "φₘi"


@dataclass
class SubdomainGrid:
    """Host grid with named cell sets and interface cells (what to_mesh(insert_interfaces(grid, names)) carries)."""
    celltype: int
    cells: np.ndarray                       # ncells x nv node ids (0-based)
    nodes: np.ndarray                       # nnodes x dim
    subdomains: dict                        # name -> ascending cell ids, insertion-ordered
    interfaces: dict = field(default_factory=dict)   # name -> (here_nodes nif x k, there_nodes nif x k)


_FACETS = {L.QUAD4: [(0, 1), (1, 2), (2, 3), (3, 0)],
           L.HEX8: [(0, 3, 2, 1), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (0, 4, 7, 3), (4, 5, 6, 7)]}


def insert_interfaces(celltype: int, cells, nodes, cellsets: dict, names, interface_name: str = "interfaces") -> SubdomainGrid:
    """Duplicate the nodes shared by the two cell sets `names = (A, B)`, re-point B's cells at the copies and create one
    interface cell per facet shared by an A cell and a B cell (here side = the A cell's facet, in its local node order;
    there side = the coincident copies in the same order)."""
    if celltype not in _FACETS:
        raise NotImplementedError("interfaces are implemented for Quadrilateral and Hexahedron grids")
    a, b = names
    cells = np.array(cells, dtype=np.int64)
    nodes = np.asarray(nodes, dtype=np.float64)
    ca, cb = np.asarray(cellsets[a], dtype=np.int64), np.asarray(cellsets[b], dtype=np.int64)
    shared = np.intersect1d(np.unique(cells[ca]), np.unique(cells[cb]))
    copy_of = {int(n): nodes.shape[0] + i for i, n in enumerate(shared)}
    facets_b = set()
    for c in cb:
        for f in _FACETS[celltype]:
            facets_b.add(tuple(sorted(int(cells[c, i]) for i in f)))
    here, there = [], []
    for c in ca:                                                   # ascending cell id, local facet order: deterministic
        for f in _FACETS[celltype]:
            fn = [int(cells[c, i]) for i in f]
            if all(n in copy_of for n in fn) and tuple(sorted(fn)) in facets_b:
                here.append(fn)
                there.append([copy_of[n] for n in fn])
    new_cells = cells.copy()
    sub = new_cells[cb]
    for old, new in copy_of.items():
        sub[sub == old] = new
    new_cells[cb] = sub
    new_nodes = np.concatenate([nodes, nodes[shared]]) if shared.size else nodes
    k = len(_FACETS[celltype][0])
    return SubdomainGrid(celltype, new_cells, new_nodes, {a: np.sort(ca), b: np.sort(cb)},
                         {interface_name: (np.array(here, dtype=np.int64).reshape(-1, k), np.array(there, dtype=np.int64).reshape(-1, k))})


def close_dofs_subdomains(grid: SubdomainGrid, order) -> tuple[np.ndarray, int]:
    """Ferrite DofHandler close! with one SubDofHandler per subdomain (creation order = `order`), one Lagrange-1 field:
    first touch over the first subdomain's cells (ascending), then the next subdomain's, ...  -> (celldofs, ndofs)"""
    node2dof = np.full(grid.nodes.shape[0], -1, dtype=np.int64)
    nxt = 0
    celldofs = np.full(grid.cells.shape, -1, dtype=np.int64)
    for name in order:
        for c in grid.subdomains[name]:
            for a_, n in enumerate(grid.cells[c]):
                if node2dof[n] < 0:
                    node2dof[n] = nxt
                    nxt += 1
                celldofs[c, a_] = node2dof[n]
    return celldofs, nxt, node2dof


def union_pattern(ndofs: int, dof_lists) -> tuple[np.ndarray, np.ndarray]:
    """allocate_matrix(dh) over several cell blocks: all dof pairs sharing a cell, sorted columns, diagonal included"""
    rows, cols = [], []
    for d in dof_lists:
        d = np.asarray(d, dtype=np.int64)
        nv = d.shape[1]
        rows.append(np.repeat(d, nv, axis=1).ravel())
        cols.append(np.tile(d, (1, nv)).ravel())
    key = np.unique(np.concatenate(rows) * ndofs + np.concatenate(cols))
    r, c = key // ndofs, key % ndofs
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(r, minlength=ndofs))]).astype(np.int64)
    return rowptr, c.astype(np.int64)


@dataclass
class PointwiseMultiODEFunction:
    """functions.jl:72: a launch pad for batches of ODE steps; children carry their own (point-blocked) state ranges"""
    functions: list                      # of api.PointwiseODEFunction (associated_states = 1-based range, layout attribute)
    x: object = None


@dataclass
class MultiDomainHeatFunction:
    """AffineODEFunction with BilinearMultiIntegrator / LinearMultiIntegrator terms (fem.jl:292-298), lowered to what
    the device needs: one sub-mesh per bulk subdomain (all sharing the global dof numbering), the interface cells, the
    union sparsity pattern."""
    dev: object
    grid: SubdomainGrid
    ndofs: int
    celldofs: np.ndarray
    dof_coords: np.ndarray
    bulk: list                           # [(name, DeviceMesh, diffusion coefficient tree, qorder)]
    interfaces: list                     # [(name, dofs nif x 2k, coords_here, coords_there, D, qorder)]
    rowptr: np.ndarray = None
    colidx: np.ndarray = None


def semidiscretize_multidomain(split_models: dict, discretization, grid: SubdomainGrid, dev, api):
    """fem.jl:434-542 (+ :245-296 for the heat part).  Returns api.GenericSplitFunction((heat, ionic), (heat_dofrange, ionic range))."""
    bulk_names = [n for n, m in split_models.items() if not isinstance(m, InterfaceDiffusionModel)]
    if_names = [n for n, m in split_models.items() if isinstance(m, InterfaceDiffusionModel)]
    claimed = np.concatenate([grid.subdomains[n] for n in bulk_names])
    if np.unique(claimed).size != claimed.size:
        raise ValueError("model subdomains are not disjoint")                    # _check_model_subdomains_disjoint
    celldofs, ndofs, node2dof = close_dofs_subdomains(grid, bulk_names)
    φsyms = {m.transmembrane_solution_symbol for n, m in split_models.items() if n in bulk_names}
    if len(φsyms) != 1:
        raise AssertionError(f"All EP models in a domain split must share the same transmembrane potential symbol, got {φsyms}")
    sym = next(iter(φsyms))
    ipc = discretization.interpolations[sym]
    qrc = api._extract_qrc(ipc)
    xdof = np.empty((ndofs, grid.nodes.shape[1]))
    used = node2dof >= 0
    xdof[node2dof[used]] = grid.nodes[used]
    bulk, dof_lists = [], []
    for n in bulk_names:
        ids = grid.subdomains[n]
        ep = split_models[n]
        sub = DeviceMesh.from_host(dev, grid.celltype, grid.cells[ids], grid.nodes, celldofs[ids], ndofs)
        bulk.append((n, sub, api.ConductivityToDiffusivityCoefficient(ep.κ, ep.Cₘ, ep.χ), qrc.order))
        dof_lists.append(celldofs[ids])
    interfaces = []
    for n in if_names:
        here, there = grid.interfaces[n]
        d = np.concatenate([node2dof[here], node2dof[there]], axis=1)
        if (d < 0).any():
            raise ValueError(f"interface {n} touches nodes that no bulk subdomain claims")
        G = split_models[n].G
        interfaces.append((n, d, grid.nodes[here], grid.nodes[there], float(G.val if hasattr(G, "val") else G), qrc.order))
        dof_lists.append(d)
    rowptr, colidx = union_pattern(ndofs, dof_lists)
    heat = MultiDomainHeatFunction(dev, grid, ndofs, celldofs, xdof, bulk, interfaces, rowptr, colidx)
    # ---- ionic part: blocks packed subdomain by subdomain, point blocked (fem.jl:472-521) ----
    inner, offset = [], 0
    heat_dofrange = np.zeros(ndofs, dtype=np.int64)
    for n in bulk_names:
        ep = split_models[n]
        subdofs = np.unique(celldofs[grid.subdomains[n]])
        lo, hi = int(subdofs.min()), int(subdofs.max())
        if hi - lo + 1 != subdofs.size:
            raise AssertionError(f"{lo + 1}:{hi + 1} does not match length(subdofs)={subdofs.size} => Subdomain is not isolated.")
        ns = api.num_states(ep.ion)
        npoints = subdofs.size
        f = api.PointwiseODEFunction(ep.ion, None, range(offset + 1, offset + ns * npoints + 1), ep.internal_state_symbol)
        f.layout = PointBlockedLayout()
        f.phi_symbol = ep.transmembrane_solution_symbol
        f.block = StateBlock(offset, npoints, ns, PointBlockedLayout())
        f.subdofs = subdofs
        inner.append(f)
        φidx = api.transmembranepotential_index(ep.ion)
        heat_dofrange[subdofs] = offset + np.arange(npoints) * ns + φidx          # state_range(block, k)[φidx], 1-based
        offset += npoints * ns
    if (heat_dofrange == 0).any():
        raise ValueError("The transmembrane potential field carries dofs that no bulk model claims")
    return api.GenericSplitFunction((heat, PointwiseMultiODEFunction(inner, None)), (heat_dofrange, range(1, offset + 1)))


def solution_indices(f, sym: str, api) -> np.ndarray:
    """1-based slots of a solution variable of a multi-subdomain split (solution_variables.jl): the shared transmembrane
    potential symbol -> heat_dofrange (ordered by dof of the heat problem); a subdomain's internal-state symbol -> the
    non-phi slots of that block, point major (all internal states of point 1, then point 2, ...)."""
    heat, ionic = f.functions
    if any(getattr(fn, "phi_symbol", None) == sym for fn in ionic.functions):
        return np.asarray(f.solution_indices[0]).copy()
    for fn in ionic.functions:
        if fn.state_symbol == sym:
            b = fn.block
            φ = api.transmembranepotential_index(fn.ode)
            k = np.arange(b.npoints)[:, None] * b.nstates
            s = np.array([j for j in range(1, b.nstates + 1) if j != φ])[None, :]
            return (b.offset + k + s).ravel()
    raise KeyError(f"unknown solution variable {sym}")


def create_initial_condition_multidomain(f, api) -> np.ndarray:
    """zeros + each block's default_initial_state, point blocked"""
    ionic = f.functions[1]
    u = np.zeros(f.solution_indices[1].stop - 1)
    for fn in ionic.functions:
        b = fn.block
        u[b.offset:b.offset + b.npoints * b.nstates] = np.tile(api.default_initial_state(fn.ode), b.npoints)
    return u


class MultiDomainIntegrator:
    """LieTrotterGodunov((BackwardEulerSolver, cell solver)) over a multi-subdomain split: per step
        phi = u[heat_dofrange]  ->  b = M phi  ->  CG on A = M - dt K  ->  u[heat_dofrange] = x  ->  one cell sweep per block
    (operatorsplitting-interface.jl:23-232 with the index sets of fem.jl:523-539).  K = sum of the bulk diffusion operators
    + the interface operators, M = sum of the bulk mass operators (BilinearMultiIntegrator)."""

    def __init__(self, prob, alg, dt, api, maxiters=10**9):
        f = prob.f
        heat, ionic = f.functions
        dev = heat.dev
        self.prob, self.alg, self.api, self.dev = prob, alg, api, dev
        self.t, self.dt, self.tstop = float(prob.tspan[0]), float(dt), float(prob.tspan[1])
        self.stats, self.sol = api.IntegratorStats(), api.Solution()
        self.iter, self.maxiters = 0, maxiters
        self.controller = alg if isinstance(alg, api.ReactionTangentController) else None
        self.R, self.dts = 0.0, []
        heat_alg, cell_alg = alg.inner_algs
        self.heat_alg = heat_alg
        self.substeps = cell_alg.substeps if isinstance(cell_alg, api.AdaptiveForwardEulerSubstepper) else 1
        self.threshold = getattr(cell_alg, "reaction_threshold", 0.1)
        n = heat.ndofs
        self.u = prob.u0 if isinstance(prob.u0, B200Vector) else B200Vector.from_host(dev, prob.u0, 1)
        self.uprev = B200Vector(dev, self.u.n, 1)
        self.phi, self.b, self.x = B200Vector(dev, n, 1), B200Vector(dev, n, 1), B200Vector(dev, n, 1)
        ix = np.ascontiguousarray(f.solution_indices[0], dtype=np.int64)
        h = C.c_void_p()
        L.call("tb_index_create", dev.h, L.ptr(ix), int(ix.size), 1, C.byref(h))
        self.heat_index = h
        # operators: one pattern, bulk parts assembled per subdomain and summed, then the interface parts
        def add(acc, part):                                              # acc + part through nz(A) = nz(M) - dt nz(K), dt = -1
            out = acc.like()
            out.axpby_values(acc, part, -1.0)
            acc.free()
            return out
        self.M = B200CSRMatrix.from_pattern(dev, heat.rowptr, heat.colidx)   # values start at zero
        self.K, self.A, tmp = self.M.like(), self.M.like(), self.M.like()
        for name, sub, D, q in heat.bulk:
            core.assemble_mass(dev, sub, tmp, q, 1.0)
            self.M = add(self.M, tmp)
            kind, data, cmchi = api._diffusion_data(D, sub)
            core.assemble_diffusion(dev, sub, tmp, q, kind, data, cmchi)
            self.K = add(self.K, tmp)
        for name, d, xh, xt, G, q in heat.interfaces:
            k = xh.shape[1]
            L.call("tb_assemble_interface_diffusion", dev.h, L.FACET_LINE2 if k == 2 else L.FACET_QUAD4, int(xh.shape[2]), int(d.shape[0]),
                   L.ptr(np.ascontiguousarray(d, dtype=np.int64)), 0, L.ptr(np.ascontiguousarray(xh, dtype=np.float64)),
                   L.ptr(np.ascontiguousarray(xt, dtype=np.float64)), int(q), float(G), tmp.h)
            self.K = add(self.K, tmp)
        tmp.free()
        self.Δt_last = 0.0
        self.iters, self.resid = [], []
        blocks = (L.CellBlock * len(ionic.functions))()
        for i, fn in enumerate(ionic.functions):
            p = np.asarray(fn.ode.params(), dtype=np.float64)
            blocks[i].offset, blocks[i].npoints = fn.block.offset, fn.block.npoints
            blocks[i].model, blocks[i].layout, blocks[i].nparams = fn.ode.model_id, L.LAYOUT_POINT_BLOCKED, p.size
            for j, v in enumerate(p):
                blocks[i].params[j] = float(v)
        self.blocks = blocks

    @property
    def cg_iterations(self):
        return self.iters

    def _step_once(self) -> bool:
        api, dev = self.api, self.dev
        t, dt = self.t, self.dt
        if self.tstop - t < dt * (1 - 1e-12):
            dt = self.tstop - t
        self.uprev.copy_from(self.u)
        if not api._isapprox(dt, self.Δt_last):                            # euler.jl:82
            self.A.axpby_values(self.M, self.K, dt)
            self.Δt_last = dt
        L.call("tb_vec_gather", self.phi.h, 0, self.u.h, 0, self.heat_index)          # uprev of the heat child
        self.M.mul(self.b, self.phi)                                                  # b = M uprev, euler.jl:85
        s = self.heat_alg.inner_solver
        s.configure(dev, self.A.nrows)
        it, rn, ok = core.cg_solve(dev, self.A, self.b, self.x, s.atol, s.rtol, s.maxiters, precond=s.precond)
        self.iters.append(it)
        self.resid.append(rn)
        if ok:
            L.call("tb_vec_scatter", self.u.h, 0, self.heat_index, self.x.h, 0)
            R = C.c_double()
            L.call("tb_cell_step_blocks", dev.h, self.blocks, len(self.blocks), self.u.h, float(t), float(dt), int(self.substeps),
                   float(self.threshold), C.byref(R) if self.controller is not None else None)
            if self.controller is not None:
                self.R = max(0.0, R.value)
        self.iter += 1
        if ok:
            self.stats.naccept += 1
            self.t = t + dt
            self.dts.append(dt)
            if self.controller is not None:
                self.dt = self.controller.next_dt(self.R)
        else:
            self.stats.nreject += 1
            self.u.copy_from(self.uprev)
            self.sol.retcode = api.ReturnCode.ConvergenceFailure
        return ok
