"""Poisson and Geselowitz (lead-field) ECG reconstructions over the C ABI:
`PoissonECGReconstructionCache` (src/modeling/electrophysiology/ecg.jl:166-380) and `Geselowitz1989ECGLeadCache`
(:382-619), with `update_ecg!` / `evaluate_ecg`.

Setup (host, once): the torso DofHandler, the nodal heart -> torso transfer (`NodalIntergridInterpolation`) and the electrode
evaluation (`PointEvalHandler`) become small rectangular sparse matrices, uploaded as B200CSRMatrix; the two diffusion
operators on the torso mesh (kappa_i extended by zero outside the heart, kappa) are assembled on the device; the ground
Dirichlet condition is applied with Ferrite's apply_zero! semantics (tb_csr_apply_zero); Geselowitz' lead fields are CG
solves on the device.  Per update (device only): one transfer SpMV, one source SpMV, then either one CG solve + an
evaluation SpMV (Poisson) or `nleads` dot products with the stored lead fields (Geselowitz, tb_vec_dots).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from . import core
from .core import B200CSRMatrix, B200Vector, DeviceMesh

_REF_HEX = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float64)


def _shape(celltype, xi):
    if celltype == L.HEX8:
        return 0.125 * np.prod(1.0 + _REF_HEX * xi, axis=1)
    if celltype == L.TET4:
        return np.array([1.0 - xi.sum(), xi[0], xi[1], xi[2]])
    raise NotImplementedError("point location is implemented for Hexahedron and Tetrahedron grids")


def _dshape_hex(xi):
    d = np.empty((8, 3))
    for a in range(3):
        f = _REF_HEX[:, a].copy()
        for b in range(3):
            if b != a:
                f = f * (1.0 + _REF_HEX[:, b] * xi[b])
        d[:, a] = 0.125 * f
    return d


def _locate(celltype, X, x, tol=1e-9):
    """reference coordinates of x in the cell with vertex coordinates X, or None if outside"""
    if celltype == L.TET4:
        A = (X[1:] - X[0]).T
        xi = np.linalg.solve(A, x - X[0])
        return xi if xi.min() >= -tol and xi.sum() <= 1.0 + tol else None
    xi = np.zeros(3)
    for _ in range(30):                                  # Newton on the trilinear map
        r = _shape(celltype, xi) @ X - x
        if np.abs(r).max() < 1e-14 * max(1.0, np.abs(X).max()):
            break
        xi = xi - np.linalg.solve((_dshape_hex(xi).T @ X).T, r)
    return xi if np.abs(xi).max() <= 1.0 + tol else None


def interpolation_rows(celltype, conn, coords, celldofs, points, cells=None):
    """For each point: (dofs, weights) of the finite-element interpolant at that point -- the first cell (ascending id, optionally
    restricted to `cells`) that contains it.  Host-side point location: bounding boxes, then the inverse map."""
    conn, coords, celldofs = np.asarray(conn), np.asarray(coords), np.asarray(celldofs)
    ids = np.arange(conn.shape[0]) if cells is None else np.asarray(cells)
    Xc = coords[conn[ids]]
    lo, hi = Xc.min(axis=1), Xc.max(axis=1)
    eps = 1e-9 * max(1.0, float(np.abs(coords).max()))
    out = []
    for x in np.asarray(points, dtype=np.float64):
        cand = np.flatnonzero(((lo - eps) <= x).all(axis=1) & ((hi + eps) >= x).all(axis=1))
        row = None
        for k in cand:
            xi = _locate(celltype, Xc[k], x)
            if xi is not None:
                row = (celldofs[ids[k]], _shape(celltype, xi))
                break
        out.append(row)
    return out


def _rows_to_csr(rows, ncols):
    rowptr, colidx, vals = [0], [], []
    for r in rows:
        if r is not None:
            d, w = r
            order = np.argsort(d, kind="stable")
            colidx.extend(int(v) for v in np.asarray(d)[order])
            vals.extend(float(v) for v in np.asarray(w)[order])
        rowptr.append(len(colidx))
    return np.array(rowptr, dtype=np.int64), np.array(colidx, dtype=np.int64), np.array(vals, dtype=np.float64)


def get_closest_vertex(x, coords) -> int:
    """get_closest_vertex(position, grid): node id of the vertex nearest to x"""
    return int(np.argmin(((np.asarray(coords) - np.asarray(x)) ** 2).sum(axis=1)))


class _TorsoSetup:
    """what both reconstructions share: the torso mesh + dofs on the device, the nodal transfer from the heart, the two
    diffusion operators (source: heart conductivity, zero outside `heart_cells`; bulk: kappa with the ground constraint)"""

    def __init__(self, api, heart_dh: DeviceMesh, celltype, torso_cells, torso_nodes, kappa_i_per_cell, kappa, ground_node: int,
                 heart_cells, qorder_source=2, qorder_bulk=2):
        dev = heart_dh.dev
        self.dev = dev
        torso_cells = np.asarray(torso_cells, dtype=np.int64)
        torso_nodes = np.asarray(torso_nodes, dtype=np.float64)
        celldofs, ndofs = api.close_dofs(torso_cells)
        self.celldofs, self.ndofs, self.celltype = celldofs, ndofs, celltype
        self.cells, self.nodes = torso_cells, torso_nodes
        self.mesh = DeviceMesh.from_host(dev, celltype, torso_cells, torso_nodes, celldofs, ndofs)
        xdof = np.empty((ndofs, 3))
        xdof[celldofs.ravel()] = torso_nodes[torso_cells.ravel()]
        self.dof_coords = xdof
        # NodalIntergridInterpolation(heart_dh, torso_dh; subdomains_to = heart domain): torso dofs on the heart domain take the
        # heart field interpolated at their node, all others 0
        hconn, hcoords, hdofs = heart_dh.download()
        to_dofs = np.unique(celldofs[np.asarray(heart_cells)])
        rows = [None] * ndofs
        found = interpolation_rows(heart_dh.celltype, hconn, hcoords, hdofs, xdof[to_dofs])
        for d, r in zip(to_dofs, found):
            rows[int(d)] = r
        rp, ci, v = _rows_to_csr(rows, heart_dh.ndofs)
        self.T = B200CSRMatrix.from_pattern(dev, rp, ci, ncols=heart_dh.ndofs)
        self.T.set_nonzeros(v)
        # operators on the torso mesh
        self.K_bulk = B200CSRMatrix.from_mesh(dev, self.mesh)
        self.K_source = self.K_bulk.like()
        kind, data, cmchi = api._diffusion_data(kappa, self.mesh)
        core.assemble_diffusion(dev, self.mesh, self.K_bulk, qorder_bulk, kind, data, cmchi)
        core.assemble_diffusion(dev, self.mesh, self.K_source, qorder_source, L.D_CELL_TENSOR, np.ascontiguousarray(kappa_i_per_cell), 1.0)
        # Dirichlet(ground, 0): apply_zero! on the bulk operator, diagonal = mean |K_ii| (Ferrite's meandiag)
        node2dof = np.full(torso_nodes.shape[0], -1, dtype=np.int64)
        node2dof[torso_cells.ravel()] = celldofs.ravel()
        gdof = np.array([node2dof[int(ground_node)]], dtype=np.int64)
        self.ground_dofs = gdof
        h = C.c_void_p()
        L.call("tb_index_create", dev.h, L.ptr(gdof), 1, 0, C.byref(h))
        self.ground_index = h
        d = B200Vector(dev, ndofs, 1)
        L.call("tb_csr_diagonal", self.K_bulk.h, d.h, 0)
        self.meandiag = float(np.abs(d.to_host()).mean())
        d.free()
        L.call("tb_csr_apply_zero", self.K_bulk.h, self.ground_index, self.meandiag)
        self.phi_t = B200Vector(dev, ndofs, 1)       # φₘ_t
        self.src = B200Vector(dev, ndofs, 1)         # κ∇φₘ_t

    def source_term(self, phi_m: B200Vector, col=0):
        """transfer!(φₘ_t, T, φₘ); mul!(κ∇φₘ_t, source_op, φₘ_t)   (ecg.jl:327-331, 605-611)"""
        self.T.mul(self.phi_t, phi_m, xcol=col)
        self.K_source.mul(self.src, self.phi_t)


def _as_device_phi(dev, phi_m, n):
    if isinstance(phi_m, B200Vector):
        return phi_m, False
    v = B200Vector.from_host(dev, np.ascontiguousarray(phi_m, dtype=np.float64)[:n], 1)
    return v, True


def kappa_per_cell(fn, cells, nodes, t=0.0):
    """evaluate an AnalyticalCoefficient((x, t) -> tensor) once per cell (at the centroid): the lowering of a piecewise
    constant conductivity field to TB_D_CELL_TENSOR data"""
    xc = np.asarray(nodes)[np.asarray(cells)].mean(axis=1)
    return np.stack([np.asarray(fn(x, t), dtype=np.float64).reshape(3, 3) for x in xc]).reshape(-1)


class PoissonECGReconstructionCache:
    """ecg.jl:166-380.  heart_dh: the heart problem's DeviceMesh; torso grid as host arrays; kappa_i_per_cell: ncells x 9
    (heart conductivity on the torso mesh, zero outside the heart); kappa: coefficient object of the bulk; electrodes: points
    inside the torso; ground_node: torso node id held at zero; heart_cells: torso cells of `torso_heart_domain`."""

    def __init__(self, api, heart_dh, celltype, torso_cells, torso_nodes, kappa_i_per_cell, kappa, electrodes, ground_node, heart_cells,
                 linear_solver=None, qorder=2):
        self.s = _TorsoSetup(api, heart_dh, celltype, torso_cells, torso_nodes, kappa_i_per_cell, kappa, ground_node, heart_cells, qorder, qorder)
        s = self.s
        rows = interpolation_rows(celltype, s.cells, s.nodes, s.celldofs, electrodes)            # PointEvalHandler
        if any(r is None for r in rows):
            raise ValueError("Poisson reconstruction setup failed! Some electrodes are not found in the torso mesh")
        rp, ci, v = _rows_to_csr(rows, s.ndofs)
        self.P = B200CSRMatrix.from_pattern(s.dev, rp, ci, ncols=s.ndofs)
        self.P.set_nonzeros(v)
        self.solver = linear_solver or api.B200CG()
        self.phi_e = B200Vector(s.dev, s.ndofs, 1)
        self.rhs = B200Vector(s.dev, s.ndofs, 1)
        self.out = B200Vector(s.dev, len(rows), 1)
        self.iters = []


class Geselowitz1989ECGLeadCache:
    """ecg.jl:382-619.  electrode_sets: per lead a list of torso NODE ids, the first is the positive electrode
    (weight 1), the others share -1 (ecg.jl:571-585)."""

    def __init__(self, api, heart_dh, celltype, torso_cells, torso_nodes, kappa_i_per_cell, kappa, electrode_sets, ground_node, heart_cells,
                 linear_solver=None, qorder=2):
        self.s = _TorsoSetup(api, heart_dh, celltype, torso_cells, torso_nodes, kappa_i_per_cell, kappa, ground_node, heart_cells, qorder, qorder)
        s = self.s
        solver = linear_solver or api.B200CG()
        nl = len(electrode_sets)
        self.Z = B200Vector(s.dev, s.ndofs, nl)                       # lead fields, one per column
        node2dof = np.full(s.nodes.shape[0], -1, dtype=np.int64)
        node2dof[s.cells.ravel()] = s.celldofs.ravel()
        b, z = B200Vector(s.dev, s.ndofs, 1), B200Vector(s.dev, s.ndofs, 1)
        self.iters = []
        for i, es in enumerate(electrode_sets):
            if len(es) < 2:
                raise AssertionError(f"Electrode set {i + 1} has too few electrodes ({len(es)}<2)")
            f = np.zeros(s.ndofs)
            f[node2dof[es[0]]] = -1.0                                  # _add_electrode!: f[dof] = -weight
            for e in es[1:]:
                f[node2dof[e]] = 1.0 / (len(es) - 1)
            f[s.ground_dofs] = 0.0
            b.upload(f)
            it, rn, conv = core.cg_solve(s.dev, s.K_bulk, b, z, solver.atol, solver.rtol, solver.maxiters, precond=solver.precond)
            if not conv:
                raise RuntimeError(f"lead field {i + 1} did not converge")
            self.iters.append(it)
            self.Z.copy_from(z, scol=0, dcol=i)
        b.free()
        z.free()


def update_ecg_(cache, phi_m, col: int = 0):
    """update_ecg!(cache, φₘ)"""
    s = cache.s
    v, tmp = _as_device_phi(s.dev, phi_m, s.T.ncols)
    s.source_term(v, col if not tmp else 0)
    if tmp:
        v.free()
    if isinstance(cache, PoissonECGReconstructionCache):
        cache.rhs.fill(0.0)
        cache.rhs.axpy(-1.0, s.src)                                    # "move to the right-hand side", ecg.jl:334
        L.call("tb_vec_fill_at", cache.rhs.h, 0, s.ground_index, 0.0)  # apply_zero!(A, b, ch), ecg.jl:336
        sv = cache.solver
        it, rn, conv = core.cg_solve(s.dev, s.K_bulk, cache.rhs, cache.phi_e, sv.atol, sv.rtol, sv.maxiters, precond=sv.precond)
        cache.iters.append(it)
        if not conv:
            raise RuntimeError("Poisson ECG: the torso solve did not converge")


def evaluate_ecg(cache):
    """evaluate_ecg(cache): Poisson -> phi_e at every electrode (ecg.jl:344-347); Geselowitz -> -Z * κ∇φₘ_t (:617-619)"""
    s = cache.s
    if isinstance(cache, PoissonECGReconstructionCache):
        cache.P.mul(cache.out, cache.phi_e)
        return cache.out.to_host()
    out = np.empty(cache.Z.ncols)
    L.call("tb_vec_dots", s.dev.h, cache.Z.h, s.src.h, 0, out)
    return -out
