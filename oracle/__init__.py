"""CPU oracle for the monodomain hot path -- TEST INFRASTRUCTURE ONLY.

A restatement (C, `tb_oracle.c`, wrapped here with ctypes/numpy) of the algorithms the reference
runs on the CPU for one LieTrotterGodunov step of a ReactionDiffusionSplit monodomain problem.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs
may import this package; the product package never does.

Parity status (see the header of tb_oracle.c and DESIGN.md): the reference is pure Julia and
cannot run in this image; known-answer tests of the reference are reproduced in
tests/test_oracle_known_answers.py; DoF numbering, CSR pattern, CG iteration counts and absolute
trajectories are **parity unpinned** (restated from the pinned package versions).

All indices are 0-based.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle.so"

QUAD4, HEX8, TRI3, TET4 = 0, 1, 2, 3
FHN, PCG2019, ALIEV_PANFILOV, TIMEPROBE = 0, 1, 2, 99
D_SCALAR, D_TENSOR, D_SPECTRAL = 0, 1, 2
SRC_NONE, SRC_BOX, SRC_BALL, SRC_COSEXP, SRC_NORMT, SRC_ENDO = 0, 1, 2, 3, 4, 5

_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> Path:
    """Compile the C restatement (gcc, seconds)."""
    src = _HERE / "tb_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        L.orc_grid_sizes.argtypes = [C.c_int, _i64p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_generate_grid.argtypes = [C.c_int, _i64p, _f64p, _f64p, _i64p, _f64p]
        L.orc_close_dofs.argtypes = [C.c_int64, C.c_int, _i64p, C.c_int64, _i64p, _i64p]
        L.orc_close_dofs.restype = C.c_int64
        L.orc_pattern.argtypes = [C.c_int64, C.c_int64, C.c_int, _i64p, _i64p, C.c_void_p]
        L.orc_pattern.restype = C.c_int64
        L.orc_quadrature.argtypes = [C.c_int, C.c_int, _f64p, _f64p]
        L.orc_quadrature.restype = C.c_int
        L.orc_shape.argtypes = [C.c_int, _f64p, _f64p, _f64p]
        L.orc_map_qp.argtypes = [C.c_int, _f64p, _f64p, _f64p, _f64p]
        L.orc_map_qp.restype = C.c_double
        L.orc_element_mass.argtypes = [C.c_int, C.c_int, _f64p, C.c_double, _f64p]
        L.orc_element_diffusion.argtypes = [C.c_int, C.c_int, _f64p, C.c_int, _f64p, C.c_double, C.c_int64, _f64p]
        L.orc_element_source.argtypes = [C.c_int, C.c_int, _f64p, C.c_int, _f64p, C.c_double, C.c_void_p, _f64p]
        L.orc_source_eval.argtypes = [C.c_int, _f64p, C.c_int, _f64p, C.c_double]
        L.orc_source_eval.restype = C.c_double
        L.orc_assemble_bilinear.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, _i64p, _f64p, _i64p, C.c_double,
                                            C.c_int, _f64p, C.c_double, _i64p, _i64p, _f64p]
        L.orc_assemble_bilinear_par.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, _i64p, _f64p, _i64p, C.c_int64, C.c_double,
                                                C.c_int, _f64p, C.c_double, _i64p, _i64p, _f64p]
        L.orc_assemble_source.argtypes = [C.c_int, C.c_int, C.c_int64, _i64p, _f64p, _i64p, C.c_int, _f64p,
                                          C.c_double, C.c_void_p, C.c_int64, _f64p]
        L.orc_ecg_plonsey.argtypes = [C.c_int, C.c_int, C.c_int64, _i64p, _f64p, _i64p, C.c_int, _f64p, C.c_double, _f64p,
                                      C.c_int, _f64p, C.c_double, _f64p, _f64p]
        L.orc_spmv.argtypes = [C.c_int64, _i64p, _i64p, _f64p, _f64p, _f64p]
        L.orc_axpby_values.argtypes = [C.c_int64, _f64p, _f64p, C.c_double, _f64p]
        L.orc_cg.argtypes = [C.c_int64, _i64p, _i64p, _f64p, _f64p, _f64p, C.c_double, C.c_double, C.c_int64,
                             C.c_int, _f64p, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
        L.orc_cg.restype = C.c_int64
        L.orc_pcg_jacobi.argtypes = [C.c_int64, _i64p, _i64p, _f64p, _f64p, _f64p, C.c_double, C.c_double, C.c_int64,
                                     _f64p, _f64p, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_int]
        L.orc_pcg_jacobi.restype = C.c_int64
        L.orc_cell_nstates.argtypes = [C.c_int]
        L.orc_cell_nstates.restype = C.c_int
        L.orc_fhn_default_params.argtypes = [_f64p]
        L.orc_pcg2019_default_params.argtypes = [_f64p]
        L.orc_aliev_panfilov_default_params.argtypes = [_f64p]
        L.orc_pcg2019_default_state.argtypes = [_f64p, _f64p]
        L.orc_cell_rhs.argtypes = [C.c_int, _f64p, _f64p, C.c_double, _f64p]
        L.orc_cell_step.argtypes = [C.c_int, _f64p, _f64p, _f64p, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                    C.c_int, C.c_double, C.c_int]
        L.orc_ltg_step.argtypes = [C.c_int64, _i64p, _i64p, _f64p, _f64p, C.c_void_p, C.c_int, _f64p, _f64p, _f64p,
                                   C.c_int64, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                                   C.c_double, C.c_int64, C.c_int, _f64p, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
        L.orc_ltg_step.restype = C.c_int64
        L.orc_stencil_apply.argtypes = [_i64p, _f64p, _f64p, C.c_double, C.c_double, _f64p, _f64p]
        L.orc_stencil_entry.argtypes = [_i64p, _f64p, _f64p, C.c_int, _i64p, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")]
        L.orc_stencil_entry.restype = C.c_double
        L.orc_stencil_ltg_step.argtypes = [_i64p, _f64p, _f64p, C.c_int, _f64p, _f64p, _f64p, C.c_int, C.c_double, C.c_double,
                                           C.c_int, C.c_double, C.c_double, C.c_double, C.c_int64, _f64p,
                                           C.POINTER(C.c_double), C.POINTER(C.c_int32)]
        L.orc_stencil_ltg_step.restype = C.c_int64
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def cell_nv(ct):
    return {QUAD4: 4, HEX8: 8, TRI3: 3, TET4: 4}[ct]


def cell_dim(ct):
    return 2 if ct in (QUAD4, TRI3) else 3


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


class Mesh:
    """Grid + closed DofHandler for one Lagrange-1 scalar field."""

    def __init__(self, ct, conn, coords):
        self.celltype = ct
        self.nv, self.dim = cell_nv(ct), cell_dim(ct)
        self.conn = np.ascontiguousarray(conn, dtype=np.int64).reshape(-1, self.nv)
        self.coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, self.dim)
        self.ncells, self.nnodes = self.conn.shape[0], self.coords.shape[0]
        self.celldofs = np.empty_like(self.conn)
        self.node2dof = np.empty(self.nnodes, dtype=np.int64)
        self.ndofs = lib().orc_close_dofs(self.ncells, self.nv, self.conn, self.nnodes, self.celldofs, self.node2dof)
        self._pattern = None

    @property
    def dof_coords(self):
        """Coordinates at dof locations (Lagrange-1: the vertex each dof sits on)."""
        x = np.empty((self.ndofs, self.dim))
        used = self.node2dof >= 0
        x[self.node2dof[used]] = self.coords[used]
        return x

    def pattern(self):
        """(rowptr, colidx) of allocate_matrix(dh) transposed to CSR."""
        if self._pattern is None:
            rowptr = np.empty(self.ndofs + 1, dtype=np.int64)
            nnz = lib().orc_pattern(self.ndofs, self.ncells, self.nv, self.celldofs, rowptr, None)
            colidx = np.empty(nnz, dtype=np.int64)
            lib().orc_pattern(self.ndofs, self.ncells, self.nv, self.celldofs, rowptr, colidx.ctypes.data)
            self._pattern = (rowptr, colidx)
        return self._pattern


def generate_grid(ct, nel, left, right) -> Mesh:
    """Ferrite generate_grid + DofHandler close! (src/mesh/generators.jl:942, fem.jl:180-182)."""
    dim = cell_dim(ct)
    nel3 = np.ones(3, dtype=np.int64)
    nel3[:dim] = nel
    ncells, nnodes = C.c_int64(), C.c_int64()
    lib().orc_grid_sizes(ct, nel3, C.byref(ncells), C.byref(nnodes))
    conn = np.empty((ncells.value, cell_nv(ct)), dtype=np.int64)
    coords = np.empty((nnodes.value, dim))
    l3, r3 = np.zeros(3), np.ones(3)
    l3[:dim], r3[:dim] = left, right
    lib().orc_generate_grid(ct, nel3, l3, r3, conn, coords)
    return Mesh(ct, conn, coords)


def quadrature(ct, order):
    pts, w = np.empty((64, cell_dim(ct))), np.empty(64)
    nq = lib().orc_quadrature(ct, order, pts, w)
    if nq == 0:
        raise ValueError(f"unsupported quadrature order {order} for cell type {ct}")
    return pts[:nq].copy(), w[:nq].copy()


def shape(ct, xi):
    N, dN = np.empty(cell_nv(ct)), np.empty((cell_nv(ct), cell_dim(ct)))
    lib().orc_shape(ct, np.ascontiguousarray(xi, dtype=np.float64), N, dN)
    return N, dN


def map_qp(ct, X, xi):
    """(detJ, N, gradN) at reference point xi of the cell with vertex coordinates X."""
    N, G = np.empty(cell_nv(ct)), np.empty((cell_nv(ct), cell_dim(ct)))
    det = lib().orc_map_qp(ct, np.ascontiguousarray(X, dtype=np.float64), np.ascontiguousarray(xi, dtype=np.float64), N, G)
    return det, N, G


def element_mass(ct, qorder, X, rho=1.0):
    Me = np.empty((cell_nv(ct),) * 2)
    lib().orc_element_mass(ct, qorder, np.ascontiguousarray(X, dtype=np.float64), rho, Me)
    return Me


def _ddata(data):
    return np.ascontiguousarray(np.atleast_1d(np.asarray(data, dtype=np.float64)).ravel())


def element_diffusion(ct, qorder, X, kind, data, cmchi=1.0, cell=0):
    Ke = np.empty((cell_nv(ct),) * 2)
    lib().orc_element_diffusion(ct, qorder, np.ascontiguousarray(X, dtype=np.float64), kind, _ddata(data), cmchi, cell, Ke)
    return Ke


def element_source(ct, qorder, X, kind, prm, t, fq=None):
    be = np.empty(cell_nv(ct))
    fqp = None if fq is None else np.ascontiguousarray(fq, dtype=np.float64)
    lib().orc_element_source(ct, qorder, np.ascontiguousarray(X, dtype=np.float64), kind, _ddata(prm), t,
                             None if fqp is None else fqp.ctypes.data, be)
    return be


def source_eval(kind, prm, x, t):
    x = np.ascontiguousarray(x, dtype=np.float64)
    return lib().orc_source_eval(kind, _ddata(prm), x.size, x, t)


def assemble_mass(mesh: Mesh, qorder=2, rho=1.0, threaded=False):
    """threaded=True: same values bit for bit (orc_assemble_bilinear_par), for the large parity cases / CPU baseline."""
    rowptr, colidx = mesh.pattern()
    vals = np.zeros(colidx.size)
    if threaded:
        lib().orc_assemble_bilinear_par(0, mesh.celltype, qorder, mesh.ncells, mesh.conn, mesh.coords, mesh.celldofs,
                                        mesh.ndofs, rho, 0, np.zeros(1), 1.0, rowptr, colidx, vals)
    else:
        lib().orc_assemble_bilinear(0, mesh.celltype, qorder, mesh.ncells, mesh.conn, mesh.coords, mesh.celldofs, rho, 0,
                                    np.zeros(1), 1.0, rowptr, colidx, vals)
    return vals


def assemble_diffusion(mesh: Mesh, qorder, kind, data, cmchi=1.0, threaded=False):
    """K as the reference assembles it: NEGATIVE semi-definite (diffusion.jl:28-50)."""
    rowptr, colidx = mesh.pattern()
    vals = np.zeros(colidx.size)
    if threaded:
        lib().orc_assemble_bilinear_par(1, mesh.celltype, qorder, mesh.ncells, mesh.conn, mesh.coords, mesh.celldofs,
                                        mesh.ndofs, 1.0, kind, _ddata(data), cmchi, rowptr, colidx, vals)
    else:
        lib().orc_assemble_bilinear(1, mesh.celltype, qorder, mesh.ncells, mesh.conn, mesh.coords, mesh.celldofs, 1.0, kind,
                                    _ddata(data), cmchi, rowptr, colidx, vals)
    return vals


def assemble_source(mesh: Mesh, qorder, kind, prm, t, fq_all=None):
    b = np.empty(mesh.ndofs)
    fqp = None if fq_all is None else np.ascontiguousarray(fq_all, dtype=np.float64)
    lib().orc_assemble_source(mesh.celltype, qorder, mesh.ncells, mesh.conn, mesh.coords, mesh.celldofs, kind,
                              _ddata(prm), t, None if fqp is None else fqp.ctypes.data, mesh.ndofs, b)
    return b


def ecg_plonsey(mesh: Mesh, qorder, kind, data, phi, electrodes, kappa_t, cmchi=1.0):
    """Plonsey1964ECGGaussCache: update_ecg!(cache, phi) then evaluate_ecg(cache, x, kappa_t) for every electrode x
    (src/modeling/electrophysiology/ecg.jl:55-160)."""
    el = np.ascontiguousarray(np.atleast_2d(np.asarray(electrodes, dtype=np.float64)))
    nq = len(quadrature(mesh.celltype, qorder)[1])
    flux = np.empty(mesh.ncells * nq * mesh.dim)
    out = np.empty(el.shape[0])
    lib().orc_ecg_plonsey(mesh.celltype, qorder, mesh.ncells, mesh.conn, mesh.coords, mesh.celldofs, kind, _ddata(data),
                          cmchi, np.ascontiguousarray(phi, dtype=np.float64), el.shape[0], el.reshape(-1), kappa_t, flux, out)
    return out


def spmv(rowptr, colidx, vals, x):
    y = np.empty(rowptr.size - 1)
    lib().orc_spmv(rowptr.size - 1, rowptr, colidx, vals, np.ascontiguousarray(x, dtype=np.float64), y)
    return y


def axpby_values(M, K, dt):
    A = np.empty_like(M)
    lib().orc_axpby_values(M.size, M, K, dt, A)
    return A


SQRT_EPS = float(np.sqrt(np.finfo(np.float64).eps))


def cg(rowptr, colidx, vals, b, atol=SQRT_EPS, rtol=SQRT_EPS, itmax=None, threaded_blas1=False):
    """LinearSolve.KrylovJL_CG defaults: abstol = reltol = sqrt(eps), maxiters = length(b), x0 = 0.
    threaded_blas1: False/0 serial sums (Krylov on Vector), True/1 threaded, 2 = order-free "exact" dot products
    (double-double, rounded once; the product's tb_cg_set_exact_dot meets it bit for bit)."""
    n = rowptr.size - 1
    itmax = n if itmax is None else itmax
    x, work = np.empty(n), np.empty(3 * n)
    rn, conv = C.c_double(), C.c_int32()
    it = lib().orc_cg(n, rowptr, colidx, vals, np.ascontiguousarray(b, dtype=np.float64), x, atol, rtol, itmax,
                      int(threaded_blas1), work, C.byref(rn), C.byref(conv))
    return x, int(it), rn.value, bool(conv.value)


def pcg_jacobi(rowptr, colidx, vals, b, atol=SQRT_EPS, rtol=SQRT_EPS, itmax=None, dot_mode=0):
    """KrylovJL_CG with a Jacobi preconditioner (precs, ldiv = false): stops on sqrt(r.z) <= atol + rtol*sqrt(r0.z0)."""
    n = rowptr.size - 1
    itmax = n if itmax is None else itmax
    x, work, dinv = np.empty(n), np.empty(4 * n), np.empty(n)
    rn, conv = C.c_double(), C.c_int32()
    it = lib().orc_pcg_jacobi(n, rowptr, colidx, np.ascontiguousarray(vals, dtype=np.float64),
                              np.ascontiguousarray(b, dtype=np.float64), x, atol, rtol, itmax, dinv, work, C.byref(rn), C.byref(conv),
                              int(dot_mode))
    return x, int(it), rn.value, bool(conv.value)


def cell_nstates(model):
    return lib().orc_cell_nstates(model)


def default_params(model):
    if model == FHN:
        p = np.empty(6)
        lib().orc_fhn_default_params(p)
    elif model == PCG2019:
        p = np.empty(36)
        lib().orc_pcg2019_default_params(p)
    elif model == ALIEV_PANFILOV:
        p = np.empty(6)
        lib().orc_aliev_panfilov_default_params(p)
    else:
        p = np.zeros(1)
    return p


def default_initial_state(model, prm=None):
    """default_initial_state (fhn.jl:19, pcg2019.jl:137-152)."""
    if model in (FHN, ALIEV_PANFILOV):
        return np.zeros(2)
    if model == PCG2019:
        u0 = np.empty(7)
        lib().orc_pcg2019_default_state(default_params(PCG2019) if prm is None else prm, u0)
        return u0
    return np.zeros(1)


def cell_rhs(model, prm, u, t=0.0):
    du = np.empty(cell_nstates(model))
    lib().orc_cell_rhs(model, _ddata(prm), np.ascontiguousarray(u, dtype=np.float64), t, du)
    return du


def cell_step(model, prm, u, n, t, dt, substeps=1, threshold=0.1, phi_idx=0, ld=None):
    """In-place cell sweep on the SoA state vector u (state s of node i at u[s*ld+i]). Returns du."""
    ld = n if ld is None else ld
    du = np.zeros_like(u)
    lib().orc_cell_step(model, _ddata(prm), u, du, n, ld, t, dt, substeps, threshold, phi_idx)
    return du


def rtc_next_dt(R, sigma_s, sigma_c, dt_bounds):
    """ReactionTangentController's step_accept_controller! (src/solver/time/rtc.jl:121-133): the next step length
    is sigma(R); sigma_s = Inf makes it a step function with R == sigma_c on the dt_max side."""
    lo, hi = dt_bounds
    if np.isinf(sigma_s):
        return lo if R > sigma_c else hi
    return (1 - 1 / (1 + np.exp((sigma_c - R) * sigma_s))) * (hi - lo) + lo


class MonodomainOracle:
    """Holds M, K, A = M - dt K and steps LTG(BackwardEuler, cell solver) like the reference."""

    def __init__(self, mesh: Mesh, model, prm, Mvals, Kvals, phi_idx=0, atol=SQRT_EPS, rtol=SQRT_EPS, itmax=None,
                 substeps=1, threshold=0.1, threaded_blas1=False, precond=None):
        self.mesh, self.model, self.prm = mesh, model, _ddata(prm)
        self.rowptr, self.colidx = mesh.pattern()
        self.M, self.K = Mvals, Kvals
        self.n = mesh.ndofs
        self.phi_idx = phi_idx
        self.atol, self.rtol = atol, rtol
        self.itmax = self.n if itmax is None else itmax
        self.substeps, self.threshold = substeps, threshold
        self.threaded_blas1 = threaded_blas1
        self.precond = precond        # None or "jacobi" (KrylovJL_CG(precs = ...))
        self.dt_last = 0.0
        self.A = None
        self.bS = None  # last assembled source vector (stays added once assembled, euler.jl:88-91)
        self.work = np.empty(5 * self.n)
        self.du = np.zeros(cell_nstates(model) * self.n)
        self.iters = []

    def step(self, u, t, dt):
        if self.A is None or not np.isclose(dt, self.dt_last, rtol=np.sqrt(np.finfo(float).eps), atol=0):
            self.A = axpby_values(self.M, self.K, dt)
            self.dt_last = dt
        if self.precond == "jacobi":
            # same LTG step composed from the exported pieces: b = M phi (+ bS); PCG; cell sweep (euler.jl:71-101)
            n, pi = self.n, self.phi_idx
            b = spmv(self.rowptr, self.colidx, self.M, u[pi * n:(pi + 1) * n])
            if self.bS is not None:
                b += self.bS
            x, it, rnv, cv = pcg_jacobi(self.rowptr, self.colidx, self.A, b, self.atol, self.rtol, self.itmax,
                                        dot_mode=2 if int(self.threaded_blas1) == 2 else 0)
            u[pi * n:(pi + 1) * n] = x
            lib().orc_cell_step(self.model, self.prm, u, self.du, n, n, t, dt, self.substeps, self.threshold, pi)
            self.iters.append(int(it))
            return int(it), rnv, bool(cv)
        rn, conv = C.c_double(), C.c_int32()
        it = lib().orc_ltg_step(self.n, self.rowptr, self.colidx, self.A, self.M,
                                None if self.bS is None else self.bS.ctypes.data, self.model, self.prm, u, self.du,
                                self.n, self.phi_idx, t, dt, self.substeps, self.threshold, self.atol, self.rtol,
                                self.itmax, int(self.threaded_blas1), self.work, C.byref(rn), C.byref(conv))
        self.iters.append(int(it))
        return int(it), rn.value, bool(conv.value)

    def reaction_tangent(self):
        """get_reaction_tangent (rtc.jl:51-78): R = 0.0; R = max(R, maximum(dumat[:, phi_m])) -- the signed maximum
        (not maximum(abs)) of the phi_m column of the cell solver's du left by the last step, never below zero."""
        return max(0.0, float(self.du[self.phi_idx * self.n:(self.phi_idx + 1) * self.n].max()))


class StencilOracle:
    """Matrix-free twin of MonodomainOracle for a UNIFORM hexahedral grid with a diagonal diffusion tensor (closed-form
    27-point operator, tb_oracle.c: orc_stencil_*).  Exists because the assembled CSR image of BASELINE config 5
    (2.7 G nonzeros) does not fit the host; vectors are in grid order (node a + (nx+1)*(b + (ny+1)*c))."""

    def __init__(self, nel, h, kappa, model, prm, atol=SQRT_EPS, rtol=SQRT_EPS, itmax=None, substeps=1, threshold=0.1, phi_idx=0):
        self.nel = np.ascontiguousarray(nel, dtype=np.int64)
        self.h = np.ascontiguousarray(np.broadcast_to(np.asarray(h, dtype=np.float64), (3,)))
        self.kappa = np.ascontiguousarray(kappa, dtype=np.float64)
        self.n = int(np.prod(self.nel + 1))
        self.model, self.prm = model, np.ascontiguousarray(prm, dtype=np.float64)
        self.atol, self.rtol, self.itmax = atol, rtol, self.n if itmax is None else itmax
        self.substeps, self.threshold, self.phi_idx = substeps, threshold, phi_idx
        self.work = np.empty(5 * self.n)
        self.du = np.zeros(cell_nstates(model) * self.n)
        self.iters = []

    def node_id(self, a, b, c):
        return a + (self.nel[0] + 1) * (b + (self.nel[1] + 1) * c)

    def node_coords(self):
        ax = [np.arange(int(n) + 1) * hh for n, hh in zip(self.nel, self.h)]
        Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
        return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)

    def apply(self, x, cm=1.0, ck=0.0):
        y = np.empty(self.n)
        lib().orc_stencil_apply(self.nel, self.h, self.kappa, cm, ck, np.ascontiguousarray(x, dtype=np.float64), y)
        return y

    def entry(self, which, node, off):
        return lib().orc_stencil_entry(self.nel, self.h, self.kappa, which, np.ascontiguousarray(node, dtype=np.int64),
                                       np.ascontiguousarray(off, dtype=np.int32))

    def step(self, u, t, dt):
        rn, conv = C.c_double(), C.c_int32()
        it = lib().orc_stencil_ltg_step(self.nel, self.h, self.kappa, self.model, self.prm, u, self.du, self.phi_idx, t, dt,
                                        self.substeps, self.threshold, self.atol, self.rtol, self.itmax, self.work,
                                        C.byref(rn), C.byref(conv))
        self.iters.append(int(it))
        return int(it), rn.value, bool(conv.value)


# ---- general preconditioned CG (KrylovJL_CG(precs = ..., ldiv = false)) -----------------------------------
def pcg(rowptr, colidx, vals, b, apply_pc, atol=SQRT_EPS, rtol=SQRT_EPS, itmax=None):
    """Krylov.jl cg! with a preconditioner given as a function z = P(r) (P applies the INVERSE, ldiv = false):
    gamma = r.z, stop on sqrt(r.z) <= atol + rtol*sqrt(r0.z0), p = z + beta p.  Vector algebra in numpy, SpMV in C."""
    n = rowptr.size - 1
    itmax = n if itmax is None else itmax
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    x, r = np.zeros(n), np.array(b, dtype=np.float64)
    z = apply_pc(r)
    p = z.copy()
    gamma = float(np.dot(r, z))
    rn = np.sqrt(gamma)
    eps = atol + rtol * rn
    solved, it = rn <= eps, 0
    while not (solved or it >= itmax):
        Ap = spmv(rowptr, colidx, vals, p)
        alpha = gamma / float(np.dot(p, Ap))
        x += alpha * p
        r -= alpha * Ap
        z = apply_pc(r)
        gnext = float(np.dot(r, z))
        rn = np.sqrt(gnext)
        solved = rn <= eps
        if not solved:
            beta = gnext / gamma
            gamma = gnext
            p = z + beta * p
        it += 1
    return x, it, float(rn), bool(solved)


def _diag(rowptr, colidx, vals):
    n = rowptr.size - 1
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    d = np.zeros(n)
    m = colidx == rows
    d[rows[m]] = vals[m]
    return d


def gershgorin_lmax(rowptr, colidx, vals):
    """max_i sum_j |a_ij| / a_ii: an upper bound of the spectrum of D^-1 A."""
    n = rowptr.size - 1
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    s = np.bincount(rows, weights=np.abs(vals), minlength=n)
    return float((s / _diag(rowptr, colidx, vals)).max())


def chebyshev_preconditioner(rowptr, colidx, vals, degree=8, ratio=30.0):
    """z = q_d(D^-1 A) D^-1 r: `degree` terms of the Chebyshev iteration for A z = r on [lmax/ratio, lmax] (Saad, Iterative
    Methods, Alg. 12.1 with the Jacobi splitting), lmax = Gershgorin bound.  Returns the apply function."""
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    dinv = 1.0 / _diag(rowptr, colidx, vals)
    hi = gershgorin_lmax(rowptr, colidx, vals)
    lo = hi / ratio
    theta, delta = 0.5 * (hi + lo), 0.5 * (hi - lo)
    sigma1 = theta / delta

    def apply(r):
        res = r.copy()
        d = (1.0 / theta) * (dinv * res)
        z = d.copy()
        rho = 1.0 / sigma1
        for _ in range(1, degree):
            res = res - spmv(rowptr, colidx, vals, d)
            rho_new = 1.0 / (2.0 * sigma1 - rho)
            d = (rho_new * rho) * d + (2.0 * rho_new / delta) * (dinv * res)
            z = z + d
            rho = rho_new
        return z
    return apply


def block_jacobi_preconditioner(rowptr, colidx, vals, nblocks, row_block=None):
    """KrylovPreconditioners.BlockJacobiPreconditioner(A, nblocks) (bak/examples-gpu/spiral-wave.jl:95-105): dense inverses of
    the diagonal blocks.  row_block = block id per row (the reference: Metis), None = equal contiguous ranges."""
    n = rowptr.size - 1
    blk = (np.arange(n) * nblocks) // n if row_block is None else np.asarray(row_block)
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    same = blk[rows] == blk[colidx]
    order = np.argsort(blk, kind="stable")
    pos = np.empty(n, dtype=np.int64)
    counts = np.bincount(blk, minlength=nblocks)
    start = np.concatenate([[0], np.cumsum(counts)])
    pos[order] = np.arange(n) - start[blk[order]]
    inv = []
    for b in range(nblocks):
        B = np.zeros((counts[b], counts[b]))
        m = same & (blk[rows] == b)
        B[pos[rows[m]], pos[colidx[m]]] = vals[m]
        inv.append(np.linalg.inv(B))
    members = [order[start[b]:start[b + 1]] for b in range(nblocks)]

    def apply(r):
        z = np.empty_like(r)
        for b in range(nblocks):
            z[members[b]] = inv[b] @ r[members[b]]
        return z
    return apply
