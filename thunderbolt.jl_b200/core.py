"""Handle objects over the C ABI: B200Device, B200Vector, DeviceMesh, B200CSRMatrix, MonodomainStepper.

These are the Python twins of the Julia types INTEGRATION.md introduces (`B200Device`,
`B200Vector{Float64}`, `B200CSRMatrix{Float64,Int32}`); every method is one C call.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

SQRT_EPS = float(np.sqrt(np.finfo(np.float64).eps))


class B200Device:
    """A CUDA context on one B200 (`src/devices.jl:1-4` maps devices to backends; this is the new one)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        h = C.c_void_p()
        L.call("tb_ctx_create", int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        self.h = h
        self.device = int(device)
        self.rank, self.nranks = 0, 1

    def close(self):
        if getattr(self, "h", None):
            L.lib().tb_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        L.call("tb_sync", self.h)

    def info(self):
        sm, mem, ma, mi = C.c_int32(), C.c_int64(), C.c_int32(), C.c_int32()
        L.call("tb_device_info", self.h, C.byref(sm), C.byref(mem), C.byref(ma), C.byref(mi))
        return {"sm_count": sm.value, "total_mem": mem.value, "cc": (ma.value, mi.value)}

    def timer_start(self):
        L.call("tb_timer_start", self.h)

    def timer_stop(self) -> float:
        ms = C.c_double()
        L.call("tb_timer_stop", self.h, C.byref(ms))
        return ms.value

    def launch_count(self) -> int:
        n = C.c_int64()
        L.call("tb_launch_count", self.h, C.byref(n))
        return n.value

    def l2_flush(self):
        L.call("tb_l2_flush", self.h)

    def profile_enable(self, on=True):
        L.call("tb_profile_enable", self.h, int(on))

    def profile_get(self):
        """(total ms, launches) of the SpMV-in-CG kernel since profile_enable(True)."""
        ms, n = C.c_double(), C.c_int64()
        L.call("tb_profile_get", self.h, C.byref(ms), C.byref(n))
        return ms.value, n.value

    # ---- CG execution model ----
    def cg_set_persistent_variant(self, variant):
        """register-resident persistent CG: 2 (default) two flag barriers per iteration, 1 three grid-wide barriers"""
        L.call("tb_cg_set_persistent_variant", self.h, int(variant))

    def cg_set_persistent(self, mode):
        """whole CG solve in one cooperative kernel: 0/False never, 1/True auto (default), 2 the TMA-staged variant
        whenever eligible (single GPU only)"""
        L.call("tb_cg_set_persistent", self.h, int(mode))

    def cg_last_path(self) -> int:
        """0 multi-kernel, 1 persistent with register-resident vectors, 2 persistent with the TMA sweep"""
        v = C.c_int32()
        L.call("tb_cg_last_path", self.h, C.byref(v))
        return v.value

    def cg_set_exact_dot(self, on=True):
        """CG dot products in double-double, rounded once: iterates independent of grid size / GPU count (solver path 0)"""
        L.call("tb_cg_set_exact_dot", self.h, int(bool(on)))

    def cg_set_block_jacobi(self, nrows: int, nblocks: int, row_block=None):
        """BlockJacobiPreconditioner(A, nblocks): row_block = block id per row, None = equal contiguous dof ranges"""
        rb = None if row_block is None else np.ascontiguousarray(row_block, dtype=np.int32)
        L.call("tb_cg_set_block_jacobi", self.h, int(nrows), int(nblocks), L.ptr(rb))
        self._bj_keepalive = rb

    def cg_set_chebyshev(self, degree: int = 8, ratio: float = 30.0):
        L.call("tb_cg_set_chebyshev", self.h, int(degree), float(ratio))

    def cg_last_path_persistent(self) -> bool:
        return self.cg_last_path() != 0

    # ---- assembly strategy ---------------------------------------------------------------------------
    def assembly_set_mode(self, mode: int):
        """2 = per-element results + ordered gather (deterministic, default), 0 = fp64 atomic scatter."""
        L.call("tb_assembly_set_mode", self.h, int(mode))

    def assembly_info(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        L.call("tb_assembly_info", self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"mode": a.value, "last_mode": b.value, "last_chunks": c.value}

    def assembly_set_scratch_budget(self, nbytes: int):
        L.call("tb_assembly_set_scratch_budget", self.h, int(nbytes))

    def assembly_release_scratch(self):
        L.call("tb_assembly_release_scratch", self.h)

    # ---- multi-GPU --------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        L.call("tb_comm_unique_id", buf)
        return buf.raw

    def comm_init(self, rank: int, nranks: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        L.call("tb_ctx_comm_init", self.h, int(rank), int(nranks), buf)
        self.rank, self.nranks = int(rank), int(nranks)

    def barrier(self):
        L.call("tb_comm_barrier", self.h)

    # NVLink peer-memory path (tb_dist.cu): export this rank's window + CG work vectors, attach everybody's
    def peer_export(self, ncols: int) -> bytes:
        buf = C.create_string_buffer(L.PEER_BLOB_BYTES)
        L.call("tb_peer_export", self.h, int(ncols), buf)
        return buf.raw

    def peer_attach(self, blobs: bytes, nranks: int):
        buf = C.create_string_buffer(blobs, len(blobs))
        L.call("tb_peer_attach", self.h, buf, int(nranks))

    def peer_enabled(self) -> bool:
        on = C.c_int32()
        L.call("tb_peer_enabled", self.h, C.byref(on))
        return bool(on.value)

    def peer_stats(self, reset=False):
        """ms one CTA of this rank waited for peers (all-reduce collects, halo flags) since the last reset, and the wait counts"""
        a, na, h, nh = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
        L.call("tb_peer_stats", self.h, C.byref(a), C.byref(na), C.byref(h), C.byref(nh), int(reset))
        return {"ar_wait_ms": a.value, "ar_waits": na.value, "halo_wait_ms": h.value, "halo_waits": nh.value}

    def allreduce_max(self, v: float) -> float:
        x = C.c_double(v)
        L.call("tb_comm_allreduce_max", self.h, C.byref(x))
        return x.value


class B200Vector:
    """Device vector with `ncols` state columns of `n` rows (host image: column c = host[c*n:(c+1)*n])."""

    def __init__(self, dev: B200Device, n: int, ncols: int = 1):
        h = C.c_void_p()
        L.call("tb_vec_create", dev.h, int(n), int(ncols), C.byref(h))
        self.h, self.dev, self.n, self.ncols = h, dev, int(n), int(ncols)

    @classmethod
    def from_host(cls, dev, host, ncols: int = 1):
        host = np.ascontiguousarray(host, dtype=np.float64).ravel()
        assert host.size % ncols == 0
        v = cls(dev, host.size // ncols, ncols)
        v.upload(host)
        return v

    def __len__(self):
        return self.n * self.ncols

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=np.float64).ravel()
        if host.size != self.n * self.ncols:
            raise ValueError(f"host vector has {host.size} entries, device vector {self.n * self.ncols}")
        L.call("tb_vec_upload", self.h, L.ptr(host))

    def to_host(self) -> np.ndarray:
        out = np.empty(self.n * self.ncols)
        L.call("tb_vec_download", self.h, L.ptr(out))
        return out

    def column(self, col: int) -> np.ndarray:
        out = np.empty(self.n)
        L.call("tb_vec_download_col", self.h, int(col), L.ptr(out), 0, self.n)
        return out

    def set_column(self, col: int, host):
        host = np.ascontiguousarray(host, dtype=np.float64).ravel()
        L.call("tb_vec_upload_col", self.h, int(col), L.ptr(host), 0, host.size)

    def fill(self, value: float, col: int = 0):
        L.call("tb_vec_fill", self.h, int(col), float(value))

    def copy_from(self, src: "B200Vector", scol: int = 0, dcol: int = 0):
        L.call("tb_vec_copy", self.h, int(dcol), src.h, int(scol))

    def axpy(self, a: float, x: "B200Vector", xcol: int = 0, ycol: int = 0):
        """self[:, ycol] += a * x[:, xcol]"""
        L.call("tb_vec_axpy", self.h, int(ycol), float(a), x.h, int(xcol))

    def devptr(self, col: int = 0):
        p, ld = C.c_void_p(), C.c_int64()
        L.call("tb_vec_devptr", self.h, int(col), C.byref(p), C.byref(ld))
        return p.value, ld.value

    def free(self):
        if getattr(self, "h", None):
            L.lib().tb_vec_destroy(self.h)
            self.h = None


class DeviceMesh:
    """Grid + closed DofHandler (one Lagrange-1 scalar field) resident in HBM."""

    def __init__(self, dev, h):
        self.dev, self.h = dev, h
        nc, nn, nd, nv, dim = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
        L.call("tb_mesh_sizes", h, C.byref(nc), C.byref(nn), C.byref(nd), C.byref(nv), C.byref(dim))
        self.ncells, self.nnodes, self.ndofs, self.nv, self.dim = nc.value, nn.value, nd.value, nv.value, dim.value
        self.celltype = None
        self.ndofs_owned = self.ndofs
        self.dof_lo = 0
        self.ghost_global = np.empty(0, dtype=np.int64)

    @classmethod
    def generate_grid(cls, dev, celltype, nel, left, right):
        dim = 2 if celltype in (L.QUAD4, L.TRI3) else 3
        nel3 = np.ones(3, dtype=np.int64)
        nel3[:dim] = nel
        l3, r3 = np.zeros(3), np.ones(3)
        l3[:dim], r3[:dim] = left, right
        h = C.c_void_p()
        L.call("tb_mesh_generate_grid", dev.h, int(celltype), nel3, l3, r3, C.byref(h))
        m = cls(dev, h)
        m.celltype = int(celltype)
        return m

    @classmethod
    def generate_grid_local(cls, dev, celltype, nel, left, right, dof_lo: int, dof_hi: int):
        """the local part [dof_lo, dof_hi) of a structured Quadrilateral / Hexahedron grid, built from the closed-form numbering:
        the same mesh as generate_grid(...).extract_local(dof_lo, dof_hi) without the global grid in HBM"""
        dim = 2 if celltype in (L.QUAD4, L.TRI3) else 3
        nel3 = np.ones(3, dtype=np.int64)
        nel3[:dim] = nel
        l3, r3 = np.zeros(3), np.ones(3)
        l3[:dim], r3[:dim] = left, right
        h, ng = C.c_void_p(), C.c_int64()
        L.call("tb_mesh_generate_grid_local", dev.h, int(celltype), nel3, l3, r3, int(dof_lo), int(dof_hi), C.byref(h), C.byref(ng))
        m = cls(dev, h)
        m.celltype = int(celltype)
        m.ndofs_owned = int(dof_hi - dof_lo)
        m.dof_lo = int(dof_lo)
        m.ghost_global = np.empty(ng.value, dtype=np.int64)
        if ng.value:
            L.call("tb_mesh_ghosts", h, m.ghost_global)
        return m

    @classmethod
    def from_host(cls, dev, celltype, conn, coords, celldofs, ndofs, index_base=0):
        conn = np.ascontiguousarray(conn, dtype=np.int64)
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        celldofs = np.ascontiguousarray(celldofs, dtype=np.int64)
        nv = {L.QUAD4: 4, L.HEX8: 8, L.TRI3: 3, L.TET4: 4}[celltype]
        dim = 2 if celltype in (L.QUAD4, L.TRI3) else 3
        h = C.c_void_p()
        L.call("tb_mesh_create", dev.h, int(celltype), conn.size // nv, coords.size // dim, conn.ravel(), coords.ravel(),
               celldofs.ravel(), int(ndofs), int(index_base), C.byref(h))
        m = cls(dev, h)
        m.celltype = int(celltype)
        return m

    def extract_local(self, dof_lo: int, dof_hi: int) -> "DeviceMesh":
        h, ng = C.c_void_p(), C.c_int64()
        L.call("tb_mesh_extract_local", self.h, int(dof_lo), int(dof_hi), C.byref(h), C.byref(ng))
        m = DeviceMesh(self.dev, h)
        m.celltype = self.celltype
        m.ndofs_owned = int(dof_hi - dof_lo)
        m.dof_lo = int(dof_lo)
        m.ghost_global = np.empty(ng.value, dtype=np.int64)
        if ng.value:
            L.call("tb_mesh_ghosts", h, m.ghost_global)
        return m

    def set_ownership(self, ndofs_owned: int, dof_lo: int, ghost_global):
        """mark a host-cut local mesh: owned dofs first (global ids dof_lo ..), ghosts after (ascending global ids)"""
        gg = np.ascontiguousarray(ghost_global, dtype=np.int64)
        L.call("tb_mesh_set_ownership", self.h, int(ndofs_owned), int(dof_lo), L.ptr(gg), int(gg.size))
        self.ndofs_owned, self.dof_lo, self.ghost_global = int(ndofs_owned), int(dof_lo), gg

    def download(self):
        conn = np.empty((self.ncells, self.nv), dtype=np.int64)
        coords = np.empty((self.nnodes, self.dim))
        celldofs = np.empty((self.ncells, self.nv), dtype=np.int64)
        L.call("tb_mesh_download", self.h, L.ptr(conn), L.ptr(coords), L.ptr(celldofs))
        return conn, coords, celldofs

    def dof_coords(self) -> np.ndarray:
        x = np.empty((self.ndofs, self.dim))
        L.call("tb_mesh_dof_coords", self.h, x.reshape(-1))
        return x

    def free(self):
        if getattr(self, "h", None):
            L.lib().tb_mesh_destroy(self.h)
            self.h = None


class B200CSRMatrix:
    """CSR operator (SELL-32 in HBM) with the reference's sparsity pattern."""

    def __init__(self, dev, h):
        self.dev, self.h = dev, h
        nr, nc, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        L.call("tb_csr_sizes", h, C.byref(nr), C.byref(nc), C.byref(nnz))
        self.nrows, self.ncols, self.nnz = nr.value, nc.value, nnz.value

    @classmethod
    def from_pattern(cls, dev, rowptr, colidx, ncols=None, index_base=0):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        colidx = np.ascontiguousarray(colidx, dtype=np.int64)
        n = rowptr.size - 1
        h = C.c_void_p()
        L.call("tb_csr_create", dev.h, n, int(n if ncols is None else ncols), rowptr, colidx, int(index_base), C.byref(h))
        return cls(dev, h)

    @classmethod
    def from_mesh(cls, dev, mesh: DeviceMesh):
        h = C.c_void_p()
        L.call("tb_csr_create_from_mesh", dev.h, mesh.h, C.byref(h))
        return cls(dev, h)

    def like(self) -> "B200CSRMatrix":
        h = C.c_void_p()
        L.call("tb_csr_create_like", self.h, C.byref(h))
        return B200CSRMatrix(self.dev, h)

    def storage(self):
        """(stored entries incl. padding, bytes of the column stream the SpMV reads, widest slice)"""
        se, cb, mw = C.c_int64(), C.c_int64(), C.c_int32()
        L.call("tb_csr_storage", self.h, C.byref(se), C.byref(cb), C.byref(mw))
        return se.value, cb.value, mw.value

    def pattern(self, index_base=0):
        rowptr = np.empty(self.nrows + 1, dtype=np.int64)
        colidx = np.empty(self.nnz, dtype=np.int64)
        L.call("tb_csr_download_pattern", self.h, rowptr, colidx, int(index_base))
        return rowptr, colidx

    def nonzeros(self) -> np.ndarray:
        vals = np.empty(self.nnz)
        L.call("tb_csr_values_download", self.h, vals)
        return vals

    def set_nonzeros(self, vals):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        if vals.size != self.nnz:
            raise ValueError("wrong number of nonzeros")
        L.call("tb_csr_values_upload", self.h, vals)

    def zero(self):
        L.call("tb_csr_zero", self.h)

    def axpby_values(self, M: "B200CSRMatrix", K: "B200CSRMatrix", dt: float):
        """nonzeros(self) .= nonzeros(M) .- dt .* nonzeros(K)   (euler.jl:104-116)."""
        L.call("tb_csr_axpby_values", self.h, M.h, K.h, float(dt))

    def mul(self, y: B200Vector, x: B200Vector, xcol=0, ycol=0):
        """mul!(y, A, x)."""
        L.call("tb_spmv", self.dev.h, self.h, x.h, int(xcol), y.h, int(ycol))

    def set_halo(self, neigh_ranks, send_ptr, send_rows, recv_ptr):
        nr = np.ascontiguousarray(neigh_ranks, dtype=np.int32)
        sp = np.ascontiguousarray(send_ptr, dtype=np.int64)
        sr = np.ascontiguousarray(send_rows, dtype=np.int64)
        rp = np.ascontiguousarray(recv_ptr, dtype=np.int64)
        L.call("tb_csr_set_halo", self.h, int(nr.size), L.ptr(nr), L.ptr(sp), L.ptr(sr), L.ptr(rp))

    def set_halo_peer(self, dst_off, dst_slot):
        do = np.ascontiguousarray(dst_off, dtype=np.int64)
        ds = np.ascontiguousarray(dst_slot, dtype=np.int32)
        L.call("tb_csr_set_halo_peer", self.h, L.ptr(do), L.ptr(ds))

    def halo_fused_capable(self) -> bool:
        v = C.c_int32()
        L.call("tb_csr_halo_fused_capable", self.h, C.byref(v))
        return bool(v.value)

    def set_halo_fused(self, on: bool):
        L.call("tb_csr_set_halo_fused", self.h, int(bool(on)))

    def free(self):
        if getattr(self, "h", None):
            L.lib().tb_csr_destroy(self.h)
            self.h = None


def quadrature(celltype, qorder):
    nq = C.c_int32()
    L.call("tb_quadrature", int(celltype), int(qorder), C.byref(nq), None, None)
    dim = 2 if celltype in (L.QUAD4, L.TRI3) else 3
    pts, w = np.empty((nq.value, dim)), np.empty(nq.value)
    L.call("tb_quadrature", int(celltype), int(qorder), C.byref(nq), L.ptr(pts), L.ptr(w))
    return pts, w


def assemble_mass(dev, mesh, M: B200CSRMatrix, qorder=2, rho=1.0):
    L.call("tb_assemble_mass", dev.h, mesh.h, int(qorder), float(rho), M.h)


def assemble_diffusion(dev, mesh, K: B200CSRMatrix, qorder, kind, data, cm_chi=1.0):
    data = np.ascontiguousarray(np.atleast_1d(np.asarray(data, dtype=np.float64)).ravel())
    L.call("tb_assemble_diffusion", dev.h, mesh.h, int(qorder), int(kind), data, data.size, float(cm_chi), K.h)


def assemble_source(dev, mesh, b: B200Vector, qorder, kind, prm, t, col=0):
    prm = np.ascontiguousarray(np.atleast_1d(np.asarray(prm, dtype=np.float64)).ravel())
    L.call("tb_assemble_source", dev.h, mesh.h, int(qorder), int(kind), L.ptr(prm), prm.size, float(t), b.h, int(col))


def assemble_source_program(dev, mesh, b: B200Vector, qorder, code, consts, t, col=0):
    """f(x, t) as a traced postfix program (api.trace_source): evaluated on the device at every quadrature point."""
    code = np.ascontiguousarray(code, dtype=np.int32)
    consts = np.ascontiguousarray(np.atleast_1d(np.asarray(consts, dtype=np.float64)).ravel())
    L.call("tb_assemble_source_program", dev.h, mesh.h, int(qorder), L.ptr(code), code.size, L.ptr(consts), consts.size,
           float(t), b.h, int(col))


def assemble_source_qp(dev, mesh, b: B200Vector, qorder, fq, col=0):
    fq = np.ascontiguousarray(fq, dtype=np.float64).ravel()
    L.call("tb_assemble_source_qp", dev.h, mesh.h, int(qorder), fq, b.h, int(col))


def ecg_plonsey(dev, mesh, qorder, kind, data, phi: B200Vector, electrodes, kappa_t, cm_chi=1.0, phicol=0):
    """update_ecg! + evaluate_ecg of Plonsey1964ECGGaussCache for all electrodes in one element sweep (ecg.jl:55-160)."""
    data = np.ascontiguousarray(np.atleast_1d(np.asarray(data, dtype=np.float64)).ravel())
    el = np.ascontiguousarray(np.atleast_2d(np.asarray(electrodes, dtype=np.float64)))
    if el.shape[1] != mesh.dim:
        raise ValueError(f"electrodes must have {mesh.dim} coordinates each")
    out = np.empty(el.shape[0])
    L.call("tb_ecg_plonsey", dev.h, mesh.h, int(qorder), int(kind), data, data.size, float(cm_chi), phi.h, int(phicol),
           el.reshape(-1), el.shape[0], float(kappa_t), out)
    return out


def cg_solve(dev, A: B200CSRMatrix, b: B200Vector, x: B200Vector, atol=SQRT_EPS, rtol=SQRT_EPS, itmax=None, bcol=0,
             xcol=0, precond=L.PRECOND_NONE):
    it, rn, cv = C.c_int64(), C.c_double(), C.c_int32()
    itmax = A.nrows if itmax is None else itmax
    L.call("tb_cg_solve_pc", dev.h, A.h, b.h, int(bcol), x.h, int(xcol), int(precond), float(atol), float(rtol), int(itmax),
           C.byref(it), C.byref(rn), C.byref(cv))
    return it.value, rn.value, bool(cv.value)


def cell_step(dev, model, params, u: B200Vector, t, dt, substeps=1, threshold=0.1, phi_idx=0, want_max=False):
    params = np.ascontiguousarray(params, dtype=np.float64)
    mx = C.c_double()
    L.call("tb_cell_step", dev.h, int(model), params, params.size, u.h, int(phi_idx), float(t), float(dt), int(substeps),
           float(threshold), C.byref(mx) if want_max else None)
    return mx.value if want_max else None


class MonodomainStepper:
    """tb_monodomain_*: the fused LieTrotterGodunov step."""

    def __init__(self, dev, M: B200CSRMatrix, K: B200CSRMatrix, model, params, phi_idx=0):
        params = np.ascontiguousarray(params, dtype=np.float64)
        h = C.c_void_p()
        L.call("tb_monodomain_create", dev.h, M.h, K.h, int(model), params, params.size, int(phi_idx), C.byref(h))
        self.h, self.dev, self.M, self.K = h, dev, M, K
        self._bS = None

    def set_cg(self, atol=SQRT_EPS, rtol=SQRT_EPS, itmax=None):
        L.call("tb_monodomain_set_cg", self.h, float(atol), float(rtol), int(self.M.nrows if itmax is None else itmax))

    def set_preconditioner(self, precond=L.PRECOND_NONE):
        L.call("tb_monodomain_set_preconditioner", self.h, int(precond))

    def set_cell_solver(self, substeps=1, threshold=0.1):
        L.call("tb_monodomain_set_cell_solver", self.h, int(substeps), float(threshold))

    def set_source(self, bS: B200Vector | None, col=0):
        self._bS = bS
        L.call("tb_monodomain_set_source", self.h, bS.h if bS is not None else None, int(col))

    def enable_timing(self, on=True):
        L.call("tb_monodomain_enable_timing", self.h, int(on))

    def section_ms(self):
        ms = (C.c_double * 3)()
        L.call("tb_monodomain_section_ms", self.h, ms)
        return list(ms)

    def step(self, u: B200Vector, t, dt):
        it, rn, cv = C.c_int64(), C.c_double(), C.c_int32()
        L.call("tb_monodomain_step", self.h, u.h, float(t), float(dt), C.byref(it), C.byref(rn), C.byref(cv))
        return it.value, rn.value, bool(cv.value)

    def step_rt(self, u: B200Vector, t, dt):
        """step + reaction tangent R (rtc.jl:51-78)"""
        it, rn, cv, R = C.c_int64(), C.c_double(), C.c_int32(), C.c_double()
        L.call("tb_monodomain_step_rt", self.h, u.h, float(t), float(dt), C.byref(it), C.byref(rn), C.byref(cv), C.byref(R))
        return it.value, rn.value, bool(cv.value), R.value

    def run(self, u: B200Vector, t0, dt, nsteps):
        it, cv = C.c_int64(), C.c_int32()
        L.call("tb_monodomain_run", self.h, u.h, float(t0), float(dt), int(nsteps), C.byref(it), C.byref(cv))
        return it.value, bool(cv.value)

    def run_host(self, u_dev: B200Vector, buf0: np.ndarray, buf1: np.ndarray, t0, dt, nsteps):
        """nsteps steps with the state in host memory between steps (step n: buf[n&1] -> buf[(n+1)&1]), transfers pipelined."""
        it, cv = C.c_int64(), C.c_int32()
        L.call("tb_monodomain_run_host", self.h, u_dev.h, L.ptr(buf0), L.ptr(buf1), float(t0), float(dt), int(nsteps),
               C.byref(it), C.byref(cv))
        return it.value, bool(cv.value)

    def set_host_chunks(self, nchunks: int):
        L.call("tb_monodomain_set_host_chunks", self.h, int(nchunks))

    def step_host(self, u_dev: B200Vector, u_in: np.ndarray, u_out: np.ndarray, t, dt):
        it, rn, cv = C.c_int64(), C.c_double(), C.c_int32()
        L.call("tb_monodomain_step_host", self.h, u_dev.h, L.ptr(u_in), L.ptr(u_out), float(t), float(dt), C.byref(it),
               C.byref(rn), C.byref(cv))
        return it.value, rn.value, bool(cv.value)

    def free(self):
        if getattr(self, "h", None):
            L.lib().tb_monodomain_destroy(self.h)
            self.h = None
