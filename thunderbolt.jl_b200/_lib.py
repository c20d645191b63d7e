"""ctypes binding of libtbolt_b200.so (the C ABI declared in include/tbolt_b200.h).

This is the same binding a Julia `ccall` layer would make; there is no CPU
fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("TB_LIB", PKG_DIR / "lib" / "libtbolt_b200.so"))   # TB_LIB: A/B runs against another build
HEADER = PKG_DIR.parent / "include" / "tbolt_b200.h"

TB_OK = 0
STATUS_NAMES = {0: "TB_OK", 1: "TB_ERR_INVALID", 2: "TB_ERR_CUDA", 3: "TB_ERR_NOMEM", 4: "TB_ERR_COMM",
                5: "TB_ERR_UNSUPPORTED"}

QUAD4, HEX8, TRI3, TET4 = 0, 1, 2, 3
FHN, PCG2019, ALIEV_PANFILOV = 0, 1, 2
D_SCALAR, D_TENSOR, D_SPECTRAL, D_CELL_TENSOR = 0, 1, 2, 3
PEER_BLOB_BYTES = 160
PRECOND_NONE, PRECOND_JACOBI, PRECOND_BLOCK_JACOBI, PRECOND_CHEBYSHEV = 0, 1, 2, 3
LAYOUT_STATE_BLOCKED, LAYOUT_POINT_BLOCKED = 0, 1
FACET_LINE2, FACET_QUAD4 = 0, 1


class CellBlock(C.Structure):
    """tb_cell_block"""
    _fields_ = [("offset", C.c_int64), ("npoints", C.c_int64), ("model", C.c_int32), ("layout", C.c_int32),
                ("nparams", C.c_int32), ("reserved", C.c_int32), ("params", C.c_double * 36)]


SRC_NONE, SRC_BOX, SRC_BALL, SRC_COSEXP, SRC_NORMT, SRC_ENDO = 0, 1, 2, 3, 4, 5


class TBError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", str(PKG_DIR / "csrc"), "clean"], check=True, capture_output=not verbose)
    r = subprocess.run(["make", "-C", str(PKG_DIR / "csrc"), "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libtbolt_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def declared_symbols() -> list[str]:
    """Every function name include/tbolt_b200.h declares."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tb_[a-z0-9_]+)\s*\(", text)))


_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_vp = C.c_void_p
_pp = C.POINTER(C.c_void_p)
_pi64 = C.POINTER(C.c_int64)
_pi32 = C.POINTER(C.c_int32)
_pf64 = C.POINTER(C.c_double)

_SIGNATURES = {
    "tb_ctx_create": [C.c_int32, _vp, _pp],
    "tb_ctx_destroy": [_vp],
    "tb_sync": [_vp],
    "tb_device_info": [_vp, _pi32, _pi64, _pi32, _pi32],
    "tb_timer_start": [_vp],
    "tb_timer_stop": [_vp, _pf64],
    "tb_launch_count": [_vp, _pi64],
    "tb_l2_flush": [_vp],
    "tb_profile_enable": [_vp, C.c_int32],
    "tb_profile_get": [_vp, _pf64, _pi64],
    "tb_comm_unique_id": [_vp],
    "tb_ctx_comm_init": [_vp, C.c_int32, C.c_int32, _vp],
    "tb_comm_barrier": [_vp],
    "tb_comm_allreduce_max": [_vp, _pf64],
    "tb_vec_create": [_vp, C.c_int64, C.c_int32, _pp],
    "tb_vec_destroy": [_vp],
    "tb_vec_sizes": [_vp, _pi64, _pi32],
    "tb_vec_upload": [_vp, _vp],
    "tb_vec_download": [_vp, _vp],
    "tb_vec_upload_col": [_vp, C.c_int32, _vp, C.c_int64, C.c_int64],
    "tb_vec_download_col": [_vp, C.c_int32, _vp, C.c_int64, C.c_int64],
    "tb_vec_fill": [_vp, C.c_int32, C.c_double],
    "tb_vec_copy": [_vp, C.c_int32, _vp, C.c_int32],
    "tb_vec_axpy": [_vp, C.c_int32, C.c_double, _vp, C.c_int32],
    "tb_vec_devptr": [_vp, C.c_int32, _pp, _pi64],
    "tb_mesh_create": [_vp, C.c_int32, C.c_int64, C.c_int64, _i64p, _f64p, _i64p, C.c_int64, C.c_int32, _pp],
    "tb_mesh_generate_grid": [_vp, C.c_int32, _i64p, _f64p, _f64p, _pp],
    "tb_mesh_destroy": [_vp],
    "tb_mesh_sizes": [_vp, _pi64, _pi64, _pi64, _pi32, _pi32],
    "tb_mesh_download": [_vp, _vp, _vp, _vp],
    "tb_mesh_dof_coords": [_vp, _f64p],
    "tb_mesh_extract_local": [_vp, C.c_int64, C.c_int64, _pp, _pi64],
    "tb_mesh_ghosts": [_vp, _i64p],
    "tb_mesh_generate_grid_local": [_vp, C.c_int32, _i64p, _f64p, _f64p, C.c_int64, C.c_int64, _pp, _pi64],
    "tb_mesh_set_ownership": [_vp, C.c_int64, C.c_int64, _vp, C.c_int64],
    "tb_csr_create": [_vp, C.c_int64, C.c_int64, _i64p, _i64p, C.c_int32, _pp],
    "tb_csr_create_from_mesh": [_vp, _vp, _pp],
    "tb_csr_create_like": [_vp, _pp],
    "tb_csr_destroy": [_vp],
    "tb_csr_sizes": [_vp, _pi64, _pi64, _pi64],
    "tb_csr_storage": [_vp, _pi64, _pi64, _pi32],
    "tb_csr_download_pattern": [_vp, _i64p, _i64p, C.c_int32],
    "tb_csr_values_download": [_vp, _f64p],
    "tb_csr_values_upload": [_vp, _f64p],
    "tb_csr_zero": [_vp],
    "tb_csr_axpby_values": [_vp, _vp, _vp, C.c_double],
    "tb_spmv": [_vp, _vp, _vp, C.c_int32, _vp, C.c_int32],
    "tb_csr_set_halo": [_vp, C.c_int32, _vp, _vp, _vp, _vp],
    "tb_peer_export": [_vp, C.c_int64, _vp],
    "tb_peer_attach": [_vp, _vp, C.c_int32],
    "tb_peer_enabled": [_vp, _pi32],
    "tb_csr_set_halo_peer": [_vp, _vp, _vp],
    "tb_peer_stats": [_vp, _pf64, _pi64, _pf64, _pi64, C.c_int32],
    "tb_csr_halo_fused_capable": [_vp, _pi32],
    "tb_csr_set_halo_fused": [_vp, C.c_int32],
    "tb_quadrature": [C.c_int32, C.c_int32, _pi32, _vp, _vp],
    "tb_assemble_mass": [_vp, _vp, C.c_int32, C.c_double, _vp],
    "tb_assemble_diffusion": [_vp, _vp, C.c_int32, C.c_int32, _f64p, C.c_int64, C.c_double, _vp],
    "tb_assemble_source": [_vp, _vp, C.c_int32, C.c_int32, _vp, C.c_int32, C.c_double, _vp, C.c_int32],
    "tb_assemble_source_qp": [_vp, _vp, C.c_int32, _f64p, _vp, C.c_int32],
    "tb_assemble_source_program": [_vp, _vp, C.c_int32, _vp, C.c_int32, _vp, C.c_int32, C.c_double, _vp, C.c_int32],
    "tb_assembly_set_mode": [_vp, C.c_int32],
    "tb_assembly_info": [_vp, _pi32, _pi32, _pi32],
    "tb_assembly_set_scratch_budget": [_vp, C.c_int64],
    "tb_assembly_release_scratch": [_vp],
    "tb_cg_solve": [_vp, _vp, _vp, C.c_int32, _vp, C.c_int32, C.c_double, C.c_double, C.c_int64, _pi64, _pf64, _pi32],
    "tb_ecg_plonsey": [_vp, _vp, C.c_int32, C.c_int32, _f64p, C.c_int64, C.c_double, _vp, C.c_int32, _f64p, C.c_int32, C.c_double,
                       _f64p],
    "tb_cg_solve_pc": [_vp, _vp, _vp, C.c_int32, _vp, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int64, _pi64, _pf64, _pi32],
    "tb_monodomain_set_preconditioner": [_vp, C.c_int32],
    "tb_cg_set_persistent": [_vp, C.c_int32],
    "tb_cg_set_persistent_variant": [_vp, C.c_int32],
    "tb_cg_last_path": [_vp, _pi32],
    "tb_cg_set_exact_dot": [_vp, C.c_int32],
    "tb_cg_set_block_jacobi": [_vp, C.c_int64, C.c_int64, _vp],
    "tb_cg_set_chebyshev": [_vp, C.c_int32, C.c_double],
    "tb_cell_step": [_vp, C.c_int32, _f64p, C.c_int32, _vp, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_double,
                     _pf64],
    "tb_index_create": [_vp, _vp, C.c_int64, C.c_int32, _pp],
    "tb_index_destroy": [_vp],
    "tb_vec_gather": [_vp, C.c_int32, _vp, C.c_int32, _vp],
    "tb_vec_scatter": [_vp, C.c_int32, _vp, _vp, C.c_int32],
    "tb_cell_step_blocks": [_vp, _vp, C.c_int32, _vp, C.c_double, C.c_double, C.c_int32, C.c_double, _pf64],
    "tb_assemble_interface_diffusion": [_vp, C.c_int32, C.c_int32, C.c_int64, _vp, C.c_int32, _vp, _vp, C.c_int32, C.c_double, _vp],
    "tb_vec_dots": [_vp, _vp, _vp, C.c_int32, _f64p],
    "tb_csr_diagonal": [_vp, _vp, C.c_int32],
    "tb_csr_apply_zero": [_vp, _vp, C.c_double],
    "tb_vec_fill_at": [_vp, C.c_int32, _vp, C.c_double],
    "tb_host_alloc": [C.c_int64, _pp],
    "tb_host_free": [_vp],
    "tb_vec_stage_col": [_vp, C.c_int32, _vp],
    "tb_stage_wait": [_vp],
    "tb_monodomain_create": [_vp, _vp, _vp, C.c_int32, _f64p, C.c_int32, C.c_int32, _pp],
    "tb_monodomain_destroy": [_vp],
    "tb_monodomain_set_cg": [_vp, C.c_double, C.c_double, C.c_int64],
    "tb_monodomain_set_cell_solver": [_vp, C.c_int32, C.c_double],
    "tb_monodomain_set_source": [_vp, _vp, C.c_int32],
    "tb_monodomain_step": [_vp, _vp, C.c_double, C.c_double, _pi64, _pf64, _pi32],
    "tb_monodomain_step_rt": [_vp, _vp, C.c_double, C.c_double, _pi64, _pf64, _pi32, _pf64],
    "tb_monodomain_run": [_vp, _vp, C.c_double, C.c_double, C.c_int64, _pi64, _pi32],
    "tb_monodomain_run_host": [_vp, _vp, _vp, _vp, C.c_double, C.c_double, C.c_int64, _pi64, _pi32],
    "tb_monodomain_set_host_chunks": [_vp, C.c_int32],
    "tb_monodomain_step_host": [_vp, _vp, _vp, _vp, C.c_double, C.c_double, _pi64, _pf64, _pi32],
    "tb_monodomain_section_ms": [_vp, _pf64],
    "tb_monodomain_enable_timing": [_vp, C.c_int32],
}

_lib = None


def lib():
    """The loaded shared library.  Raises if it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C thunderbolt.jl_b200/csrc`).  There is no CPU fallback.")
        L = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
        L.tb_version.restype = C.c_int32
        L.tb_last_error.restype = C.c_char_p
        for name, args in _SIGNATURES.items():
            if not hasattr(L, name) and "TB_LIB" in os.environ:
                continue                      # an older build under test: calls to what it lacks fail loudly
            f = getattr(L, name)
            f.argtypes = args
            f.restype = C.c_int32
        _lib = L
    return _lib


def check(status: int):
    if status != TB_OK:
        raise TBError(status, lib().tb_last_error().decode(errors="replace"))


def call(name: str, *args):
    check(getattr(lib(), name)(*args))


def ptr(a):
    """void* of a numpy array (or None)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)
