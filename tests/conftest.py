import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no GPU in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def tb():
    """The product package.  The shared library is a build artefact (git-ignored); a fresh checkout that has not run
    __graft_entry__.build() yet gets it built here (nvcc cross-compiles without a GPU) -- the product itself never
    builds or falls back at import time."""
    import thunderbolt_jl_b200 as tb
    if not tb._lib.LIB_PATH.exists():
        tb._lib.build()
    return tb


@pytest.fixture(scope="session")
def dev(tb):
    d = tb.B200Device(0)
    tb.set_default_device(d)
    yield d
    d.close()


@pytest.fixture(scope="session")
def hostmath():
    """The product's host/device inline arithmetic compiled for the host (tests/hostmath)."""
    d = ROOT / "tests" / "hostmath"
    so = d / "libhostmath.so"
    srcs = [d / "hostmath.cpp", ROOT / "thunderbolt.jl_b200" / "csrc" / "tb_cells.cuh",
            ROOT / "thunderbolt.jl_b200" / "csrc" / "tb_elements.cuh", ROOT / "thunderbolt.jl_b200" / "csrc" / "tb_grid.cuh"]
    if not so.exists() or so.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas",
                        "-shared", "-o", str(so), str(srcs[0])], check=True)
    L = C.CDLL(str(so))
    f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L.hm_cell_node_step.argtypes = [C.c_int, C.c_int, f64, f64, C.c_double, C.c_double, C.c_int, C.c_double]
    L.hm_cell_node_step.restype = C.c_double
    L.hm_tables.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), f64, f64, f64, f64]
    L.hm_element_matrix.argtypes = [C.c_int, C.c_int, C.c_int, f64, C.c_double, C.c_int, f64, C.c_double, C.c_int64, f64]
    L.hm_element_diffusion_full.argtypes = [C.c_int, C.c_int, f64, C.c_int, f64, C.c_double, C.c_int64, f64]
    L.hm_grid_dofs.argtypes = [C.c_int, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS"),
                               np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")]
    L.hm_element_source.argtypes = [C.c_int, C.c_int, f64, C.c_int, f64, C.c_double, C.c_void_p, f64]
    L.hm_program_eval.argtypes = [C.c_int, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS"), C.c_int, f64, C.c_int,
                                  f64, C.c_int, C.c_double, f64]
    return L


# the distorted hexahedron of the reference's geometry fixture (test/test_coefficients.jl:243-252)
DISTORTED_HEX = np.array([[0.0, 0.0, 0.0], [1.3, 0.1, 0.0], [1.1, 1.4, -0.2], [0.2, 1.0, 0.1], [-0.1, 0.2, 1.2],
                          [1.5, 0.0, 1.0], [1.2, 1.1, 1.4], [0.0, 1.3, 1.1]])
