"""The C-ABI library loads and exports exactly what include/tbolt_b200.h declares (no compute calls:
there is no GPU here), and the product never touches the oracle."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(tb):
    L = tb._lib.lib()
    declared = tb.declared_symbols()
    assert len(declared) >= 60
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert L.tb_version() == 100


def test_binding_covers_every_declared_symbol(tb):
    bound = set(tb._lib._SIGNATURES) | {"tb_version", "tb_last_error"}
    assert set(tb.declared_symbols()) == bound


def test_every_declaration_cites_the_reference():
    text = (ROOT / "include" / "tbolt_b200.h").read_text()
    assert len(re.findall(r"[a-z_/\-A-Z0-9]+\.jl:\d+", text)) >= 40


def test_no_gpu_fails_loudly(tb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(tb.TBError) as e:
        tb.B200Device(0)
    assert e.value.status in (1, 2, 5)
    assert tb._lib.lib().tb_last_error()


def test_null_handles_are_rejected_not_dereferenced(tb):
    L = tb._lib.lib()
    assert L.tb_sync(None) == 1
    assert L.tb_vec_create(None, 10, 1, C.byref(C.c_void_p())) == 1
    assert L.tb_cg_solve(None, None, None, 0, None, 0, 0.0, 0.0, 1, None, None, None) == 1
    assert L.tb_monodomain_step(None, None, 0.0, 1.0, None, None, None) == 1
    assert b"NULL" in L.tb_last_error()
    nq = C.c_int32()
    assert L.tb_quadrature(1, 2, C.byref(nq), None, None) == 0 and nq.value == 8
    assert L.tb_quadrature(3, 7, C.byref(nq), None, None) == 5


def test_product_never_imports_the_oracle():
    pkg = ROOT / "thunderbolt.jl_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + [ROOT / "thunderbolt_jl_b200.py"]:
        text = p.read_text()
        assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), p
        assert "liboracle" not in text and "tb_oracle" not in text, p


def test_header_is_plain_c():
    """the drop-in boundary is a C ABI: the header must compile as C99 on its own (no C++ types, no missing includes)"""
    import subprocess
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                        str(ROOT / "include" / "tbolt_b200.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
