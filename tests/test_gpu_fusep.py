"""-m gpu: the fused-p SpMV experiment (TB_SPMV_FUSEP=1: p = r + beta p formed inside the SpMV's gather) produces the same bits
as the three-kernel iteration."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import thunderbolt_jl_b200 as tb
dev = tb.B200Device(0)
dev.cg_set_persistent(0)
md = tb.generate_mesh(tb.Hexahedron, (40, 33, 21), (0, 0, 0), (10.0, 8.25, 5.25), device=dev)
M = tb.B200CSRMatrix.from_mesh(dev, md); K = M.like(); A = M.like()
tb.core.assemble_mass(dev, md, M, 2, 1.0)
tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_TENSOR, np.diag([0.3, 0.1, 0.05]), 1.0)
A.axpby_values(M, K, 0.7)
b = np.random.default_rng(3).standard_normal(md.ndofs)
bv, xv = tb.B200Vector.from_host(dev, b), tb.B200Vector(dev, md.ndofs)
it, rn, conv = tb.core.cg_solve(dev, A, bv, xv)
np.save(sys.argv[1], np.concatenate([[it, rn, float(conv), dev.cg_last_path()], xv.to_host()]))
""" % str(ROOT)


def test_fused_p_spmv_is_bitwise_the_three_kernel_iteration(tmp_path):
    out = {}
    for flag in ("0", "1"):
        f = tmp_path / f"x{flag}.npy"
        r = subprocess.run([sys.executable, "-c", SCRIPT, str(f)], env=dict(os.environ, TB_SPMV_FUSEP=flag), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        out[flag] = np.load(f)
    assert out["0"][2] == 1.0 and out["0"][3] == 0.0 and out["0"][0] > 10
    assert np.array_equal(out["0"], out["1"])
