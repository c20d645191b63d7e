"""-m gpu: preconditioned CG beyond point Jacobi -- BlockJacobiPreconditioner(A, nblocks) with the semantics
of the reference's GPU example (bak/examples-gpu/spiral-wave.jl:95-105) and a Chebyshev polynomial preconditioner --
against the oracle's restatements (oracle.pcg + block_jacobi_preconditioner / chebyshev_preconditioner)."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pytestmark = pytest.mark.gpu


def _hex_problem(tb, dev, O, nel=(14, 12, 9), dt=1.0):
    lengths = tuple(0.25 * n for n in nel)
    md = tb.generate_mesh(tb.Hexahedron, nel, (0, 0, 0), lengths, device=dev)
    mo = O.generate_grid(O.HEX8, nel, (0, 0, 0), lengths)
    M = tb.B200CSRMatrix.from_mesh(dev, md)
    K, A = M.like(), M.like()
    D = np.diag([0.3, 0.1, 0.05])
    tb.core.assemble_mass(dev, md, M, 2, 1.0)
    tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_TENSOR, D, 1.0)
    A.axpby_values(M, K, dt)
    rp, ci = mo.pattern()
    Ao = O.axpby_values(O.assemble_mass(mo, 2), O.assemble_diffusion(mo, 2, O.D_TENSOR, D), dt)
    assert np.array_equal(A.nonzeros(), Ao)
    return md, mo, M, K, A, rp, ci, Ao


@pytest.mark.parametrize("degree,ratio", [(1, 30.0), (4, 30.0), (8, 30.0), (8, 100.0), (16, 300.0)])
def test_chebyshev_pcg_matches_oracle(tb, dev, oracle, degree, ratio):
    O = oracle
    md, mo, M, K, A, rp, ci, Ao = _hex_problem(tb, dev, O)
    b = np.random.default_rng(5).standard_normal(md.ndofs)
    bv, xv = tb.B200Vector.from_host(dev, b), tb.B200Vector(dev, md.ndofs)
    dev.cg_set_chebyshev(degree, ratio)
    it, rn, conv = tb.core.cg_solve(dev, A, bv, xv, precond=tb._lib.PRECOND_CHEBYSHEV)
    xo, ito, rno, convo = O.pcg(rp, ci, Ao, b, O.chebyshev_preconditioner(rp, ci, Ao, degree, ratio))
    assert conv and convo and abs(it - ito) <= 1, (it, ito)
    x = xv.to_host()
    assert np.abs(x - xo).max() <= 1e-6 * np.abs(xo).max()
    if it == ito:
        assert np.abs(x - xo).max() <= 1e-10 * np.abs(xo).max() and abs(rn - rno) <= 1e-6 * rno
    it0 = tb.core.cg_solve(dev, A, bv, xv)[0]
    itj = tb.core.cg_solve(dev, A, bv, xv, precond=tb._lib.PRECOND_JACOBI)[0]
    assert it <= itj <= it0 + 1 and (degree < 4 or it * 2 <= itj)          # the polynomial really cuts the iteration count


@pytest.mark.parametrize("nblocks,partition", [(1, "contiguous"), (7, "contiguous"), (40, "contiguous"), (25, "random"), (300, "contiguous")])
def test_block_jacobi_pcg_matches_oracle(tb, dev, oracle, nblocks, partition):
    O = oracle
    md, mo, M, K, A, rp, ci, Ao = _hex_problem(tb, dev, O, nel=(10, 9, 6))
    n = md.ndofs
    b = np.random.default_rng(6).standard_normal(n)
    bv, xv = tb.B200Vector.from_host(dev, b), tb.B200Vector(dev, n)
    rb = None if partition == "contiguous" else np.random.default_rng(7).integers(0, nblocks, n).astype(np.int32)
    dev.cg_set_block_jacobi(n, nblocks, rb)
    it, rn, conv = tb.core.cg_solve(dev, A, bv, xv, precond=tb._lib.PRECOND_BLOCK_JACOBI)
    xo, ito, rno, convo = O.pcg(rp, ci, Ao, b, O.block_jacobi_preconditioner(rp, ci, Ao, nblocks, rb))
    assert conv and convo and abs(it - ito) <= 1, (it, ito)
    assert np.abs(xv.to_host() - xo).max() <= 1e-6 * np.abs(xo).max()
    if nblocks == 1:
        assert it <= 2                                                       # the "preconditioner" is A^-1
    itj = tb.core.cg_solve(dev, A, bv, xv, precond=tb._lib.PRECOND_JACOBI)[0]
    if partition == "contiguous":
        assert it <= itj + 1


def test_block_jacobi_with_one_row_blocks_is_jacobi(tb, dev, oracle):
    O = oracle
    md, mo, M, K, A, rp, ci, Ao = _hex_problem(tb, dev, O, nel=(8, 7, 5))
    n = md.ndofs
    b = np.random.default_rng(8).standard_normal(n)
    bv, x1, x2 = tb.B200Vector.from_host(dev, b), tb.B200Vector(dev, n), tb.B200Vector(dev, n)
    dev.cg_set_block_jacobi(n, n, None)
    dev.cg_set_persistent(0)
    a = tb.core.cg_solve(dev, A, bv, x1, precond=tb._lib.PRECOND_BLOCK_JACOBI)
    bb = tb.core.cg_solve(dev, A, bv, x2, precond=tb._lib.PRECOND_JACOBI)
    dev.cg_set_persistent(1)
    assert a[0] == bb[0] and np.abs(x1.to_host() - x2.to_host()).max() <= 1e-12 * np.abs(x2.to_host()).max()


@pytest.mark.parametrize("precs", ["chebyshev", "block_jacobi"])
def test_stepper_with_general_preconditioner(tb, dev, oracle, precs):
    """the fused LieTrotterGodunov step with the preconditioner selected like `precs` selects it in the reference:
    phi within 1e-6 of the unpreconditioned oracle trajectory (both converge to sqrt(eps)), fewer iterations"""
    O = oracle
    nel, lengths = (12, 12, 6), (3.0, 3.0, 1.5)
    md = tb.generate_mesh(tb.Hexahedron, nel, (0, 0, 0), lengths, device=dev)
    M = tb.B200CSRMatrix.from_mesh(dev, md)
    K = M.like()
    D = np.diag([0.3, 0.1, 0.05])
    tb.core.assemble_mass(dev, md, M, 2, 1.0)
    tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_TENSOR, D, 1.0)
    ion = tb.FHNModel()
    N = md.ndofs
    x = md.dof_coords()
    u0 = np.concatenate([np.where(x[:, 0] <= 1.5, 1.0, 0.0), np.zeros(N)])
    res = {}
    for name in ("none", precs):
        st = tb.MonodomainStepper(dev, M, K, ion.model_id, ion.params())
        if name == "chebyshev":
            dev.cg_set_chebyshev(6, 30.0)
            st.set_preconditioner(tb._lib.PRECOND_CHEBYSHEV)
        elif name == "block_jacobi":
            dev.cg_set_block_jacobi(N, 24, None)
            st.set_preconditioner(tb._lib.PRECOND_BLOCK_JACOBI)
        u = tb.B200Vector.from_host(dev, u0, 2)
        its = [st.step(u, float(s), 1.0)[0] for s in range(10)]
        res[name] = (its, u.to_host())
        st.free()
        u.free()
    assert np.abs(res[precs][1][:N] - res["none"][1][:N]).max() <= 1e-6
    assert sum(res[precs][0]) < sum(res["none"][0])
