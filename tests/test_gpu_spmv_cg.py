"""-m gpu: SpMV, A = M - dt K, and CG through the C ABI vs the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["persistent", "persistent_tma", "multikernel"], autouse=True)
def cg_path(request, dev):
    """Every test of this module runs with all three CG execution models: the one-launch cooperative kernels that small
    (register-resident) and mid-size (TMA sweep) operators get by default (tb_cg_small.cu) and the
    three-kernels-per-iteration path large and multi-GPU ones take (tb_cg.cu)."""
    dev.cg_set_persistent({"persistent": 1, "persistent_tma": 2, "multikernel": 0}[request.param])
    yield request.param
    dev.cg_set_persistent(1)


def _system(tb, dev, O, ct, nel, dt, kappa):
    dim = len(nel)
    mo = O.generate_grid(ct, nel, (0.0,) * dim, tuple(0.25 * n for n in nel))
    md = tb.DeviceMesh.from_host(dev, ct, mo.conn, mo.coords, mo.celldofs, mo.ndofs)
    M = tb.B200CSRMatrix.from_mesh(dev, md)
    K, A = M.like(), M.like()
    Mo, Ko = O.assemble_mass(mo, 2), O.assemble_diffusion(mo, 2, O.D_TENSOR, np.diag(kappa[:dim]))
    # upload the oracle's values so that SpMV/CG are compared on IDENTICAL operators
    M.set_nonzeros(Mo); K.set_nonzeros(Ko)
    A.axpby_values(M, K, dt)
    Ao = O.axpby_values(Mo, Ko, dt)
    return mo, md, M, K, A, Mo, Ko, Ao


@pytest.mark.parametrize("ct,nel", [(0, (31, 17)), (1, (9, 8, 7)), (3, (6, 5, 4)), (2, (12, 9))])
def test_axpby_and_spmv_bitwise(tb, dev, oracle, ct, nel):
    O = oracle
    mo, md, M, K, A, Mo, Ko, Ao = _system(tb, dev, O, ct, nel, 0.37, (0.13, 0.02, 0.02))
    assert np.array_equal(A.nonzeros(), Ao)                           # a - dt*b without fma: bitwise
    rp, ci = mo.pattern()
    x = np.random.default_rng(0).standard_normal(mo.ndofs)
    xd, yd = tb.B200Vector.from_host(dev, x), tb.B200Vector(dev, mo.ndofs)
    for mat, vals in ((A, Ao), (M, Mo), (K, Ko)):
        mat.mul(yd, xd)
        assert np.array_equal(yd.to_host(), O.spmv(rp, ci, vals, x))  # sequential row sums, no fma: bitwise
    with pytest.raises(tb.TBError):
        A.mul(xd, xd)                                                 # aliasing refused
    for h in (M, K, A, xd, yd, md):
        h.free()


@pytest.mark.parametrize("ct,nel,dt", [(0, (40, 40), 1.0), (1, (12, 10, 8), 0.5), (3, (6, 6, 6), 0.5), (1, (16, 16, 4), 0.01)])
def test_cg_matches_oracle(tb, dev, oracle, ct, nel, dt):
    O = oracle
    mo, md, M, K, A, Mo, Ko, Ao = _system(tb, dev, O, ct, nel, dt, (0.13, 0.02, 0.02))
    rp, ci = mo.pattern()
    rng = np.random.default_rng(1)
    u = rng.standard_normal(mo.ndofs)
    b = O.spmv(rp, ci, Mo, u)
    bd, xd = tb.B200Vector.from_host(dev, b), tb.B200Vector(dev, mo.ndofs)
    xd.fill(123.0)                                                    # x0 = 0 regardless of what x holds
    for atol, rtol in ((O.SQRT_EPS, O.SQRT_EPS), (1e-6, 1e-5), (1e-14, 1e-14)):
        xo, ito, rno, convo = O.cg(rp, ci, Ao, b, atol, rtol)
        it, rn, conv = tb.core.cg_solve(dev, A, bd, xd, atol, rtol)
        assert conv == convo and abs(it - ito) <= 1, (it, ito)       # north star: iteration counts within +-1
        if it == ito:
            assert rn == pytest.approx(rno, rel=1e-6)
            x = xd.to_host()
            assert np.abs(x - xo).max() <= 1e-12 * np.abs(xo).max()
    # run-to-run determinism of the two-stage reductions
    it1, rn1, _ = tb.core.cg_solve(dev, A, bd, xd)
    x1 = xd.to_host()
    it2, rn2, _ = tb.core.cg_solve(dev, A, bd, xd)
    assert it1 == it2 and rn1 == rn2 and np.array_equal(x1, xd.to_host())
    for h in (M, K, A, bd, xd, md):
        h.free()


def test_cg_failure_is_a_flag_not_an_error(tb, dev, oracle):
    """Non-convergence must come back as converged = False (-> ReturnCode.MaxIters, euler.jl:95-100)."""
    O = oracle
    mo, md, M, K, A, Mo, Ko, Ao = _system(tb, dev, O, 0, (30, 30), 50.0, (0.13, 0.02, 0.02))
    rp, ci = mo.pattern()
    b = np.sin(np.arange(mo.ndofs) * 0.1)
    bd, xd = tb.B200Vector.from_host(dev, b), tb.B200Vector(dev, mo.ndofs)
    for itmax in (0, 1, 3):
        xo, ito, rno, convo = O.cg(rp, ci, Ao, b, itmax=itmax)
        it, rn, conv = tb.core.cg_solve(dev, A, bd, xd, itmax=itmax)
        assert (it, conv) == (ito, convo) == (itmax, False)
        assert np.abs(xd.to_host() - xo).max() <= 1e-13 * max(np.abs(xo).max(), 1.0)
        assert rn == pytest.approx(rno, rel=1e-10)
    # zero right-hand side: solved at iteration 0
    bd.fill(0.0)
    it, rn, conv = tb.core.cg_solve(dev, A, bd, xd)
    assert (it, rn, conv) == (0, 0.0, True) and not xd.to_host().any()
    # NaN in the operator: no hang, no error status, just not converged
    bad = Ao.copy(); bad[5] = np.nan
    A.set_nonzeros(bad)
    bd.upload(b)
    it, rn, conv = tb.core.cg_solve(dev, A, bd, xd, itmax=20)
    assert not conv and it == 20
    for h in (M, K, A, bd, xd, md):
        h.free()


def test_cg_path_selection(tb, dev, cg_path):
    """Small operators take the one-launch cooperative kernel (when enabled), operators with more than
    148*16*32*4 rows never do; both paths report the same iteration count on the same system."""
    def solve(nel):
        md = tb.generate_mesh(tb.Quadrilateral, nel, (0.0, 0.0), (1.0, 1.0), device=dev)
        M = tb.B200CSRMatrix.from_mesh(dev, md)
        K, A = M.like(), M.like()
        tb.core.assemble_mass(dev, md, M, 2, 1.0)
        tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_SCALAR, [1e-5], 1.0)
        A.axpby_values(M, K, 1.0)
        x = md.dof_coords()
        b = tb.B200Vector.from_host(dev, np.sin(7 * x[:, 0]) * np.cos(5 * x[:, 1]))
        xs = tb.B200Vector(dev, md.ndofs)
        it, rn, conv = tb.core.cg_solve(dev, A, b, xs)
        path = dev.cg_last_path()
        y = tb.B200Vector(dev, md.ndofs)
        A.mul(y, xs)
        res = np.linalg.norm(y.to_host() - b.to_host()) / np.linalg.norm(b.to_host())
        for h in (M, K, A, b, xs, y, md):
            h.free()
        return it, conv, path, res
    it, conv, path, res = solve((200, 150))
    assert conv and path == {"persistent": 1, "persistent_tma": 2, "multikernel": 0}[cg_path] and res < 1e-7
    it2, conv2, path2, res2 = solve((900, 900))        # 811 801 rows: beyond the register-resident limit -> TMA kernel
    assert conv2 and path2 == (0 if cg_path == "multikernel" else 2) and res2 < 1e-7


@pytest.mark.parametrize("ct,nel,dt", [(0, (40, 40), 1.0), (1, (12, 10, 8), 0.5), (3, (6, 6, 6), 0.5)])
def test_jacobi_pcg_matches_oracle(tb, dev, oracle, ct, nel, dt, cg_path):
    """KrylovJL_CG with a Jacobi preconditioner (z = D^-1 r, stop on sqrt(r.z)) against the oracle's
    restatement, on warped meshes where the diagonal really varies; fewer iterations than plain CG."""
    O = oracle
    mo, md, M, K, A, Mo, Ko, Ao = _system(tb, dev, O, ct, nel, dt, (0.13, 0.02, 0.02))
    rp, ci = mo.pattern()
    rng = np.random.default_rng(3)
    b = O.spmv(rp, ci, Mo, rng.standard_normal(mo.ndofs))
    bd, xd = tb.B200Vector.from_host(dev, b), tb.B200Vector(dev, mo.ndofs)
    for atol, rtol in ((O.SQRT_EPS, O.SQRT_EPS), (1e-12, 1e-12)):
        xo, ito, rno, convo = O.pcg_jacobi(rp, ci, Ao, b, atol, rtol)
        it, rn, conv = tb.core.cg_solve(dev, A, bd, xd, atol, rtol, precond=tb._lib.PRECOND_JACOBI)
        assert dev.cg_last_path() == {"persistent": 1, "persistent_tma": 2, "multikernel": 0}[cg_path]
        assert conv == convo and abs(it - ito) <= 1, (it, ito)
        if it == ito:
            assert rn == pytest.approx(rno, rel=1e-6)
            assert np.abs(xd.to_host() - xo).max() <= 1e-11 * np.abs(xo).max()
    it_plain, _, _ = tb.core.cg_solve(dev, A, bd, xd, 1e-12, 1e-12)
    assert it <= it_plain + 2      # (near-)uniform diagonals: Jacobi must at least not hurt
    for h in (M, K, A, bd, xd, md):
        h.free()


@pytest.mark.parametrize("ct,nel", [(0, (64, 48)), (0, (256, 256)), (1, (50, 50, 50)), (1, (64, 64, 60))])
def test_persistent_variants_bitwise(tb, dev, oracle, ct, nel, cg_path):
    """The register-resident persistent CG with two flag barriers per iteration (direction of the gathered columns formed on the
    fly) against the three-grid.sync kernel: same partial-sum order, same unfused p = z + beta p, hence the same bits -- x,
    iteration count and residual norm -- at 1, 2 and 4 rows per lane, plain and Jacobi-preconditioned, from b and fused
    with b = M u."""
    if cg_path != "persistent":
        pytest.skip("register-resident kernels only")
    md = tb.generate_mesh([tb.Quadrilateral, tb.Hexahedron][ct], nel, (0.0,) * len(nel), tuple(0.25 * n for n in nel), device=dev)
    M = tb.B200CSRMatrix.from_mesh(dev, md)
    K, A = M.like(), M.like()
    tb.core.assemble_mass(dev, md, M, 2, 1.0)
    tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_TENSOR, np.diag([0.13, 0.02, 0.02][:len(nel)]), 1.0)
    A.axpby_values(M, K, 0.7)
    n = md.ndofs
    rng = np.random.default_rng(4)
    bd, xd = tb.B200Vector.from_host(dev, rng.standard_normal(n)), tb.B200Vector(dev, n)
    res = {}
    for variant in (1, 2):
        dev.cg_set_persistent_variant(variant)
        for precond in (tb._lib.PRECOND_NONE, tb._lib.PRECOND_JACOBI):
            for tol in ((oracle.SQRT_EPS, oracle.SQRT_EPS), (1e-13, 1e-13)):
                it, rn, conv = tb.core.cg_solve(dev, A, bd, xd, *tol, precond=precond)
                assert conv and dev.cg_last_path() == 1
                res[(variant, precond, tol)] = (it, rn, xd.to_host().copy())
        ion = tb.ParametrizedFHNModel()
        st = tb.MonodomainStepper(dev, M, K, ion.model_id, ion.params())
        u = tb.B200Vector.from_host(dev, np.concatenate([rng.uniform(0, 1, n), np.zeros(n)]) if variant == 1 else res["u0"], 2)
        if variant == 1:
            res["u0"] = u.to_host().copy()
        its = [st.step(u, float(k), 0.7)[0] for k in range(5)]
        res[(variant, "step")] = (its, u.to_host().copy())
        st.free(); u.free()
    dev.cg_set_persistent_variant(2)
    for key in [k for k in res if isinstance(k, tuple) and k[0] == 1 and k[1] != "step"]:
        a, b = res[key], res[(2,) + key[1:]]
        assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]), key
    assert res[(1, "step")][0] == res[(2, "step")][0] and np.array_equal(res[(1, "step")][1], res[(2, "step")][1])
    for h in (M, K, A, bd, xd, md):
        h.free()
