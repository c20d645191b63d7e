"""The reference's own known-answer checks for the path, run against the oracle.
Each test names the reference test it restates."""
import numpy as np
import pytest

from conftest import DISTORTED_HEX


def test_backward_euler_steady_state(oracle):
    """test/test_time_integrator.jl:13-41: 4x4 Q1 grid, D = I, pure Neumann, u = 1 is steady."""
    O = oracle
    m = O.generate_grid(O.QUAD4, (4, 4), (0.0, 0.0), (1.0, 1.0))
    rp, ci = m.pattern()
    M = O.assemble_mass(m, 2)
    K = O.assemble_diffusion(m, 2, O.D_TENSOR, [1.0, 0.0, 0.0, 1.0])
    A = O.axpby_values(M, K, 0.1)
    u = np.ones(m.ndofs)
    b = O.spmv(rp, ci, M, u)
    u, it, rn, conv = O.cg(rp, ci, A, b)
    assert conv and np.allclose(u, 1.0, atol=1e-4)
    for _ in range(9):                       # solve! to t = 1 with dt = 0.1
        b = O.spmv(rp, ci, M, u)
        u, it, rn, conv = O.cg(rp, ci, A, b)
        assert conv
    assert np.allclose(u, 1.0, atol=1e-4)


def test_substepper_evaluates_each_substep_at_its_own_time(oracle):
    """test/test_time_integrator.jl:275-294: closed form with t0 = 2, dt = 0.4, 4 substeps."""
    O = oracle
    t0, dt, sub = 2.0, 0.4, 4
    u = np.zeros(1)
    # threshold 0 forces the substepping branch regardless of |du| (the reference's default 0.1 is also exceeded)
    O.cell_step(O.TIMEPROBE, [0.0], u, 1, t0, dt, substeps=sub, threshold=0.1)
    dts = dt / sub
    assert u[0] == pytest.approx(sum(dts * (1.0 + np.sin(t0 + s * dts)) for s in range(sub)), rel=1e-15)


def test_pcg2019_default_initial_state(oracle):
    """src/modeling/cells/pcg2019.jl:137-152 / test/test_solution_variables.jl:100-111."""
    O = oracle
    p = O.default_params(O.PCG2019)
    u0 = O.default_initial_state(O.PCG2019)
    sig = lambda phi, E, k, s: 1.0 / (1.0 + np.exp(s * (phi - E) / k))
    assert u0[0] == -85.0
    expect = [sig(-85.0, -78.7, 5.93, 1.0), sig(-85.0, -52.244, 6.5472, -1.0), sig(-85.0, -15.7, 4.6, 1.0),
              sig(-85.0, -47.9286, 4.9314, 1.0), sig(-85.0, 24.6, 12.1, -1.0), sig(-85.0, -26.6, 6.5, -1.0)]
    assert np.allclose(u0[1:], expect, rtol=1e-15)
    # resting state: gates sit at their infinity values, so only phi moves
    du = O.cell_rhs(O.PCG2019, p, u0)
    assert np.all(np.abs(du[1:]) < 1e-12)


def test_fhn_rhs_values(oracle):
    """src/modeling/cells/fhn.jl:21-34 with defaults a=.1 b=.5 c=1 d=0 e=.01 f=1."""
    O = oracle
    du = O.cell_rhs(O.FHN, O.default_params(O.FHN), [0.5, 0.2])
    assert du[0] == pytest.approx(0.5 * 0.5 * 0.4 - 0.2)
    assert du[1] == pytest.approx(0.01 * (0.25 - 0.2))


def test_layout_is_state_blocked(oracle):
    """test/test_solution_variables.jl:113-127: state s of point i lives at u[s*N + i]."""
    O = oracle
    n = 5
    u = np.zeros(2 * n)
    u[:n] = np.linspace(0.0, 1.0, n)          # phi block
    u[n:] = 0.1                                # s block
    ref = u.copy()
    O.cell_step(O.FHN, O.default_params(O.FHN), u, n, 0.0, 0.5)
    for i in range(n):
        du = O.cell_rhs(O.FHN, O.default_params(O.FHN), [ref[i], ref[n + i]])
        assert u[i] == ref[i] + 0.5 * du[0] and u[n + i] == ref[n + i] + 0.5 * du[1]


def test_forward_euler_and_adaptive_agree_but_differ(oracle):
    """test/integration/test_electrophysiology.jl:76-95: 8x8 quads on [-2.5,2.5]^2, FHN, stimulus
    |x| < 0.1 && t < 2 -> 0.01, initial condition phi0 = max(1 - |x|, 0) (simple_initializer!, :8-27),
    tspan (0,10), dt = 1: FE and adaptive agree to 1e-2 and differ at 1e-8."""
    O = oracle
    m = O.generate_grid(O.QUAD4, (8, 8), (-2.5, -2.5), (2.5, 2.5))
    M = O.assemble_mass(m, 2)
    K = O.assemble_diffusion(m, 2, O.D_TENSOR, [4.5e-4, 0, 0, 2.0e-4])
    x = m.dof_coords
    res = []
    for sub in (1, 10):
        orc = O.MonodomainOracle(m, O.FHN, O.default_params(O.FHN), M, K, substeps=sub)
        u = np.zeros(2 * m.ndofs)
        u[:m.ndofs] = np.maximum(1.0 - np.linalg.norm(x, axis=1), 0.0)
        for s in range(10):
            t = float(s)
            if 0.0 <= t + 1.0 <= 2.1:
                orc.bS = O.assemble_source(m, 2, O.SRC_BALL, [0.1, 2.0, 0.01], t + 1.0)
            it, rn, conv = orc.step(u, t, 1.0)
            assert conv
        res.append(u.copy())
    assert np.allclose(res[0], res[1], rtol=1e-2, atol=1e-2 * np.abs(res[0]).max())
    assert not np.allclose(res[0], res[1], rtol=1e-8, atol=0)


def test_forward_euler_and_adaptive_agree_on_the_lv(oracle):
    """test/integration/test_electrophysiology.jl:101-121: the same wave-propagation case on generate_ideal_lv_mesh(4, 1, 1),
    kappa = (4.5e-4, 2e-4, 2e-4), where the reference asserts FE ~ adaptive to rtol 1e-4 (isapprox on the whole solution
    vector: |u - v| <= rtol max(|u|, |v|) in the 2-norm).  The reference's mesh is hexahedra + wedges; here its conforming
    tetrahedral split (this project's addition, lv.tetrahedralize) carries the same nodes."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from thunderbolt_jl_b200 import lv
    O = oracle
    nodes, hexes, wedges, prm = lv.generate_ideal_lv_mesh(4, 1, 1)
    tets = lv.tetrahedralize(nodes, hexes, wedges)
    m = O.Mesh(O.TET4, tets, nodes)
    M = O.assemble_mass(m, 2)
    K = O.assemble_diffusion(m, 2, O.D_TENSOR, np.diag([4.5e-4, 2.0e-4, 2.0e-4]).ravel())
    x = m.dof_coords
    res = []
    for sub in (1, 10):
        orc = O.MonodomainOracle(m, O.FHN, O.default_params(O.FHN), M, K, substeps=sub)
        u = np.zeros(2 * m.ndofs)
        u[:m.ndofs] = np.maximum(1.0 - np.linalg.norm(x, axis=1), 0.0)            # simple_initializer!, :8-27
        u0 = u.copy()
        for s in range(10):
            t = float(s)
            if 0.0 <= t + 1.0 <= 2.1:
                orc.bS = O.assemble_source(m, 2, O.SRC_BALL, [0.1, 2.0, 0.01], t + 1.0)
            it, rn, conv = orc.step(u, t, 1.0)
            assert conv
        assert not np.allclose(u, u0)                                               # `integrator.u ≉ u₀`, :60
        res.append(u.copy())
    a, b = res
    assert np.linalg.norm(a - b) <= 1e-4 * max(np.linalg.norm(a), np.linalg.norm(b))


def test_cartesian_coordinate_and_analytical_coefficient_values(oracle, hostmath):
    """test/test_coefficients.jl:73-105: on generate_grid(Line, (2,)) the Cartesian coordinate at xi = 0 / 0.1 is -0.5 / -0.45 in
    cell 1 and 0.5 / 0.55 in cell 2, and AnalyticalCoefficient((x, t) -> norm(x) + t) is 0.5 / 0.45 / 0.5 / 0.55 (+ t).  This
    library has no 1-D cells: the same numbers on the x coordinate of the two-quadrilateral strip [-1, 1] x [0, 1] (first
    reference coordinate = the Line's), through the oracle's shape functions (x_q = sum_a N_a x_a, coefficients.jl:279-292)
    and through the product's traced-program evaluator for f."""
    from thunderbolt_jl_b200 import trace as T
    O = oracle
    m = O.generate_grid(O.QUAD4, (2, 1), (-1.0, 0.0), (1.0, 1.0))
    want_x = {(0, 0.0): -0.5, (0, 0.1): -0.45, (1, 0.0): 0.5, (1, 0.1): 0.55}
    prog = T.trace_source(lambda x, t: abs(x[0]) + t, 2)                       # norm of the Line's 1-vector
    assert prog is not None
    for (cell, xi1), xw in want_x.items():
        for eta in (-0.3, 0.0, 0.7):
            N, _ = O.shape(O.QUAD4, [xi1, eta])
            xq = N @ m.coords[m.conn[cell]]
            assert np.isclose(xq[0], xw, rtol=0, atol=1e-15)
            for t in (0.0, 1.0):
                out = np.zeros(1)
                rc = hostmath.hm_program_eval(2, prog.code, len(prog), np.ascontiguousarray(prog.consts) if prog.consts.size else np.zeros(1),
                                              prog.consts.size, np.ascontiguousarray(xq), 1, t, out)
                assert rc == 0 and np.isclose(out[0], abs(xw) + t, rtol=0, atol=1e-15)


def test_distorted_hex_geometry(oracle):
    """test/test_coefficients.jl:239-279 compares two implementations of the mapping on this fixture;
    here: partition of unity, gradient of a linear field, detJ*w sums to the volume (independent facts)."""
    O = oracle
    pts, w = O.quadrature(O.HEX8, 2)
    assert pts.shape == (8, 3) and np.allclose(w, 1.0)
    vol = 0.0
    lin = DISTORTED_HEX @ np.array([0.3, -1.2, 0.7]) + 2.0      # u(x) = a.x + c at the vertices
    for q in range(8):
        det, N, G = O.map_qp(O.HEX8, DISTORTED_HEX, pts[q])
        assert det > 0
        assert N.sum() == pytest.approx(1.0, abs=1e-15)
        assert np.allclose(G.sum(axis=0), 0.0, atol=1e-14)
        assert np.allclose(lin @ G, [0.3, -1.2, 0.7], atol=1e-13)    # isoparametric map reproduces linear fields
        assert np.allclose(N @ DISTORTED_HEX, _trilinear(DISTORTED_HEX, pts[q]), atol=1e-15)
        vol += det * w[q]
    # volume by a much finer rule
    p4, w4 = O.quadrature(O.HEX8, 4)
    vol4 = sum(O.map_qp(O.HEX8, DISTORTED_HEX, p)[0] * ww for p, ww in zip(p4, w4))
    assert vol == pytest.approx(vol4, rel=1e-12)


def _trilinear(X, xi):
    sx = np.array([-1, 1, 1, -1, -1, 1, 1, -1.0]); sy = np.array([-1, -1, 1, 1, -1, -1, 1, 1.0]); sz = np.array([-1, -1, -1, -1, 1, 1, 1, 1.0])
    N = 0.125 * (1 + sx * xi[0]) * (1 + sy * xi[1]) * (1 + sz * xi[2])
    return N @ X


def test_conductivity_to_diffusivity(oracle):
    """test/test_coefficients.jl:164-188: ConductivityToDiffusivityCoefficient(kappa, Cm=2, chi=0.5) == kappa."""
    O = oracle
    X = DISTORTED_HEX
    k = [1.0, 0.2, 0.0, 0.2, 2.0, 0.1, 0.0, 0.1, 0.5]
    K1 = O.element_diffusion(O.HEX8, 2, X, O.D_TENSOR, k, cmchi=2.0 * 0.5)
    K2 = O.element_diffusion(O.HEX8, 2, X, O.D_TENSOR, k, cmchi=1.0)
    assert np.array_equal(K1, K2)
    K3 = O.element_diffusion(O.HEX8, 2, X, O.D_TENSOR, k, cmchi=2.0)
    assert np.allclose(K3, K1 / 2.0, rtol=1e-15)


def test_ltg_order_heat_then_cells(oracle):
    """test/test_os_gearing.jl:471-508 idea: LTG advances children sequentially, each from the previous
    child's result.  With K = 0 the heat child is the identity (A = M), so one LTG step must equal one
    cell step to CG tolerance; with the cell model frozen (dt_cell effect zero at rest) it is pure BE."""
    O = oracle
    m = O.generate_grid(O.QUAD4, (3, 3), (0, 0), (1, 1))
    M = O.assemble_mass(m, 2)
    prm = O.default_params(O.FHN)
    u = np.concatenate([np.linspace(0, 1, m.ndofs), np.full(m.ndofs, 0.05)])
    u_ref = u.copy()
    orc = O.MonodomainOracle(m, O.FHN, prm, M, np.zeros_like(M), atol=1e-15, rtol=1e-15)
    orc.step(u, 0.0, 0.25)
    O.cell_step(O.FHN, prm, u_ref, m.ndofs, 0.0, 0.25)
    assert np.allclose(u, u_ref, rtol=1e-12, atol=1e-14)


def test_spectral_tensor_coefficient_known_tensors(oracle):
    """test/test_coefficients.jl:107-141 (SpectralTensorCoefficient): eigenvector e1 with eigenvalues (-1, 0) gives
    diag(-1, 0[, 0]); equal eigenvalues give -I whatever the frame.  Checked through the element matrix: the spectral
    coefficient over a constant frame must assemble exactly what the equivalent constant tensor assembles; a rotated
    frame gives R diag(lambda) R^T; and orthogonalize_system (microstructure.jl:176-187) makes a skewed input frame
    equivalent to its Gram-Schmidt image."""
    O = oracle
    from conftest import DISTORTED_HEX
    X = DISTORTED_HEX
    nv = 8

    def spectral(lam, f, s, n, cell=0):
        frame = np.tile(np.concatenate([f, s, n]), (nv, 1)).ravel()
        return O.element_diffusion(O.HEX8, 2, X, O.D_SPECTRAL, np.concatenate([lam, frame]), cell=cell)

    e1, e2, e3 = np.eye(3)
    K = spectral([-1.0, 0.0, 0.0], e1, e2, e3)
    assert np.allclose(K, O.element_diffusion(O.HEX8, 2, X, O.D_TENSOR, np.diag([-1.0, 0.0, 0.0]).ravel()), rtol=0, atol=1e-15)
    K = spectral([-1.0, -1.0, -1.0], e1, e2, e3)
    assert np.allclose(K, O.element_diffusion(O.HEX8, 2, X, O.D_TENSOR, (-np.eye(3)).ravel()), rtol=0, atol=1e-15)
    assert np.allclose(K, O.element_diffusion(O.HEX8, 2, X, O.D_SCALAR, [-1.0]), rtol=0, atol=1e-15)
    # rotated orthonormal frame
    rng = np.random.default_rng(0)
    Q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    lam = np.array([0.1334, 0.0176, 0.0051])
    D = Q @ np.diag(lam) @ Q.T
    K = spectral(lam, Q[:, 0], Q[:, 1], Q[:, 2])
    Kt = O.element_diffusion(O.HEX8, 2, X, O.D_TENSOR, D.ravel())
    assert np.allclose(K, Kt, rtol=0, atol=1e-13 * np.abs(Kt).max())
    # skewed, unnormalised input frame: normalise each, then Gram-Schmidt without renormalising (utils.jl:131-139);
    # f stays, s loses its f-component (the result is orthogonal but s, n are no longer unit vectors)
    f, s, n = 2.0 * Q[:, 0], Q[:, 1] + 0.4 * Q[:, 0], 0.5 * Q[:, 2] + 0.3 * Q[:, 1]
    fh, sh, nh = f / np.linalg.norm(f), s / np.linalg.norm(s), n / np.linalg.norm(n)
    w1 = fh
    w2 = sh - (w1 @ sh) * w1
    w3 = nh - (w1 @ nh) * w1 - (w2 @ nh) * w2
    D = lam[0] * np.outer(w1, w1) + lam[1] * np.outer(w2, w2) + lam[2] * np.outer(w3, w3)
    K = spectral(lam, f, s, n)
    Kt = O.element_diffusion(O.HEX8, 2, X, O.D_TENSOR, D.ravel())
    assert np.allclose(K, Kt, rtol=0, atol=1e-13 * np.abs(Kt).max())


def test_field_coefficient_interpolates_nodal_data(oracle):
    """test/test_coefficients.jl:39-71 (FieldCoefficient): per-element nodal data are interpolated with the shape values
    at the quadrature point.  A frame that varies linearly over the element must give, at order-1 quadrature (one point,
    the centroid), exactly the tensor of the nodal average."""
    O = oracle
    m = O.generate_grid(O.HEX8, (1, 1, 1), (0, 0, 0), (1, 1, 1))
    X = m.coords[m.conn[0]]
    rng = np.random.default_rng(1)
    f = np.tile([1.0, 0.0, 0.0], (8, 1)) + 0.2 * rng.standard_normal((8, 3))
    s = np.tile([0.0, 1.0, 0.0], (8, 1))
    n = np.tile([0.0, 0.0, 1.0], (8, 1))
    lam = np.array([0.3, 0.1, 0.05])
    data = np.concatenate([lam, np.concatenate([f, s, n], axis=1).ravel()])
    K = O.element_diffusion(O.HEX8, 1, X, O.D_SPECTRAL, data)
    fa = f.mean(axis=0)                      # shape values at the centroid are all 1/8
    fh = fa / np.linalg.norm(fa)
    w2 = s[0] - (fh @ s[0]) * fh
    w3 = n[0] - (fh @ n[0]) * fh - (w2 @ n[0]) * w2
    D = lam[0] * np.outer(fh, fh) + lam[1] * np.outer(w2, w2) + lam[2] * np.outer(w3, w3)
    Kt = O.element_diffusion(O.HEX8, 1, X, O.D_TENSOR, D.ravel())
    assert np.allclose(K, Kt, rtol=0, atol=1e-14 * np.abs(Kt).max())
