"""Poisson and Geselowitz ECG reconstructions (src/modeling/electrophysiology/ecg.jl:166-619) on the reference's
own test geometry and with its own assertions (test/integration/test_ecg.jl:6-230: equilibrium, idempotence, the x^3 planar
wave whose lead difference is -2*0.37 +- 1e-2, zero on the orthogonal leads), plus a host (scipy) restatement built from the
oracle's operators."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

SIZE = 2.0


def _geometry(O):
    heart = O.generate_grid(O.HEX8, (6, 6, 6), (-1, -1, -1), (1, 1, 1))
    hx = np.sign(heart.coords) * heart.coords ** 2                       # transform_coordinates!(x -> sign(x) x^2)
    torso = O.generate_grid(O.HEX8, (16, 16, 16), (-SIZE,) * 3, (SIZE,) * 3)
    heart_cells = np.flatnonzero((np.abs(torso.coords[torso.conn]).max(axis=2) <= 1.0 + 1e-12).all(axis=1))
    kap_i = np.zeros((torso.ncells, 9))
    kap_i[heart_cells] = np.eye(3).ravel()
    electrodes = np.array([[0, 0, 0], [-SIZE, 0, 0], [SIZE, 0, 0], [0, -SIZE, 0], [0, SIZE, 0], [0, 0, -SIZE], [0, 0, SIZE]], dtype=np.float64)
    return heart, hx, torso, heart_cells, kap_i, electrodes


def test_point_location_and_interpolation_rows(tb, oracle):
    """interpolation_rows reproduces linear functions exactly, on hexahedra (Newton on the trilinear map) and tetrahedra"""
    O = oracle
    rng = np.random.default_rng(2)
    for ct, tbct in ((O.HEX8, tb.Hexahedron), (O.TET4, tb.Tetrahedron)):
        m = O.generate_grid(ct, (3, 4, 2), (-1, 0, 0.5), (1, 2, 1.5))
        X = m.coords + (0.03 * rng.standard_normal(m.coords.shape) if ct == O.TET4 else 0.0)
        f = lambda x: 2.0 * x[..., 0] - 0.5 * x[..., 1] + 3.0 * x[..., 2] + 1.0
        nodal = np.empty(m.ndofs)
        nodal[m.celldofs.ravel()] = f(X[m.conn.ravel()])
        pts = np.array([-1, 0, 0.5]) + rng.random((40, 3)) * [1.9, 1.9, 0.9] + 0.04
        rows = tb.ecg.interpolation_rows(tbct, m.conn, X, m.celldofs, pts)
        assert all(r is not None for r in rows)
        vals = np.array([w @ nodal[d] for d, w in rows])
        assert np.abs(vals - f(pts)).max() < 1e-12
        assert tb.ecg.interpolation_rows(tbct, m.conn, X, m.celldofs, [[5.0, 5.0, 5.0]])[0] is None


def _host_reference(O, heart, hx, torso, heart_cells, kap_i, electrodes, phi_m, ground_dof):
    """scipy restatement: K_bulk (ground row/column replaced), K_source, tensor-product transfer, sparse direct solve"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    rp, ci = torso.pattern()
    n = torso.ndofs
    Kb = sp.csr_matrix((O.assemble_diffusion(torso, 2, O.D_TENSOR, np.eye(3)), ci, rp), shape=(n, n)).tolil()
    Ks = sp.csr_matrix((O.assemble_diffusion(torso, 2, 3, kap_i.ravel()), ci, rp), shape=(n, n))
    md = np.abs(Kb.diagonal()).mean()
    Kb[ground_dof, :] = 0.0
    Kb[:, ground_dof] = 0.0
    Kb[ground_dof, ground_dof] = md
    Kb = Kb.tocsc()
    # transfer: separable grids -> 1D piecewise-linear interpolation per axis (independent of the product's point location)
    g1 = np.sign(np.linspace(-1, 1, 7)) * np.linspace(-1, 1, 7) ** 2
    hnode = {tuple(np.round(x, 12)): d for x, d in zip(hx[heart.conn.ravel()], heart.celldofs.ravel())}
    xd = torso.dof_coords
    phi_t = np.zeros(n)
    for d in np.unique(torso.celldofs[heart_cells]):
        w = []
        for a in range(3):
            x = xd[d, a]
            k = min(max(np.searchsorted(g1, x, side="right") - 1, 0), 5)
            t = (x - g1[k]) / (g1[k + 1] - g1[k])
            w.append(((k, 1 - t), (k + 1, t)))
        v = 0.0
        for (i, wi) in w[0]:
            for (j, wj) in w[1]:
                for (k, wk) in w[2]:
                    if wi * wj * wk != 0.0:
                        v += wi * wj * wk * phi_m[hnode[(round(g1[i], 12), round(g1[j], 12), round(g1[k], 12))]]
        phi_t[d] = v
    src = Ks @ phi_t
    rhs = -src
    rhs[ground_dof] = 0.0
    phi_e = spl.spsolve(Kb, rhs)
    rows = [None] * len(electrodes)
    return Kb, src, phi_e, phi_t


@pytest.mark.gpu
def test_poisson_and_geselowitz_on_the_reference_blocks(tb, dev, oracle):
    O = oracle
    heart, hx, torso, heart_cells, kap_i, electrodes = _geometry(O)
    heart_dh = tb.DeviceMesh.from_host(dev, tb.Hexahedron, heart.conn, hx, heart.celldofs, heart.ndofs)
    ground_node = tb.ecg.get_closest_vertex([0.0, 0.0, 0.0], torso.coords)
    kappa = tb.ConstantCoefficient(np.eye(3))
    tight = tb.B200CG(atol=1e-12, rtol=1e-12)
    poisson = tb.ecg.PoissonECGReconstructionCache(tb.api, heart_dh, tb.Hexahedron, torso.conn, torso.coords, kap_i, kappa, electrodes,
                                                   ground_node, heart_cells, linear_solver=tight)
    el_nodes = [tb.ecg.get_closest_vertex(e, torso.coords) for e in electrodes]
    leads = [[el_nodes[0], el_nodes[i]] for i in range(1, 7)]
    gesel = tb.ecg.Geselowitz1989ECGLeadCache(tb.api, heart_dh, tb.Hexahedron, torso.conn, torso.coords, kap_i, kappa, leads, ground_node,
                                              heart_cells, linear_solver=tight, qorder=3)
    xh = np.empty((heart.ndofs, 3))
    xh[heart.celldofs.ravel()] = hx[heart.conn.ravel()]
    # Equilibrium
    u = np.zeros(heart.ndofs)
    for c in (poisson, gesel):
        tb.ecg.update_ecg_(c, u)
        assert np.abs(tb.ecg.evaluate_ecg(c)).max() <= 1e-14
    assert tb.ecg.evaluate_ecg(poisson).size == 7 and tb.ecg.evaluate_ecg(gesel).size == 6
    # Idempotence
    u = np.random.default_rng(0).standard_normal(heart.ndofs)
    for c in (poisson, gesel):
        tb.ecg.update_ecg_(c, u)
        v1 = tb.ecg.evaluate_ecg(c)
        tb.ecg.update_ecg_(c, u)
        assert np.array_equal(tb.ecg.evaluate_ecg(c), v1)
    # Planar wave x^3 and -x^3 (test_ecg.jl:137-230)
    for sgn in (1.0, -1.0):
        u = sgn * xh[:, 0] ** 3
        tb.ecg.update_ecg_(poisson, u)
        pv = tb.ecg.evaluate_ecg(poisson)
        assert abs(pv[0]) <= 1e-12                                          # ground
        assert abs((pv[2] - pv[1]) - sgn * (-2 * 0.37)) <= 1e-2
        assert abs(pv[4] - pv[3]) <= 1e-4 and abs(pv[6] - pv[5]) <= 1e-4
        tb.ecg.update_ecg_(gesel, u)
        gv = tb.ecg.evaluate_ecg(gesel)
        assert abs((gv[1] - gv[0]) - sgn * (-2 * 0.37)) <= 1e-2
        assert abs(gv[3] - gv[2]) <= 1e-4 and abs(gv[5] - gv[4]) <= 1e-4
    # against the host restatement (oracle operators + scipy direct solve)
    u = np.random.default_rng(1).standard_normal(heart.ndofs) + xh[:, 0] ** 3
    gdof = int(poisson.s.ground_dofs[0])
    assert np.array_equal(poisson.s.celldofs, torso.celldofs)
    Kb, src, phi_e, phi_t = _host_reference(O, heart, hx, torso, heart_cells, kap_i, electrodes, u, gdof)
    tb.ecg.update_ecg_(poisson, u)
    assert np.abs(poisson.s.phi_t.to_host() - phi_t).max() <= 1e-12 * np.abs(phi_t).max()
    assert np.abs(poisson.s.src.to_host() - src).max() <= 1e-12 * np.abs(src).max()
    pe = poisson.phi_e.to_host()
    assert np.abs(pe - phi_e).max() <= 1e-8 * np.abs(phi_e).max()
    node2dof = np.empty(torso.nnodes, dtype=np.int64)
    node2dof[torso.conn.ravel()] = torso.celldofs.ravel()
    assert np.abs(tb.ecg.evaluate_ecg(poisson) - phi_e[node2dof[el_nodes]]).max() <= 1e-8 * np.abs(phi_e).max()
    # Geselowitz: V_i = -Z_i . src with K_bulk Z_i = f_i
    import scipy.sparse.linalg as spl
    tb.ecg.update_ecg_(gesel, u)
    gv = tb.ecg.evaluate_ecg(gesel)
    for i, es in enumerate(leads):
        f = np.zeros(torso.ndofs)
        f[node2dof[es[0]]] = -1.0
        f[node2dof[es[1]]] = 1.0
        f[gdof] = 0.0
        Zi = spl.spsolve(Kb, f)
        # qorder 3 on the source operator of the lead cache vs 2 in the host reference: same integrals (trilinear on boxes)
        assert abs(gv[i] + Zi @ src) <= 1e-7 * max(1.0, abs(Zi @ src))
