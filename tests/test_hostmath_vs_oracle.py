"""The product's host/device inline arithmetic (tb_cells.cuh, tb_elements.cuh), compiled for the host,
against the oracle -- catches arithmetic slips before any GPU time is spent."""
import ctypes as C

import numpy as np
import pytest

from conftest import DISTORTED_HEX

CT = {"QUAD4": 0, "HEX8": 1, "TRI3": 2, "TET4": 3}


def _cell(O, ct):
    if O.cell_dim(ct) == 2:
        m = O.generate_grid(ct, (2, 2), (0.0, 0.0), (1.0, 1.5))
    else:
        m = O.generate_grid(ct, (2, 2, 2), (0.0, 0.0, 0.0), (1.0, 1.5, 0.8))
    rng = np.random.default_rng(3)
    return np.ascontiguousarray(m.coords[m.conn[1]] + 0.02 * rng.standard_normal((m.nv, m.dim)))


@pytest.mark.parametrize("name", list(CT))
@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_tables(oracle, hostmath, name, order):
    O, ct = oracle, CT[name]
    nq = C.c_int()
    xi, w, N, dN = np.zeros(64 * 3), np.zeros(64), np.zeros(64 * 8), np.zeros(64 * 24)
    rc = hostmath.hm_tables(ct, order, C.byref(nq), xi, w, N, dN)
    try:
        pts, wo = O.quadrature(ct, order)
    except ValueError:
        assert rc == 1
        return
    assert rc == 0 and nq.value == len(wo)
    nv, dim = O.cell_nv(ct), O.cell_dim(ct)
    assert np.array_equal(xi[:nq.value * dim].reshape(-1, dim), pts) and np.array_equal(w[:nq.value], wo)
    for q in range(nq.value):
        No, dNo = O.shape(ct, pts[q])
        assert np.array_equal(N[q * nv:(q + 1) * nv], No)
        assert np.array_equal(dN[q * nv * dim:(q + 1) * nv * dim].reshape(nv, dim), dNo)


@pytest.mark.parametrize("name", list(CT))
def test_element_kernels(oracle, hostmath, name):
    O, ct = oracle, CT[name]
    X = _cell(O, ct)
    nv, dim = O.cell_nv(ct), O.cell_dim(ct)
    out = np.zeros((nv, nv))
    hostmath.hm_element_matrix(ct, 2, 0, X.ravel(), 1.3, 0, np.zeros(1), 1.0, 0, out.reshape(-1))
    assert np.allclose(out, O.element_mass(ct, 2, X, 1.3), rtol=1e-15, atol=0)
    rng = np.random.default_rng(5)
    A = rng.standard_normal((dim, dim))
    D = A @ A.T + np.eye(dim)
    for kind, data in ((0, np.array([0.37])), (1, D.ravel())):
        hostmath.hm_element_matrix(ct, 2, 1, X.ravel(), 1.0, kind, np.ascontiguousarray(data), 2.0, 0, out.reshape(-1))
        ref = O.element_diffusion(ct, 2, X, kind, data, cmchi=2.0)
        assert np.allclose(out, ref, rtol=1e-13, atol=1e-16 * np.abs(ref).max())
    if dim == 3:
        lam = np.array([0.3, 0.1, 0.05])
        fsn = rng.standard_normal((3, nv, 9))       # three cells worth; use cell 2
        data = np.concatenate([lam, fsn.ravel()])
        hostmath.hm_element_matrix(ct, 2, 1, X.ravel(), 1.0, 2, data, 1.0, 2, out.reshape(-1))
        ref = O.element_diffusion(ct, 2, X, 2, data, cmchi=1.0, cell=2)
        assert np.allclose(out, ref, rtol=1e-12, atol=1e-15 * np.abs(ref).max())
    be = np.zeros(nv)
    for kind, prm in ((1, [0.9, 2.0, 0.5]), (2, [1.2, 2.0, 0.01]), (3, [0.0]), (4, [0.0]), (5, [0.6, 2.0, 0.5, 0.25])):
        p8 = np.zeros(8); p8[:len(prm)] = prm
        hostmath.hm_element_source(ct, 2, X.ravel(), kind, p8, 0.4, None, be)
        assert np.allclose(be, O.element_source(ct, 2, X, kind, p8, 0.4), rtol=1e-14, atol=1e-18)
    fq = rng.standard_normal(len(O.quadrature(ct, 2)[1]))
    hostmath.hm_element_source(ct, 2, X.ravel(), 0, np.zeros(8), 0.0, fq.ctypes.data, be)
    assert np.allclose(be, O.element_source(ct, 2, X, 0, np.zeros(8), 0.0, fq=fq), rtol=1e-14, atol=1e-18)


@pytest.mark.parametrize("name", list(CT))
def test_full_element_matrices_bitwise(oracle, hostmath, name):
    """What the deterministic gather assembly stores per element: mass (symmetric by construction) and the
    full diffusion matrix, bit for bit the oracle's."""
    O, ct = oracle, CT[name]
    X = _cell(O, ct)
    nv, dim = O.cell_nv(ct), O.cell_dim(ct)
    out = np.zeros((nv, nv))
    hostmath.hm_element_matrix(ct, 2, 0, X.ravel(), 1.3, 0, np.zeros(1), 1.0, 0, out.reshape(-1))
    assert np.array_equal(out, O.element_mass(ct, 2, X, 1.3))
    rng = np.random.default_rng(5)
    A = rng.standard_normal((dim, dim))
    D = A @ A.T + np.eye(dim)
    cases = [(0, np.array([0.37]), 0), (1, D.ravel(), 0)]
    if dim == 3:
        cases.append((2, np.concatenate([[0.3, 0.1, 0.05], rng.standard_normal((3, nv, 9)).ravel()]), 2))
    for kind, data, cell in cases:
        hostmath.hm_element_diffusion_full(ct, 2, X.ravel(), kind, np.ascontiguousarray(data), 2.0, cell, out.reshape(-1))
        assert np.array_equal(out, O.element_diffusion(ct, 2, X, kind, data, cmchi=2.0, cell=cell))


def test_distorted_hex_fixture(oracle, hostmath):
    """test/test_coefficients.jl:239-279 coordinates through the product's mapping."""
    O = oracle
    out = np.zeros((8, 8))
    hostmath.hm_element_matrix(1, 2, 0, DISTORTED_HEX.ravel(), 1.0, 0, np.zeros(1), 1.0, 0, out.reshape(-1))
    assert np.allclose(out, O.element_mass(O.HEX8, 2, DISTORTED_HEX), rtol=1e-15)
    hostmath.hm_element_matrix(1, 2, 1, DISTORTED_HEX.ravel(), 1.0, 0, np.ones(1), 1.0, 0, out.reshape(-1))
    assert np.allclose(out, O.element_diffusion(O.HEX8, 2, DISTORTED_HEX, 0, [1.0]), rtol=1e-13, atol=1e-16)
    assert np.allclose(out.sum(axis=1), 0.0, atol=1e-14)


def test_fhn_node_step_bitwise(oracle, hostmath):
    O = oracle
    prm = O.default_params(O.FHN)
    rng = np.random.default_rng(7)
    for _ in range(200):
        u = np.array([rng.uniform(-0.3, 1.2), rng.uniform(-0.1, 0.3)])
        for sub in (1, 10):
            a, b = u.copy(), u.copy()
            O.cell_step(O.FHN, prm, a, 1, 0.3, 0.7, substeps=sub, threshold=0.1)
            hostmath.hm_cell_node_step(0, int(sub > 1), prm, b, 0.3, 0.7, sub, 0.1)
            assert np.array_equal(a, b)


def test_aliev_panfilov_node_step_bitwise(oracle, hostmath):
    """aliev-panfilov.jl:15-34, states (s, phi): rational in the state, so bitwise like FHN."""
    O = oracle
    prm = O.default_params(O.ALIEV_PANFILOV)
    assert np.array_equal(prm, [1.0 / 12.9, 8.0, 0.05, 0.002, 0.2, 0.3])          # aliev-panfilov.jl:2-7
    rng = np.random.default_rng(8)
    for _ in range(200):
        u = np.array([rng.uniform(0.0, 2.5), rng.uniform(-0.1, 1.1)])
        for sub in (1, 10):
            a, b = u.copy(), u.copy()
            du = O.cell_step(O.ALIEV_PANFILOV, prm, a, 1, 0.3, 0.7, substeps=sub, threshold=0.1, phi_idx=1)
            d = hostmath.hm_cell_node_step(2, int(sub > 1), prm, b, 0.3, 0.7, sub, 0.1)
            assert np.array_equal(a, b) and d == du[1]
    # closed form of one rhs evaluation
    s_, phi = 0.4, 0.6
    du = O.cell_rhs(O.ALIEV_PANFILOV, prm, [s_, phi])
    ct, k, a_, e0, m1, m2 = prm
    assert du[1] == ct * (k * phi * (phi - 1.0) * (phi - a_) - phi * s_)
    assert du[0] == ct * (e0 + s_ * m1 / (phi + m2)) * (-s_ - k * phi * (phi - a_ - 1.0))


def test_pcg2019_node_step(oracle, hostmath):
    O = oracle
    prm = O.default_params(O.PCG2019)
    rng = np.random.default_rng(9)
    u0 = O.default_initial_state(O.PCG2019)
    for _ in range(200):
        u = u0.copy()
        u[0] = rng.uniform(-90.0, 40.0)
        u[1:] = np.clip(u[1:] + rng.uniform(-0.3, 0.3, 6), 0.0, 1.0)
        for sub in (1, 10):
            a, b = u.copy(), u.copy()
            du = O.cell_step(O.PCG2019, prm, a, 1, 0.0, 0.01, substeps=sub, threshold=0.1)
            d = hostmath.hm_cell_node_step(1, int(sub > 1), prm, b, 0.0, 0.01, sub, 0.1)
            assert np.array_equal(a, b)          # same libm exp on the host: bitwise
            assert d == du[0]


@pytest.mark.parametrize("nel", [(4, 3, 5), (1, 1, 1), (2, 5, 1), (7, 1, 3), (5, 4), (1, 3), (6, 1)])
def test_closed_form_first_touch_numbering(hostmath, oracle, nel):
    """tb_grid_dof (the numbering tb_mesh_generate_grid_local builds local meshes from, no global grid in HBM) against the
    oracle's close! restatement on the whole grid"""
    O = oracle
    dim = len(nel)
    m = O.generate_grid(O.HEX8 if dim == 3 else O.QUAD4, nel, (0,) * dim, (1,) * dim)
    got = np.empty(m.nnodes, dtype=np.int64)
    hostmath.hm_grid_dofs(dim, np.array(list(nel) + [1] * (3 - dim), dtype=np.int64), got)
    assert np.array_equal(got, m.node2dof)
