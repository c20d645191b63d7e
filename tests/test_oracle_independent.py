"""Independent (numpy/scipy) cross-checks of the C oracle: closed forms and a second implementation
of pattern, element matrices, SpMV and CG written without looking at tb_oracle.c's structure."""
import numpy as np
import pytest
import scipy.sparse as sp

ALL = ["QUAD4", "HEX8", "TRI3", "TET4"]


def _grid(O, name, n=3):
    ct = getattr(O, name)
    if O.cell_dim(ct) == 2:
        return O.generate_grid(ct, (n, n + 1), (0.0, -1.0), (2.0, 1.0))
    return O.generate_grid(ct, (n, n + 1, 2), (0.0, -1.0, 0.5), (2.0, 1.0, 1.5))


@pytest.mark.parametrize("name", ALL)
def test_counts_and_first_touch_numbering(oracle, name):
    O = oracle
    m = _grid(O, name)
    nx = 4
    assert m.ndofs == m.nnodes
    # first cell gets dofs 0..nv-1 in local vertex order; every dof id appears; ids are first-touch ordered
    assert list(m.celldofs[0]) == list(range(m.nv))
    flat = m.celldofs.ravel()
    first = {}
    for p, d in enumerate(flat):
        first.setdefault(int(d), p)
    order = sorted(first, key=first.get)
    assert order == list(range(m.ndofs))
    # node numbering is x-fastest
    assert np.allclose(m.coords[1] - m.coords[0], [2.0 / 3] + [0.0] * (m.dim - 1))
    assert m.coords[nx][0] == 0.0 and m.coords[nx][1] > m.coords[0][1]


def test_closed_form_sizes(oracle):
    O = oracle
    m = O.generate_grid(O.QUAD4, (16, 8), (0, 0), (1, 1))
    assert (m.ndofs, m.pattern()[1].size) == (17 * 9, (3 * 17 - 2) * (3 * 9 - 2))
    m = O.generate_grid(O.HEX8, (4, 5, 3), (0, 0, 0), (1, 1, 1))
    assert (m.ndofs, m.pattern()[1].size) == (5 * 6 * 4, (3 * 5 - 2) * (3 * 6 - 2) * (3 * 4 - 2))
    m = O.generate_grid(O.TET4, (2, 2, 2), (0, 0, 0), (1, 1, 1))
    assert m.ncells == 6 * 8
    m = O.generate_grid(O.TRI3, (3, 2), (0, 0), (1, 1))
    assert m.ncells == 12


@pytest.mark.parametrize("name", ALL)
def test_pattern_matches_scipy(oracle, name):
    O = oracle
    m = _grid(O, name)
    rp, ci = m.pattern()
    i = np.repeat(m.celldofs, m.nv, axis=1).ravel()
    j = np.tile(m.celldofs, (1, m.nv)).ravel()
    P = sp.coo_matrix((np.ones(i.size), (i, j)), shape=(m.ndofs, m.ndofs)).tocsr()
    P.sum_duplicates()
    P.sort_indices()
    assert np.array_equal(P.indptr, rp) and np.array_equal(P.indices, ci)
    # sorted, diagonal present, structurally symmetric
    for r in range(m.ndofs):
        cols = ci[rp[r]:rp[r + 1]]
        assert np.all(np.diff(cols) > 0) and r in cols
    assert (P != P.T).nnz == 0


def _np_element(O, ct, qorder, X, D=None):
    """numpy element matrices: Me = sum_q N N^T detJ w ; Ke = - sum_q G D G^T detJ w"""
    pts, w = O.quadrature(ct, qorder)
    nv = O.cell_nv(ct)
    Me, Ke = np.zeros((nv, nv)), np.zeros((nv, nv))
    for p, ww in zip(pts, w):
        N, dN = O.shape(ct, p)
        J = X.T @ dN
        G = dN @ np.linalg.inv(J)
        dO = np.linalg.det(J) * ww
        Me += np.outer(N, N) * dO
        if D is not None:
            Ke -= G @ D @ G.T * dO
    return Me, Ke


@pytest.mark.parametrize("name", ALL)
def test_element_matrices_match_numpy(oracle, name):
    O = oracle
    ct = getattr(O, name)
    m = _grid(O, name)
    rng = np.random.default_rng(0)
    X = m.coords[m.conn[1]] + 0.03 * rng.standard_normal((m.nv, m.dim))
    A = rng.standard_normal((m.dim, m.dim))
    D = A @ A.T + np.eye(m.dim)
    Me, Ke = _np_element(O, ct, 2, X, D)
    assert np.allclose(O.element_mass(ct, 2, X, 1.0), Me, rtol=1e-13, atol=1e-16)
    assert np.allclose(O.element_diffusion(ct, 2, X, O.D_TENSOR, D), Ke, rtol=1e-12, atol=1e-15)
    Ks = O.element_diffusion(ct, 2, X, O.D_SCALAR, [0.7])
    assert np.allclose(Ks, _np_element(O, ct, 2, X, 0.7 * np.eye(m.dim))[1], rtol=1e-12, atol=1e-15)
    assert np.allclose(Ke.sum(axis=1), 0.0, atol=1e-13)       # constants are in the kernel of K


def test_hex_mass_closed_form(oracle):
    O = oracle
    h = 0.25
    X = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float) * h
    Me = O.element_mass(O.HEX8, 2, X)
    # h^3/216 * {8 self, 4 edge, 2 face diagonal, 1 body diagonal}
    dist = np.abs(X[:, None, :] - X[None, :, :]).sum(-1) / h
    expect = h ** 3 / 216.0 * np.choose(dist.round().astype(int), [8, 4, 2, 1])
    assert np.allclose(Me, expect, rtol=1e-14)


def test_spectral_tensor_coefficient(oracle):
    """Sum lambda_i v_i (x) v_i with normalise + Gram-Schmidt (microstructure.jl:136-187, utils.jl:131-139)."""
    O = oracle
    m = O.generate_grid(O.TET4, (2, 2, 2), (0, 0, 0), (1, 1, 1))
    lam = np.array([0.3, 0.12, 0.05])
    # constant orthonormal frame (ep04_geselowitz-ecg.jl:36-41): f = z, s = y, n = x
    fsn = np.tile(np.array([0, 0, 1, 0, 1, 0, 1, 0, 0.0]), (m.ncells, m.nv, 1))
    data = np.concatenate([lam, fsn.ravel()])
    K1 = O.assemble_diffusion(m, 2, O.D_SPECTRAL, data)
    K2 = O.assemble_diffusion(m, 2, O.D_TENSOR, np.diag([0.05, 0.12, 0.3]))
    assert np.allclose(K1, K2, rtol=1e-13, atol=1e-16)
    # non-orthogonal, non-normalised input frame gets orthogonalised the reference's way
    X = m.coords[m.conn[3]]
    f, s, n = np.array([1.0, 0.2, 0.0]), np.array([0.3, 2.0, 0.1]), np.array([0.1, 0.1, 0.7])
    d = np.concatenate([lam, np.tile(np.concatenate([f, s, n]), 4)])
    v1, v2, v3 = f / np.linalg.norm(f), s / np.linalg.norm(s), n / np.linalg.norm(n)
    w2 = v2 - (v1 @ v2) * v1
    w3 = v3 - (v1 @ v3) * v1 - (w2 @ v3) * w2
    D = lam[0] * np.outer(v1, v1) + lam[1] * np.outer(w2, w2) + lam[2] * np.outer(w3, w3)
    assert np.allclose(O.element_diffusion(O.TET4, 2, X, O.D_SPECTRAL, d, cell=0),
                       O.element_diffusion(O.TET4, 2, X, O.D_TENSOR, D), rtol=1e-13, atol=1e-16)


@pytest.mark.parametrize("name", ALL)
def test_assembly_spmv_cg_against_scipy(oracle, name):
    O = oracle
    m = _grid(O, name, 4)
    rp, ci = m.pattern()
    M = O.assemble_mass(m, 2)
    K = O.assemble_diffusion(m, 2, O.D_SCALAR, [0.05])
    vol = 2.0 * 2.0 * (1.0 if m.dim == 3 else 1.0)
    assert M.sum() == pytest.approx(vol, rel=1e-13)
    Ms, Ks = sp.csr_matrix((M, ci, rp)), sp.csr_matrix((K, ci, rp))
    assert abs(Ms - Ms.T).max() < 1e-15 and abs(Ks - Ks.T).max() < 1e-15
    assert np.abs(Ks @ np.ones(m.ndofs)).max() < 1e-13
    # K negative semi-definite, A = M - dt K SPD
    A = O.axpby_values(M, K, 0.7)
    As = sp.csr_matrix((A, ci, rp))
    assert np.allclose(As.toarray(), (Ms - 0.7 * Ks).toarray(), rtol=0, atol=0)
    ev = np.linalg.eigvalsh(Ks.toarray())
    assert ev.max() < 1e-12
    rng = np.random.default_rng(1)
    x = rng.standard_normal(m.ndofs)
    assert np.allclose(O.spmv(rp, ci, A, x), As @ x, rtol=1e-13, atol=1e-15)
    b = As @ x
    xs, it, rn, conv = O.cg(rp, ci, A, b, atol=1e-14, rtol=1e-14)
    assert conv and np.allclose(xs, x, rtol=1e-9, atol=1e-11)
    assert rn <= 1e-14 + 1e-14 * np.linalg.norm(b)


def test_cg_recurrence_matches_python_restatement(oracle):
    """Krylov.jl cg! restated a second time in numpy; iterates must agree to rounding, counts exactly."""
    O = oracle
    m = O.generate_grid(O.QUAD4, (12, 12), (0, 0), (1, 1))
    rp, ci = m.pattern()
    A = O.axpby_values(O.assemble_mass(m), O.assemble_diffusion(m, 2, O.D_SCALAR, [1.0]), 0.01)
    As = sp.csr_matrix((A, ci, rp))
    b = np.sin(np.arange(m.ndofs) * 0.37)
    x = np.zeros_like(b); r = b.copy(); p = r.copy(); gamma = r @ r
    eps = np.sqrt(np.finfo(float).eps); tol = eps + eps * np.sqrt(gamma)
    it = 0
    while np.sqrt(gamma) > tol and it < m.ndofs:
        Ap = As @ p
        alpha = gamma / (p @ Ap)
        x += alpha * p; r -= alpha * Ap
        gn = r @ r
        it += 1
        if np.sqrt(gn) <= tol:
            gamma = gn
            break
        p = r + (gn / gamma) * p
        gamma = gn
    xo, ito, rn, conv = O.cg(rp, ci, A, b)
    assert conv and ito == it
    assert np.allclose(xo, x, rtol=1e-12, atol=1e-14)
    # itmax reached -> converged False, iterations == itmax (retcode MaxIters path, euler.jl:95-100)
    xo, ito, rn, conv = O.cg(rp, ci, A, b, itmax=3)
    assert (not conv) and ito == 3


def test_source_vector(oracle):
    O = oracle
    m = O.generate_grid(O.HEX8, (4, 4, 4), (0, 0, 0), (2, 2, 2))
    b = O.assemble_source(m, 2, O.SRC_NORMT, [0.0], 1.5)
    # sum_j b_j = int f dx by the same quadrature
    pts, w = O.quadrature(O.HEX8, 2)
    tot = 0.0
    for c in range(m.ncells):
        X = m.coords[m.conn[c]]
        for p, ww in zip(pts, w):
            det, N, _ = O.map_qp(O.HEX8, X, p)
            tot += (np.linalg.norm(N @ X) + 1.5) * det * ww
    assert b.sum() == pytest.approx(tot, rel=1e-13)
    # box stimulus of the Niederer-style benchmark: only cells near the corner contribute, off after tmax
    b1 = O.assemble_source(m, 2, O.SRC_BOX, [1.5, 2.0, 0.5], 0.01)
    assert b1.max() > 0 and np.count_nonzero(b1) < m.ndofs
    assert not O.assemble_source(m, 2, O.SRC_BOX, [1.5, 2.0, 0.5], 2.05).any()
    # host-evaluated per-qp values reproduce the built-in family
    fq = np.empty((m.ncells, 8))
    for c in range(m.ncells):
        X = m.coords[m.conn[c]]
        for q, p in enumerate(pts):
            N, _ = O.shape(O.HEX8, p)
            fq[c, q] = O.source_eval(O.SRC_COSEXP, [0.0], N @ X, 0.3)
    assert np.allclose(O.assemble_source(m, 2, O.SRC_COSEXP, [0.0], 0.3),
                       O.assemble_source(m, 2, O.SRC_NONE, [0.0], 0.3, fq_all=fq), rtol=1e-13, atol=1e-18)
