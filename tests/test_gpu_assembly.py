"""-m gpu: element-loop assembly of M, K and the stimulus vector through the C ABI vs the oracle.
Pattern is bit exact (test_gpu_mesh_pattern); values agree to rounding: the atomic scatter order is
not the oracle's element order, and the upper triangle is mirrored, so the tolerance is 1e-13 relative
to the largest entry."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [("QUAD4", (13, 7)), ("HEX8", (6, 5, 4)), ("TRI3", (8, 9)), ("TET4", (4, 5, 3))]
RTOL = 1e-13


def _pair(tb, dev, O, name, nel, warp=0.0):
    ct = getattr(O, name)
    dim = len(nel)
    mo = O.generate_grid(ct, nel, (0.0, -1.0, 0.5)[:dim], (2.5, 1.0, 2.0)[:dim])
    if warp:
        rng = np.random.default_rng(11)
        mo.coords += warp * rng.standard_normal(mo.coords.shape)
    md = tb.DeviceMesh.from_host(dev, ct, mo.conn, mo.coords, mo.celldofs, mo.ndofs)
    return mo, md


def _close(a, b):
    return np.allclose(a, b, rtol=0, atol=RTOL * np.abs(b).max())


@pytest.mark.parametrize("name,nel", CASES)
@pytest.mark.parametrize("warp", [0.0, 0.02])
def test_mass_and_diffusion(tb, dev, oracle, name, nel, warp):
    O = oracle
    mo, md = _pair(tb, dev, O, name, nel, warp)
    M = tb.B200CSRMatrix.from_mesh(dev, md)
    K = M.like()
    tb.core.assemble_mass(dev, md, M, 2, 1.0)
    assert _close(M.nonzeros(), O.assemble_mass(mo, 2, 1.0))
    tb.core.assemble_mass(dev, md, M, 2, 2.5)                     # re-assembly zeroes first
    assert _close(M.nonzeros(), O.assemble_mass(mo, 2, 2.5))
    tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_SCALAR, [0.37], 1.0)
    assert _close(K.nonzeros(), O.assemble_diffusion(mo, 2, O.D_SCALAR, [0.37]))
    rng = np.random.default_rng(2)
    B = rng.standard_normal((mo.dim, mo.dim))
    D = B @ B.T + np.eye(mo.dim)
    tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_TENSOR, D, 0.5 * 2.0)
    Ko = O.assemble_diffusion(mo, 2, O.D_TENSOR, D, cmchi=1.0)
    Kd = K.nonzeros()
    assert _close(Kd, Ko)
    # structural facts at any size: symmetric, rows sum to zero, sum(M) = volume
    import scipy.sparse as sp
    rp, ci = K.pattern()
    Ks = sp.csr_matrix((Kd, ci, rp))
    assert abs(Ks - Ks.T).max() <= 1e-14 * abs(Kd).max()
    assert np.abs(Ks @ np.ones(mo.ndofs)).max() <= 1e-12 * abs(Kd).max()
    for h in (M, K, md):
        h.free()


def test_quadrature_orders(tb, dev, oracle):
    """test/test_elements.jl:29-42 assembles with QuadratureRuleCollection(3)."""
    O = oracle
    mo, md = _pair(tb, dev, O, "HEX8", (3, 3, 3), 0.03)
    M = tb.B200CSRMatrix.from_mesh(dev, md)
    for q in (1, 2, 3, 4):
        tb.core.assemble_mass(dev, md, M, q, 1.0)
        assert _close(M.nonzeros(), O.assemble_mass(mo, q, 1.0))
        tb.core.assemble_diffusion(dev, md, M, q, 0, [1.0], 1.0)
        assert _close(M.nonzeros(), O.assemble_diffusion(mo, q, 0, [1.0]))
    with pytest.raises(tb.TBError):
        tb.core.assemble_mass(dev, md, M, 5, 1.0)
    pts, w = tb.core.quadrature(tb.Hexahedron, 2)
    po, wo = O.quadrature(O.HEX8, 2)
    assert np.array_equal(pts, po) and np.array_equal(w, wo)
    M.free(); md.free()


@pytest.mark.parametrize("name,nel", [("HEX8", (5, 4, 3)), ("TET4", (4, 4, 4))])
def test_spectral_fibre_tensor(tb, dev, oracle, name, nel):
    """config 4: SpectralTensorCoefficient over per-element nodal f,s,n (microstructure.jl:280-333)."""
    O = oracle
    mo, md = _pair(tb, dev, O, name, nel, 0.01)
    rng = np.random.default_rng(4)
    lam = np.array([0.1334, 0.0176, 0.0176])
    # a rotating, deliberately non-orthonormal frame per element node
    ang = rng.uniform(-np.pi / 3, np.pi / 3, (mo.ncells, mo.nv))
    f = np.stack([np.cos(ang), np.sin(ang), 0.1 * np.ones_like(ang)], -1)
    s = np.stack([-np.sin(ang), np.cos(ang), 0.05 * np.ones_like(ang)], -1) * 1.3
    n = np.tile([0.02, 0.01, 0.9], (mo.ncells, mo.nv, 1))
    data = np.concatenate([lam, np.stack([f, s, n], axis=2).ravel()])
    K = tb.B200CSRMatrix.from_mesh(dev, md)
    tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_SPECTRAL, data, 1.0)
    assert _close(K.nonzeros(), O.assemble_diffusion(mo, 2, O.D_SPECTRAL, data))
    with pytest.raises(tb.TBError):
        tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_SPECTRAL, data[:-1], 1.0)
    K.free(); md.free()


@pytest.mark.parametrize("name,nel", CASES)
def test_source_vector(tb, dev, oracle, name, nel):
    O = oracle
    mo, md = _pair(tb, dev, O, name, nel, 0.01)
    b = tb.B200Vector(dev, mo.ndofs, 1)
    for kind, prm, t in ((O.SRC_BOX, [1.5, 2.0, 0.5], 0.01), (O.SRC_BALL, [1.6, 2.0, 0.01], 1.0), (O.SRC_COSEXP, [0.0], 0.3),
                         (O.SRC_NORMT, [0.0], 1.5), (O.SRC_ENDO, [0.6, 2.0, 0.5, 0.25], 0.5), (O.SRC_BOX, [1.5, 2.0, 0.5], 2.5)):
        tb.core.assemble_source(dev, md, b, 2, kind, prm, t)
        bo = O.assemble_source(mo, 2, kind, prm, t)
        assert np.allclose(b.to_host(), bo, rtol=0, atol=1e-13 * max(np.abs(bo).max(), 1e-300))
    nq = len(O.quadrature(getattr(O, name), 2)[1])
    fq = np.random.default_rng(6).standard_normal((mo.ncells, nq))
    tb.core.assemble_source_qp(dev, md, b, 2, fq)
    bo = O.assemble_source(mo, 2, O.SRC_NONE, [0.0], 0.0, fq_all=fq)
    assert np.allclose(b.to_host(), bo, rtol=0, atol=1e-13 * np.abs(bo).max())
    b.free(); md.free()


def test_c3_shape_properties(tb, dev):
    """BASELINE config 3 shape at reduced size (the full 500x100x100 sweep is bench.py --workload c3):
    size-independent properties of the assembled operators on hex and tet meshes."""
    for ct, vol in ((tb.Hexahedron, 8.0), (tb.Tetrahedron, 8.0)):
        md = tb.generate_mesh(ct, (100, 20, 20), (-1, -1, -1), (1, 1, 1), device=dev)
        M = tb.B200CSRMatrix.from_mesh(dev, md)
        K = M.like()
        tb.core.assemble_mass(dev, md, M, 2, 1.0)
        tb.core.assemble_diffusion(dev, md, K, 2, 1, np.diag([0.1334, 0.0176, 0.0176]), 1.0)
        one = tb.B200Vector.from_host(dev, np.ones(md.ndofs))
        y = tb.B200Vector(dev, md.ndofs)
        M.mul(y, one)
        assert y.to_host().sum() == pytest.approx(vol, rel=1e-12)
        K.mul(y, one)
        assert np.abs(y.to_host()).max() < 1e-12
        b = tb.B200Vector(dev, md.ndofs)
        tb.core.assemble_source(dev, md, b, 2, tb._lib.SRC_NORMT, [0.0], 0.0)
        assert b.to_host().sum() > 0
        for h in (M, K, one, y, b, md):
            h.free()
