"""-m gpu: the default assembly strategy (per-element results + ordered row gather, tb_assembly.cu mode 2)
is BITWISE the oracle's sequential element loop -- values, not just pattern -- for every cell type, with one
chunk or many, and falls back to the atomic scatter (values to rounding) only when the scratch cannot fit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [("QUAD4", (13, 7)), ("HEX8", (6, 5, 4)), ("TRI3", (8, 9)), ("TET4", (4, 5, 3)), ("HEX8", (40, 9, 7))]
# (mesh, fraction of all cells the scratch may hold): small enough for several row chunks, large enough for the
# cell range one 32-row slice touches (a few grid planes on these tiny meshes)
CHUNKED = [("QUAD4", (13, 7), 0.7), ("HEX8", (6, 5, 4), 0.85), ("TRI3", (8, 9), 0.85), ("TET4", (4, 5, 6), 0.85),
           ("HEX8", (40, 9, 7), 0.5)]
ALL = [(n, e, None) for n, e in CASES] + CHUNKED


def _pair(tb, dev, O, name, nel, warp=0.02, shuffle=False):
    ct = getattr(O, name)
    dim = len(nel)
    mo = O.generate_grid(ct, nel, (0.0, -1.0, 0.5)[:dim], (2.5, 1.0, 2.0)[:dim])
    rng = np.random.default_rng(11)
    mo.coords += warp * rng.standard_normal(mo.coords.shape)
    if shuffle:   # a numbering without locality: cell order permuted (dof numbering follows first touch in the new order)
        perm = rng.permutation(mo.ncells)
        mo = O.Mesh(ct, mo.conn[perm], mo.coords)
    md = tb.DeviceMesh.from_host(dev, ct, mo.conn, mo.coords, mo.celldofs, mo.ndofs)
    return mo, md


@pytest.fixture()
def gather(dev):
    dev.assembly_set_mode(2)
    dev.assembly_set_scratch_budget(8 << 30)
    yield dev
    dev.assembly_set_mode(2)
    dev.assembly_set_scratch_budget(8 << 30)


@pytest.mark.parametrize("name,nel,frac", ALL)
def test_matrices_bitwise(tb, gather, oracle, name, nel, frac):
    O, dev = oracle, gather
    mo, md = _pair(tb, dev, O, name, nel)
    if frac:
        dev.assembly_set_scratch_budget(int(frac * mo.ncells) * mo.nv * mo.nv * 8)
    M = tb.B200CSRMatrix.from_mesh(dev, md)
    K = M.like()
    tb.core.assemble_mass(dev, md, M, 2, 1.3)
    info = dev.assembly_info()
    assert info["last_mode"] == 2 and (info["last_chunks"] > 1) == bool(frac)
    assert np.array_equal(M.nonzeros(), O.assemble_mass(mo, 2, 1.3))
    tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_SCALAR, [0.37], 1.0)
    assert np.array_equal(K.nonzeros(), O.assemble_diffusion(mo, 2, O.D_SCALAR, [0.37]))
    rng = np.random.default_rng(2)
    B = rng.standard_normal((mo.dim, mo.dim))
    D = B @ B.T + np.eye(mo.dim)
    tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_TENSOR, D, 2.0)
    Ko = O.assemble_diffusion(mo, 2, O.D_TENSOR, D, cmchi=2.0)
    assert np.array_equal(K.nonzeros(), Ko)
    if mo.dim == 3:
        data = np.concatenate([[0.1334, 0.0176, 0.0176], rng.standard_normal((mo.ncells, mo.nv, 9)).ravel()])
        tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_SPECTRAL, data, 1.0)
        assert np.array_equal(K.nonzeros(), O.assemble_diffusion(mo, 2, O.D_SPECTRAL, data))
    # higher quadrature orders go through the same path
    if name in ("QUAD4", "HEX8"):
        tb.core.assemble_diffusion(dev, md, K, 3, tb._lib.D_TENSOR, D, 1.0)
        assert np.array_equal(K.nonzeros(), O.assemble_diffusion(mo, 3, O.D_TENSOR, D))
    for h in (M, K, md):
        h.free()


@pytest.mark.parametrize("name,nel,frac", ALL)
def test_source_vectors_bitwise(tb, gather, oracle, name, nel, frac):
    O, dev = oracle, gather
    mo, md = _pair(tb, dev, O, name, nel)
    if frac:
        dev.assembly_set_scratch_budget(int(frac * mo.ncells) * mo.nv * 8)
    b = tb.B200Vector(dev, mo.ndofs, 1)
    for kind, prm, t in ((O.SRC_BOX, [1.5, 2.0, 0.5], 0.01), (O.SRC_BALL, [1.6, 2.0, 0.01], 1.0), (O.SRC_NORMT, [0.0], 1.5),
                         (O.SRC_ENDO, [0.6, 2.0, 0.5, 0.25], 0.5), (O.SRC_BOX, [1.5, 2.0, 0.5], 2.5)):
        tb.core.assemble_source(dev, md, b, 2, kind, prm, t)
        info = dev.assembly_info()
        assert info["last_mode"] == 2 and (info["last_chunks"] > 1) == bool(frac)
        # exp() in SRC_ENDO is the only device/host libm difference: all other families are bitwise
        bo = O.assemble_source(mo, 2, kind, prm, t)
        if kind == O.SRC_ENDO:
            assert np.allclose(b.to_host(), bo, rtol=4e-16, atol=0)
        else:
            assert np.array_equal(b.to_host(), bo)
    nq = len(O.quadrature(getattr(O, name), 2)[1])
    fq = np.random.default_rng(6).standard_normal((mo.ncells, nq))
    tb.core.assemble_source_qp(dev, md, b, 2, fq)
    assert np.array_equal(b.to_host(), O.assemble_source(mo, 2, O.SRC_NONE, [0.0], 0.0, fq_all=fq))
    b.free(); md.free()


def test_fallback_and_explicit_atomic_mode(tb, gather, oracle):
    """A cell numbering without locality makes every row chunk span (almost) all cells: with a scratch too small for
    that the call must fall back to the atomic scatter and say so; mode 0 can also be requested."""
    O, dev = oracle, gather
    mo, md = _pair(tb, dev, O, "HEX8", (9, 8, 7), shuffle=True)
    M = tb.B200CSRMatrix.from_mesh(dev, md)
    Mo = O.assemble_mass(mo, 2, 1.0)
    tb.core.assemble_mass(dev, md, M, 2, 1.0)                 # big scratch: gather works on any numbering
    assert dev.assembly_info()["last_mode"] == 2 and np.array_equal(M.nonzeros(), Mo)
    dev.assembly_set_scratch_budget(16 << 10)                # 32 hex matrices: no chunk of 32 rows fits
    tb.core.assemble_mass(dev, md, M, 2, 1.0)
    assert dev.assembly_info()["last_mode"] == 0
    assert np.allclose(M.nonzeros(), Mo, rtol=0, atol=1e-13 * np.abs(Mo).max())
    dev.assembly_set_scratch_budget(8 << 30)
    dev.assembly_set_mode(0)
    tb.core.assemble_mass(dev, md, M, 2, 1.0)
    assert dev.assembly_info()["last_mode"] == 0
    assert np.allclose(M.nonzeros(), Mo, rtol=0, atol=1e-13 * np.abs(Mo).max())
    with pytest.raises(tb.TBError):
        dev.assembly_set_mode(1)
    M.free(); md.free()


def test_generated_grid_large_is_deterministic_and_chunked(tb, gather):
    """Size-independent properties on a mesh far larger than the oracle handles in seconds: two assemblies are
    bitwise identical, a chunked run equals the one-chunk run, sum(M) = volume, K 1 = 0."""
    dev = gather
    md = tb.generate_mesh(tb.Hexahedron, (96, 64, 48), (0, 0, 0), (24.0, 16.0, 12.0), device=dev)
    M = tb.B200CSRMatrix.from_mesh(dev, md)
    K = M.like()
    D = np.diag([0.1334, 0.0176, 0.0176])
    tb.core.assemble_mass(dev, md, M, 2, 1.0)
    tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_TENSOR, D, 1.0)
    m1, k1 = M.nonzeros(), K.nonzeros()
    dev.assembly_set_scratch_budget(24 << 20)
    tb.core.assemble_mass(dev, md, M, 2, 1.0)
    assert dev.assembly_info()["last_chunks"] > 1
    tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_TENSOR, D, 1.0)
    assert np.array_equal(M.nonzeros(), m1) and np.array_equal(K.nonzeros(), k1)
    one = tb.B200Vector.from_host(dev, np.ones(md.ndofs))
    y = tb.B200Vector(dev, md.ndofs)
    M.mul(y, one)
    assert y.to_host().sum() == pytest.approx(24.0 * 16.0 * 12.0, rel=1e-12)
    K.mul(y, one)
    assert np.abs(y.to_host()).max() < 1e-12
    for h in (M, K, one, y, md):
        h.free()
