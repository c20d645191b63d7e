"""-m gpu: device-side generate_grid / DoF numbering / sparsity pattern against the oracle: BIT EXACT."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [("QUAD4", (17, 9)), ("QUAD4", (64, 64)), ("HEX8", (7, 5, 6)), ("HEX8", (33, 1, 1)), ("TRI3", (9, 14)),
         ("TET4", (5, 6, 4)), ("HEX8", (1, 1, 1)), ("QUAD4", (1, 1))]


def _boxes(dim):
    return ((0.0, -1.0, 0.5)[:dim], (2.5, 1.0, 2.0)[:dim])


@pytest.mark.parametrize("name,nel", CASES)
def test_generate_grid_numbering_and_pattern(tb, dev, oracle, name, nel):
    O = oracle
    ct = getattr(O, name)
    left, right = _boxes(len(nel))
    mo = O.generate_grid(ct, nel, left, right)
    md = tb.generate_mesh(ct, nel, left, right, device=dev)
    assert (md.ncells, md.nnodes, md.ndofs, md.nv, md.dim) == (mo.ncells, mo.nnodes, mo.ndofs, mo.nv, mo.dim)
    conn, coords, celldofs = md.download()
    assert np.array_equal(conn, mo.conn)
    assert np.array_equal(coords, mo.coords)              # same formula, IEEE division: bitwise
    assert np.array_equal(celldofs, mo.celldofs)          # first-touch DoF numbering
    assert np.array_equal(md.dof_coords(), mo.dof_coords)
    A = tb.B200CSRMatrix.from_mesh(dev, md)
    rp, ci = A.pattern()
    rpo, cio = mo.pattern()
    assert A.nnz == cio.size and np.array_equal(rp, rpo) and np.array_equal(ci, cio)
    rp1, ci1 = A.pattern(index_base=1)
    assert np.array_equal(rp1, rpo + 1) and np.array_equal(ci1, cio + 1)
    A.free(); md.free()


def test_host_mesh_and_pattern_upload_roundtrip(tb, dev, oracle):
    """The drop-in path: Ferrite's grid/celldofs/pattern come from the host, 1-based."""
    O = oracle
    mo = O.generate_grid(O.TET4, (4, 3, 5), (0, 0, 0), (1, 1, 1))
    md = tb.DeviceMesh.from_host(dev, O.TET4, mo.conn + 1, mo.coords, mo.celldofs + 1, mo.ndofs, index_base=1)
    conn, coords, celldofs = md.download()
    assert np.array_equal(conn, mo.conn) and np.array_equal(celldofs, mo.celldofs) and np.array_equal(coords, mo.coords)
    rpo, cio = mo.pattern()
    A = tb.B200CSRMatrix.from_pattern(dev, rpo + 1, cio + 1, index_base=1)
    rp, ci = A.pattern()
    assert np.array_equal(rp, rpo) and np.array_equal(ci, cio)
    B = tb.B200CSRMatrix.from_mesh(dev, md)
    assert np.array_equal(B.pattern()[1], cio)
    vals = np.random.default_rng(0).standard_normal(cio.size)
    A.set_nonzeros(vals)
    assert np.array_equal(A.nonzeros(), vals)
    A2 = A.like()
    assert not A2.nonzeros().any()
    # to_mesh: host numbering done by the API's own close_dofs
    mh = tb.to_mesh(O.TET4, mo.conn, mo.coords, device=dev)
    assert np.array_equal(mh.download()[2], mo.celldofs)
    for h in (A, A2, B, md, mh):
        h.free()


def test_bad_patterns_are_rejected(tb, dev):
    rp = np.array([0, 2, 4], dtype=np.int64)
    with pytest.raises(tb.TBError):
        tb.B200CSRMatrix.from_pattern(dev, rp, np.array([1, 0, 0, 1]))        # unsorted row
    with pytest.raises(tb.TBError):
        tb.B200CSRMatrix.from_pattern(dev, rp, np.array([0, 1, 0, 5]))        # column out of range
    with pytest.raises(tb.TBError):
        tb.generate_mesh(1, (0, 1, 1), (0, 0, 0), (1, 1, 1), device=dev)       # empty grid


def test_c2_size_pattern_closed_form(tb, dev):
    """BASELINE config 2 at full size: N = 129*129*33, nnz = 385^2*97."""
    md = tb.generate_mesh(tb.Hexahedron, (128, 128, 32), (0, 0, 0), (32, 32, 8), device=dev)
    assert md.ndofs == 549153
    A = tb.B200CSRMatrix.from_mesh(dev, md)
    assert A.nnz == 385 * 385 * 97
    rp, ci = A.pattern()
    assert np.all(np.diff(rp) >= 8) and np.all(np.diff(rp) <= 27)
    rows = np.repeat(np.arange(md.ndofs), np.diff(rp))
    assert np.all(np.diff(ci)[np.diff(rows) == 0] > 0)                 # sorted within rows
    import scipy.sparse as sp
    P = sp.csr_matrix((np.ones(ci.size), ci, rp))
    assert (P != P.T).nnz == 0 and np.all(P.diagonal() == 1)
    A.free(); md.free()


@pytest.mark.parametrize("celltype,nel", [("Hexahedron", (7, 5, 9)), ("Hexahedron", (3, 4, 2)), ("Hexahedron", (1, 1, 6)), ("Quadrilateral", (9, 11)),
                                           ("Quadrilateral", (2, 3))])
def test_local_structured_generator_equals_extract_of_the_global_grid(tb, dev, celltype, nel):
    """tb_mesh_generate_grid_local (closed-form first-touch numbering, slab-sized temporaries) against
    tb_mesh_extract_local(tb_mesh_generate_grid(...)): identical cells, nodes, coordinates, dofs and ghost lists"""
    ct = getattr(tb, celltype)
    dim = len(nel)
    left, right = (0.0,) * dim, tuple(0.25 * n + 0.1 for n in nel)
    full = tb.generate_mesh(ct, nel, left, right, device=dev)
    n = full.ndofs
    rng = np.random.default_rng(5)
    cuts = sorted(set([0, n] + [int(v) for v in rng.integers(1, n, 4)] + [n // 2, n // 3]))
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        a = full.extract_local(lo, hi)
        b = tb.DeviceMesh.generate_grid_local(dev, ct, nel, left, right, lo, hi)
        assert (a.ncells, a.nnodes, a.ndofs, a.ndofs_owned) == (b.ncells, b.nnodes, b.ndofs, b.ndofs_owned)
        assert np.array_equal(a.ghost_global, b.ghost_global)
        for x, y in zip(a.download(), b.download()):
            assert np.array_equal(x, y)
        assert np.array_equal(a.dof_coords(), b.dof_coords())
        a.free()
        b.free()
    full.free()
