"""-m gpu: parity on the BENCHMARKED configuration's shape (BASELINE config 5, `bench.py` default workload).

C5 itself (512x512x384, 2.7 G nonzeros) does not fit the host as an assembled oracle, so
  * a C5-SHAPED slab the oracle can afford -- the full 512x512 cross-section, 16 layers, 4.47 M dofs, the bench's physics,
    initial condition and tolerances, the multi-kernel CG path large operators take and the compressed column stream on --
    is compared against the ASSEMBLED oracle: pattern / numbering bit exact, 1e-10 after one step, iterations +-1;
  * the same slab is compared against the oracle's closed-form (matrix-free) operator, which is what generates the
    full-size golden checksums bench.py's `parity` block checks at every GPU count (tests/golden/c5_checksum.json).
"""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu

KAPPA = (0.0295, 0.0131, 0.0131)     # bench.py WORKLOADS["c5"]
H, DT = 0.25, 1.0


def rel_linf(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def _bench_ic(x, lengths):
    import bench
    return bench.initial_state(x, "fhn", lengths, None)


@pytest.fixture(scope="module")
def slab(tb, dev, oracle):
    O = oracle
    nel = (512, 512, 16)
    lengths = tuple(n * H for n in nel)
    dev.cg_set_persistent(0)                                   # what C5 runs: three kernels per iteration, TMA-staged SpMV
    md = tb.generate_mesh(tb.Hexahedron, nel, (0.0, 0.0, 0.0), lengths, device=dev)
    M = tb.B200CSRMatrix.from_mesh(dev, md)
    K = M.like()
    tb.core.assemble_mass(dev, md, M, 2, 1.0)
    tb.core.assemble_diffusion(dev, md, K, 2, tb._lib.D_TENSOR, np.diag(KAPPA), 1.0)
    mo = O.generate_grid(O.HEX8, nel, (0.0, 0.0, 0.0), lengths)
    yield dict(nel=nel, lengths=lengths, md=md, M=M, K=K, mo=mo)
    dev.cg_set_persistent(1)
    for h in (K, M, md):
        h.free()


def test_c5_shaped_slab_against_assembled_oracle(tb, dev, oracle, slab):
    O = oracle
    md, M, K, mo = slab["md"], slab["M"], slab["K"], slab["mo"]
    N = md.ndofs
    assert N == 513 * 513 * 17
    # numbering and pattern bit exact, column stream compressed (the C5 configuration of the SpMV)
    assert np.array_equal(md.download()[2], mo.celldofs)
    rp, ci = M.pattern()
    rpo, cio = mo.pattern()
    assert np.array_equal(rp, rpo) and np.array_equal(ci, cio)
    stored, col_bytes, _ = M.storage()
    assert col_bytes < 0.3 * 4 * stored, "column stream is not compressed on a structured grid"   # thin slab: 0.23; C5: 0.10
    Mo = O.assemble_mass(mo, 2, threaded=True)
    Ko = O.assemble_diffusion(mo, 2, O.D_TENSOR, np.diag(KAPPA), threaded=True)
    assert np.array_equal(M.nonzeros(), Mo) and np.array_equal(K.nonzeros(), Ko)      # ordered gather: bitwise
    ion = tb.FHNModel()
    st = tb.MonodomainStepper(dev, M, K, ion.model_id, ion.params())
    st.set_cg(O.SQRT_EPS, O.SQRT_EPS, None)
    u0 = _bench_ic(md.dof_coords(), slab["lengths"])
    assert np.array_equal(u0, _bench_ic(mo.dof_coords, slab["lengths"]))
    u = tb.B200Vector.from_host(dev, u0, 2)
    orc = O.MonodomainOracle(mo, O.FHN, O.default_params(O.FHN), Mo, Ko, threaded_blas1=True)
    uo = u0.copy()
    t = 0.0
    for s in range(5):
        it, rn, conv = st.step(u, t, DT)
        ito, rno, convo = orc.step(uo, t, DT)
        assert dev.cg_last_path() == 0
        assert conv and convo and abs(it - ito) <= 1, (s, it, ito)
        t += DT
        if s == 0:
            h = u.to_host()
            assert rel_linf(h[:N], uo[:N]) <= 1e-10, rel_linf(h[:N], uo[:N])
            assert rel_linf(h[N:], uo[N:]) <= 1e-10
    h = u.to_host()
    assert rel_linf(h[:N], uo[:N]) <= 1e-9 and rel_linf(h[N:], uo[N:]) <= 1e-9
    st.free()
    u.free()


def test_c5_shaped_slab_against_closed_form_operator(tb, dev, oracle, slab):
    """The matrix-free oracle that produces tests/golden/c5_checksum.json, on a size where the assembled oracle also
    runs: GPU (assembled on the device) vs closed-form CPU operator, node by node through the coordinates."""
    O = oracle
    md, M, K = slab["md"], slab["M"], slab["K"]
    nel, N = slab["nel"], md.ndofs
    S = O.StencilOracle(nel, H, np.array(KAPPA), O.FHN, O.default_params(O.FHN))
    x = md.dof_coords()
    idx = np.rint(x / H).astype(np.int64)
    g = idx[:, 0] + (nel[0] + 1) * (idx[:, 1] + (nel[1] + 1) * idx[:, 2])          # dof -> grid node
    ug = _bench_ic(S.node_coords(), slab["lengths"])
    u0 = np.concatenate([ug[:N][g], ug[N:][g]])
    assert np.array_equal(u0, _bench_ic(x, slab["lengths"]))
    ion = tb.FHNModel()
    st = tb.MonodomainStepper(dev, M, K, ion.model_id, ion.params())
    u = tb.B200Vector.from_host(dev, u0, 2)
    t = 0.0
    for s in range(3):
        it, rn, conv = st.step(u, t, DT)
        its, rns, convs = S.step(ug, t, DT)
        assert conv and convs and abs(it - its) <= 1
        t += DT
        h = u.to_host()
        assert rel_linf(h[:N], ug[:N][g]) <= (1e-10 if s == 0 else 1e-9)
    st.free()
    u.free()
