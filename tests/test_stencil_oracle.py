"""CPU: the oracle's matrix-free closed-form operator (tb_oracle.c: orc_stencil_*) against its assembled CSR image, the
threaded oracle assembly against the sequential element loop, the committed full-size C5 golden checksums, and the
closed-form tables bench.py's parity block uses for its stencil probes."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _grid_ids(m, nel, h):
    idx = np.rint(m.dof_coords / h).astype(np.int64)
    return idx, idx[:, 0] + (nel[0] + 1) * (idx[:, 1] + (nel[1] + 1) * idx[:, 2])


@pytest.mark.parametrize("nel,h", [((6, 5, 4), 0.25), ((3, 3, 3), 0.1), ((9, 2, 5), 0.5)])
def test_closed_form_operator_equals_assembled(oracle, nel, h):
    O = oracle
    kap = np.array([0.0295, 0.0131, 0.0171])
    m = O.generate_grid(O.HEX8, nel, (0, 0, 0), tuple(h * n for n in nel))
    rp, ci = m.pattern()
    Mv, Kv = O.assemble_mass(m, 2), O.assemble_diffusion(m, 2, O.D_TENSOR, np.diag(kap))
    S = O.StencilOracle(nel, h, kap, O.FHN, O.default_params(O.FHN))
    idx, g = _grid_ids(m, nel, h)
    assert np.array_equal(np.sort(g), np.arange(m.ndofs))
    x = np.random.default_rng(1).standard_normal(m.ndofs)
    xg = np.empty(m.ndofs)
    xg[g] = x
    for cm, ck, vals in ((1.0, 0.0, Mv), (0.0, 1.0, Kv), (1.0, -0.7, Mv - 0.7 * Kv)):
        y = O.spmv(rp, ci, vals, x)
        assert np.abs(S.apply(xg, cm, ck)[g] - y).max() <= 1e-13 * np.abs(y).max()
    # entry by entry
    for r in (0, m.ndofs // 2, m.ndofs - 1):
        for k in range(rp[r], rp[r + 1]):
            off = idx[ci[k]] - idx[r]
            assert abs(S.entry(0, idx[r], off) - Mv[k]) <= 1e-14 * np.abs(Mv).max()
            assert abs(S.entry(1, idx[r], off) - Kv[k]) <= 1e-13 * np.abs(Kv).max()


def test_closed_form_steps_equal_assembled_steps(oracle):
    O = oracle
    nel, h = (8, 8, 4), 0.25
    kap = np.array([0.0295, 0.0131, 0.0131])
    m = O.generate_grid(O.HEX8, nel, (0, 0, 0), tuple(h * n for n in nel))
    Mv, Kv = O.assemble_mass(m, 2), O.assemble_diffusion(m, 2, O.D_TENSOR, np.diag(kap))
    S = O.StencilOracle(nel, h, kap, O.FHN, O.default_params(O.FHN))
    _, g = _grid_ids(m, nel, h)
    xc, n = m.dof_coords, m.ndofs
    u = np.concatenate([np.where((xc[:, 0] <= 1.0) & (xc[:, 1] <= 1.0), 1.0, 0.0), np.where(xc[:, 1] >= 1.0, 0.1, 0.0)])
    ug = np.empty(2 * n)
    ug[g], ug[n + g] = u[:n], u[n:]
    orc = O.MonodomainOracle(m, O.FHN, O.default_params(O.FHN), Mv, Kv)
    for s in range(10):
        a, b = orc.step(u, float(s), 1.0), S.step(ug, float(s), 1.0)
        assert a[0] == b[0] and a[2] and b[2]
        assert np.abs(ug[g] - u[:n]).max() <= 1e-12 and np.abs(ug[n + g] - u[n:]).max() <= 1e-12


@pytest.mark.parametrize("ct,nel", [("HEX8", (9, 7, 5)), ("TET4", (5, 4, 3)), ("QUAD4", (13, 9)), ("TRI3", (6, 7))])
def test_threaded_assembly_is_bitwise_the_sequential_loop(oracle, ct, nel):
    O = oracle
    dim = len(nel)
    m = O.generate_grid(getattr(O, ct), nel, (0,) * dim, tuple(0.3 * n for n in nel))
    D = np.diag([0.0295, 0.0131, 0.0171])[:dim, :dim]
    assert np.array_equal(O.assemble_mass(m, 2), O.assemble_mass(m, 2, threaded=True))
    assert np.array_equal(O.assemble_diffusion(m, 2, O.D_TENSOR, D), O.assemble_diffusion(m, 2, O.D_TENSOR, D, threaded=True))


def test_c5_golden_checksum_file_is_complete():
    """The committed full-size golden (generated on the CPU by tests/golden/make_c5_checksum.py) has what bench.py reads."""
    g = json.loads((ROOT / "tests" / "golden" / "c5_checksum.json").read_text())
    assert g["nel"] == [512, 512, 384] and g["dofs"] == 101320065 and len(g["steps"]) == 3
    assert all(len(s["phi_samples"]) == len(g["sample_nodes"]) for s in g["steps"])
    assert all(s["converged"] for s in g["steps"])
    import bench
    W = bench.WORKLOADS["c5"]
    assert list(W["nel"]) == g["nel"] and W["h"] == g["h"] and list(W["kappa"]) == g["kappa"] and W["dt"] == g["dt"]


def test_parity_block_tables_match_the_oracle_closed_form(oracle):
    """bench.py's stencil probes use their own numpy restatement of the 1D tables (the product side must not import the
    oracle): it has to agree with orc_stencil_entry."""
    from scripts import parity_block as pb
    O = oracle
    nel, h = (7, 5, 9), 0.25
    kap = np.array([0.0295, 0.0131, 0.0171])
    S = O.StencilOracle(nel, h, kap, O.FHN, O.default_params(O.FHN))
    T = [pb._tables_1d(n, h) for n in nel]
    rng = np.random.default_rng(0)
    for _ in range(200):
        node = np.array([rng.integers(0, n + 1) for n in nel])
        off = rng.integers(-1, 2, 3)
        if np.any(node + off < 0) or np.any(node + off > np.array(nel)):
            continue
        m1 = [T[d][0][node[d], off[d] + 1] for d in range(3)]
        k1 = [T[d][1][node[d], off[d] + 1] for d in range(3)]
        Me = m1[0] * m1[1] * m1[2]
        Ke = -(kap[0] * k1[0] * m1[1] * m1[2] + kap[1] * m1[0] * k1[1] * m1[2] + kap[2] * m1[0] * m1[1] * k1[2])
        assert abs(Me - S.entry(0, node, off)) <= 1e-18 and abs(Ke - S.entry(1, node, off)) <= 1e-16
    for n in (1, 2, 3, 5, 8, 16, 384, 512):
        P = pb._probe_positions(n)
        assert P[0] == 0 and (n < 3 or P[-1] == n) and (len(P) < 2 or np.diff(P).min() >= 3)
