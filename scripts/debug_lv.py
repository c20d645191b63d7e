"""GPU debug: LV config-4 first-step mismatch (tests/test_lv_config4.py)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import oracle as O
import thunderbolt_jl_b200 as tb
from thunderbolt_jl_b200 import lv

O.build()
dev = tb.B200Device(0); tb.set_default_device(dev)
nodes, hexes, wedges, prm = lv.generate_ideal_lv_mesh(16, 3, 8)
tets = lv.tetrahedralize(nodes, hexes, wedges)
fsn = lv.odb25lt_fibres(prm, tets)
k1, kr = 0.17 * 0.62 / (0.17 + 0.62), 0.019 * 0.24 / (0.019 + 0.24)
mo = O.Mesh(O.TET4, tets, nodes)
mesh = tb.to_mesh(tb.Tetrahedron, tets, nodes, device=dev)
data = np.concatenate([[k1, kr, kr], np.ascontiguousarray(fsn).reshape(tets.shape[0], 4, 9).ravel()])
Mo, Ko = O.assemble_mass(mo, 2), O.assemble_diffusion(mo, 2, O.D_SPECTRAL, data)
M = tb.B200CSRMatrix.from_mesh(dev, mesh); K = M.like()
tb.core.assemble_mass(dev, mesh, M, 2, 1.0)
tb.core.assemble_diffusion(dev, mesh, K, 2, tb._lib.D_SPECTRAL, data, 1.0)
print("M rel", np.abs(M.nonzeros() - Mo).max() / np.abs(Mo).max(), "K rel", np.abs(K.nonzeros() - Ko).max() / np.abs(Ko).max())
N = mo.ndofs
bS = tb.B200Vector(dev, N, 1)
tb.core.assemble_source(dev, mesh, bS, 2, tb._lib.SRC_ENDO, [0.0, 2.0, 0.5, 0.25], 0.01)
bo = O.assemble_source(mo, 2, O.SRC_ENDO, [0.0, 2.0, 0.5, 0.25], 0.01)
bg = bS.to_host()
print("bS rel", np.abs(bg - bo).max() / np.abs(bo).max(), "nonzero", (bo != 0).sum(), (bg != 0).sum(), "mismatch rows", np.flatnonzero(np.abs(bg - bo) > 1e-12 * np.abs(bo).max())[:20])
rp, ci = M.pattern()
w = np.diff(rp); print("row widths min/max", w.min(), w.max(), "storage", M.storage())
# SpMV check on this irregular pattern
x = np.random.default_rng(0).standard_normal(N)
xv = tb.B200Vector.from_host(dev, x); yv = tb.B200Vector(dev, N, 1)
M.mul(yv, xv)
import scipy.sparse as sp
yo = sp.csr_matrix((Mo, ci, rp), shape=(N, N)) @ x
print("spmv rel", np.abs(yv.to_host() - yo).max() / np.abs(yo).max())
# CG alone: A = M - dt K, b = M u + bS
dt = 0.01
A = M.like(); A.axpby_values(M, K, dt)
Ao = Mo - dt * Ko
print("A rel", np.abs(A.nonzeros() - Ao).max() / np.abs(Ao).max())
u0 = np.full(N, -85.0)
b = sp.csr_matrix((Mo, ci, rp), shape=(N, N)) @ u0 + bo
bv = tb.B200Vector.from_host(dev, b); xs = tb.B200Vector(dev, N, 1)
it, rn, cv = tb.core.cg_solve(dev, A, bv, xs)
xo, ito, rno, cvo = O.cg(rp, ci, Ao, b) if hasattr(O, "cg") else (None, None, None, None)
print("gpu cg", it, rn, cv, "oracle cg", ito, rno, cvo)
if xo is not None:
    print("x rel", np.abs(xs.to_host() - xo).max() / np.abs(xo).max())
# full stepper vs oracle step, with per-step diagnostics
ion = tb.PCG2019()
st = tb.MonodomainStepper(dev, M, K, ion.model_id, ion.params())
st.set_cg(); st.set_cell_solver(1, 0.1); st.set_source(bS)
uinit = np.repeat(tb.default_initial_state(ion), N)
u = tb.B200Vector.from_host(dev, uinit, 7)
orc = O.MonodomainOracle(mo, O.PCG2019, O.default_params(O.PCG2019), Mo, Ko)
orc.bS = bo
uo = uinit.copy()
for s in range(3):
    r = st.step(u, s * dt, dt); ro = orc.step(uo, s * dt, dt)
    h = u.to_host()
    print("step", s, "gpu", r, "oracle", ro, "phi rel", np.abs(h[:N] - uo[:N]).max() / np.abs(uo[:N]).max())
