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
micro = tb.OrthotropicMicrostructureModel(tb.FieldCoefficient(fsn[:, :, 0]), tb.FieldCoefficient(fsn[:, :, 1]), tb.FieldCoefficient(fsn[:, :, 2]))
kappa = tb.SpectralTensorCoefficient(micro, tb.ConstantCoefficient((k1, kr, kr)))
proto = tb.AnalyticalTransmembraneStimulationProtocol(
    tb.AnalyticalCoefficient(tb.UniformEndocardialActivation(transmural_depth=0.0, tmax=0.2, amplitude=0.3), tb.CartesianCoordinateSystem()), [(-np.inf, np.inf)])
model = tb.MonodomainModel(tb.ConstantCoefficient(1.0), tb.ConstantCoefficient(1.0), kappa, proto, tb.PCG2019(), "φₘ", "s")
odeform = tb.semidiscretize(tb.ReactionDiffusionSplit(model), tb.FiniteElementDiscretization({"φₘ": tb.LagrangeCollection(1)}), mesh)
u0 = tb.create_initial_condition(odeform)
integ = tb.init(tb.OperatorSplittingProblem(odeform, u0.copy(), (0.0, 2.0)),
                tb.LieTrotterGodunov((tb.BackwardEulerSolver(inner_solver=tb.B200CG()), tb.ForwardEulerCellSolver())), dt=0.01)
data = np.concatenate([[k1, kr, kr], np.ascontiguousarray(fsn).reshape(tets.shape[0], 4, 9).ravel()])
Mo, Ko = O.assemble_mass(mo, 2), O.assemble_diffusion(mo, 2, O.D_SPECTRAL, data)
hc = integ.caches[0]
print("M bitwise", np.array_equal(hc.M.A.nonzeros(), Mo), "K bitwise", np.array_equal(hc.K.A.nonzeros(), Ko), dev.assembly_info())
orc = O.MonodomainOracle(mo, O.PCG2019, O.default_params(O.PCG2019), Mo, Ko)
uo = u0.copy(); N = mo.ndofs
t, dt = 0.0, 0.01
for step in range(100):
    orc.bS = O.assemble_source(mo, 2, O.SRC_ENDO, [0.0, 0.2, 0.3, 0.25], t + dt)
    ro = orc.step(uo, t, dt)
    ok = tb.step_(integ)
    bg = hc.source_term.b.to_host()
    h = integ.u.to_host()
    if abs(hc.iters[-1] - ro[0]) > 0 or step % 10 == 0: print(step, "ok", ok, "gpu it/rn", hc.iters[-1], hc.resid[-1], "oracle", ro, "bS rel", np.abs(bg[:N] - orc.bS).max() / np.abs(orc.bS).max(),
          "phi rel", np.abs(h[:N] - uo[:N]).max() / np.abs(uo[:N]).max(), dev.assembly_info(), flush=True)
    if not ok:
        break
    t += dt
