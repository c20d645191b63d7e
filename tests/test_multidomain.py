"""Multi-subdomain reaction-diffusion split: PointwiseMultiODEFunction with PointBlockedLayout blocks, the
scattered heat_dofrange view, BilinearInterfaceDiffusionIntegrator -- the reference's "Pacemaker subdomain" case
(test/integration/test_electrophysiology.jl:124-195).  CPU part: host logic + oracle element kernel; GPU part: parity."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _pacemaker_grid(tb, O, n=64):
    """generate_grid(Quadrilateral, (n, n), (-2.5, -2.5), (2.5, 2.5)); Pacemaker = cells with all nodes |x|_inf <= 0.75"""
    g = O.generate_grid(O.QUAD4, (n, n), (-2.5, -2.5), (2.5, 2.5))
    xc = g.coords[g.conn]                                             # addcellset!(grid, name, x -> ...) : all nodes of the cell
    pace = np.flatnonzero((np.abs(xc).max(axis=2) <= 0.75).all(axis=1))
    myo = np.setdiff1d(np.arange(g.ncells), pace)
    return tb.insert_interfaces(tb.Quadrilateral, g.conn, g.coords, {"Pacemaker": pace, "Myocardium": myo}, ("Pacemaker", "Myocardium"))


def _models(tb):
    coeff = tb.ConstantCoefficient(tb.SymmetricTensor(2, (4.5e-4, 0, 2.0e-4)))
    one = tb.ConstantCoefficient(1.0)
    pace = tb.FHNModel(a=-0.5, b=1.0, c=-0.6, d=0.0, e=0.001, f=50 * 0.001)
    return {"Pacemaker": tb.MonodomainModel(one, one, coeff, tb.NoStimulationProtocol(), pace, "φₘ", "s1"),
            "Myocardium": tb.MonodomainModel(one, one, coeff, tb.NoStimulationProtocol(), tb.FHNModel(), "φₘ", "s2"),
            "interfaces": tb.InterfaceDiffusionModel(tb.ConstantCoefficient(1.0), "φₘ", "φₘi")}


def test_state_range_layouts(tb):
    """solution_variables.jl:60-68"""
    b = tb.StateBlock(10, 4, 3, tb.PointBlockedLayout())
    assert list(tb.state_range(b, 1)) == [11, 12, 13] and list(tb.state_range(b, 4)) == [20, 21, 22]
    b = tb.StateBlock(10, 4, 3, tb.StateBlockedLayout())
    assert list(tb.state_range(b, 1)) == [11, 15, 19] and list(tb.state_range(b, 4)) == [14, 18, 22]


def test_insert_interfaces_duplicates_the_common_boundary(tb, oracle):
    g = _pacemaker_grid(tb, oracle, 16)
    here, there = g.interfaces["interfaces"]
    pace, myo = g.subdomains["Pacemaker"], g.subdomains["Myocardium"]
    assert np.intersect1d(np.unique(g.cells[pace]), np.unique(g.cells[myo])).size == 0          # subdomains isolated
    assert np.array_equal(g.nodes[here], g.nodes[there]) and not np.intersect1d(here, there).size   # coincident copies
    # the pacemaker block of a 16x16 grid on [-2.5, 2.5]^2 with |x| <= 0.75: 4 x 4 cells -> 16 boundary edges, 16 boundary nodes
    assert pace.size == 16 and here.shape == (16, 2) and g.nodes.shape[0] == 17 * 17 + 16
    # every interface edge belongs to one pacemaker cell (here) and one myocardium cell (there)
    for h, t_ in zip(here, there):
        assert any(set(h) <= set(c) for c in g.cells[pace]) and any(set(t_) <= set(c) for c in g.cells[myo])


def test_interface_element_matrix_closed_form(oracle):
    """straight interface edge of length l, D: K_e = -D * l * [[1/3, 1/6], [1/6, 1/3]] (x) [[1, -1], [-1, 1]] (jump blocks)"""
    import ctypes as C
    O = oracle
    L_ = O.lib()
    f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L_.orc_interface_diffusion_element.argtypes = [C.c_int, C.c_int, C.c_int, f64, f64, C.c_double, f64]
    X = np.array([[0.0, 0.0], [0.6, 0.8]])                      # length 1
    Ke = np.empty(16)
    L_.orc_interface_diffusion_element(2, 2, 2, X.ravel(), X.ravel(), 2.5, Ke)
    m = np.array([[1 / 3, 1 / 6], [1 / 6, 1 / 3]])
    ref = -2.5 * np.block([[m, -m], [-m, m]])
    assert np.allclose(Ke.reshape(4, 4), ref, rtol=0, atol=1e-15)
    assert abs(Ke.sum()) < 1e-15                                  # constants are in the kernel: no flux without a jump


def _oracle_operators(O, grid, order, celldofs, ndofs, rowptr, colidx, D2, G):
    """M, K on the union pattern: bulk subdomains in `order` (sequential element loop each), then the interface cells"""
    import ctypes as C
    L_ = O.lib()
    i64 = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
    f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L_.orc_assemble_interface_diffusion.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, i64, f64, f64, C.c_double, i64, i64, f64]
    L_.orc_assemble_interface_diffusion.restype = C.c_int
    M, K = np.zeros(colidx.size), np.zeros(colidx.size)
    for name in order:
        ids = grid.subdomains[name]
        conn, cd = np.ascontiguousarray(grid.cells[ids]), np.ascontiguousarray(celldofs[ids])
        L_.orc_assemble_bilinear(0, O.QUAD4, 2, ids.size, conn, np.ascontiguousarray(grid.nodes), cd, 1.0, 0, np.zeros(1), 1.0, rowptr, colidx, M)
        L_.orc_assemble_bilinear(1, O.QUAD4, 2, ids.size, conn, np.ascontiguousarray(grid.nodes), cd, 1.0, O.D_TENSOR, np.ascontiguousarray(D2.ravel()), 1.0, rowptr, colidx, K)
    return M, K


@pytest.mark.gpu
@pytest.mark.parametrize("cell_solver,tight", [("fe", False), ("adaptive", False), ("fe", True)])
def test_pacemaker_subdomain_against_oracle(tb, dev, oracle, cell_solver, tight):
    """tight = False: LinearSolve's default tolerances -- CG iterations +-1, phi to the CG stopping error; tight = True: the
    linear solves converged far below 1e-10 on both sides, then the 1e-10-after-one-step rule applies to the whole split step"""
    O = oracle
    tol = dict(atol=1e-14, rtol=1e-13) if tight else {}
    grid = _pacemaker_grid(tb, O, 64)
    models = _models(tb)
    odeform = tb.semidiscretize(tb.ReactionDiffusionSplit(models),
                                tb.FiniteElementDiscretization({"φₘ": tb.LagrangeCollection(1), "φₘi": tb.LagrangeCollection(1)}), grid)
    heat, ionic = odeform.functions
    hd = np.asarray(odeform.solution_indices[0])
    n = heat.ndofs
    assert n == grid.nodes.shape[0] and tb.solution_size(odeform) == 2 * n
    # index contract (fem.jl:472-521): blocks packed in subdomain order, point blocked, phi_m = first state of every point
    b0, b1 = ionic.functions[0].block, ionic.functions[1].block
    assert (b0.offset, b0.nstates, b1.offset) == (0, 2, 2 * b0.npoints) and b0.npoints + b1.npoints == n
    assert np.array_equal(np.sort(hd), 1 + 2 * np.arange(n))
    # simple_initializer!: phi_0 = max(1 - |x|, 0)
    u0 = tb.create_initial_condition(odeform)
    u0[hd - 1] = np.maximum(1.0 - np.linalg.norm(heat.dof_coords, axis=1), 0.0)
    cs = tb.ForwardEulerCellSolver() if cell_solver == "fe" else tb.AdaptiveForwardEulerSubstepper()
    integ = tb.init(tb.OperatorSplittingProblem(odeform, u0.copy(), (0.0, 10.0)),
                    tb.LieTrotterGodunov((tb.BackwardEulerSolver(inner_solver=tb.B200CG(**tol)), cs)), dt=1.0)
    # ---- oracle: same numbering, same operators, same split -------------------------------------------------------
    rowptr, colidx = heat.rowptr, heat.colidx
    D2 = np.array([[4.5e-4, 0.0], [0.0, 2.0e-4]])
    Mo, Ko = _oracle_operators(O, grid, ["Pacemaker", "Myocardium"], heat.celldofs, n, rowptr, colidx, D2, 1.0)
    name, d, xh, xt, G, q = heat.interfaces[0]
    miss = O.lib().orc_assemble_interface_diffusion(2, 2, q, d.shape[0], np.ascontiguousarray(d), np.ascontiguousarray(xh),
                                                    np.ascontiguousarray(xt), G, rowptr, colidx, Ko)
    assert miss == 0
    assert np.array_equal(integ.M.pattern()[1], colidx)
    assert np.abs(integ.M.nonzeros() - Mo).max() <= 1e-15 * np.abs(Mo).max()
    assert np.abs(integ.K.nonzeros() - Ko).max() <= 1e-14 * np.abs(Ko).max()
    import ctypes as C
    L_ = O.lib()
    f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L_.orc_cell_step_strided.argtypes = [C.c_int, f64, f64, f64, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int]
    uo, du = u0.copy(), np.zeros_like(u0)
    Ao = O.axpby_values(Mo, Ko, 1.0)
    sub = 10 if cell_solver == "adaptive" else 1
    prm = [np.ascontiguousarray(f.ode.params(), dtype=np.float64) for f in ionic.functions]
    for step in range(10):
        assert tb.step_(integ)
        phi = uo[hd - 1].copy()
        x, ito, rno, convo = O.cg(rowptr, colidx, Ao, O.spmv(rowptr, colidx, Mo, phi), **tol)
        assert convo and abs(integ.cg_iterations[-1] - ito) <= (3 if tight else 1)
        uo[hd - 1] = x
        for f, p in zip(ionic.functions, prm):
            b = f.block
            seg = uo[b.offset:b.offset + 2 * b.npoints]
            dseg = du[b.offset:b.offset + 2 * b.npoints]
            L_.orc_cell_step_strided(O.FHN, p, seg, dseg, b.npoints, 2, 1, float(step), 1.0, sub, 0.1, 0)
        if step == 0:
            h = integ.u.to_host()
            assert np.abs(h - uo).max() <= (1e-10 if tight else 1e-7) * np.abs(uo).max()
    h = integ.u.to_host()
    assert np.abs(h - uo).max() <= (1e-9 if tight else 1e-6) * np.abs(uo).max()
    assert not np.allclose(h, u0)                                     # `integrator.u ≉ u₀`
    assert integ.stats.naccept == 10 and integ.t == 10.0


@pytest.mark.gpu
def test_pacemaker_reference_assertions(tb, dev, oracle):
    """the reference's own checks: FE vs adaptive agree to rtol 1e-3; the RTC run takes a different number of steps and
    stays within 5e-2 (test_electrophysiology.jl:166-176)"""
    grid = _pacemaker_grid(tb, oracle, 64)
    odeform = tb.semidiscretize(tb.ReactionDiffusionSplit(_models(tb)),
                                tb.FiniteElementDiscretization({"φₘ": tb.LagrangeCollection(1), "φₘi": tb.LagrangeCollection(1)}), grid)
    heat = odeform.functions[0]
    hd = np.asarray(odeform.solution_indices[0])
    u0 = tb.create_initial_condition(odeform)
    u0[hd - 1] = np.maximum(1.0 - np.linalg.norm(heat.dof_coords, axis=1), 0.0)
    ltg = tb.LieTrotterGodunov((tb.BackwardEulerSolver(), tb.ForwardEulerCellSolver()))
    runs = {}
    for name, alg in (("fe", ltg), ("adaptive", tb.LieTrotterGodunov((tb.BackwardEulerSolver(), tb.AdaptiveForwardEulerSubstepper()))),
                      ("rtc", tb.ReactionTangentController(ltg, 0.5, 1.0, (0.5, 2.0)))):
        integ = tb.init(tb.OperatorSplittingProblem(odeform, u0.copy(), (0.0, 10.0)), alg, dt=1.0)
        while integ.t < 10.0 - 1e-12:
            assert tb.step_(integ)
        runs[name] = (integ.u.to_host(), integ.stats.naccept)
    a, b, c = runs["fe"][0], runs["adaptive"][0], runs["rtc"][0]
    assert np.linalg.norm(a - b) <= 1e-3 * max(np.linalg.norm(a), np.linalg.norm(b))
    assert np.linalg.norm(a - c) <= 5e-2 * max(np.linalg.norm(a), np.linalg.norm(c))
    assert runs["rtc"][1] != runs["fe"][1]


@pytest.mark.gpu
def test_blocked_cell_sweep_layouts_and_models(tb, dev, oracle):
    """tb_cell_step_blocks: three blocks (FHN point-blocked, PCG2019 point-blocked, Aliev-Panfilov state-blocked) in one
    flat vector against the oracle's strided sweep, FE and adaptive"""
    import ctypes as C
    O = oracle
    Lp = tb._lib
    L_ = O.lib()
    f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L_.orc_cell_step_strided.argtypes = [C.c_int, f64, f64, f64, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int]
    rng = np.random.default_rng(11)
    spec = [(Lp.FHN, O.FHN, 1001, 2, Lp.LAYOUT_POINT_BLOCKED, 0), (Lp.PCG2019, O.PCG2019, 777, 7, Lp.LAYOUT_POINT_BLOCKED, 0),
            (Lp.ALIEV_PANFILOV, O.ALIEV_PANFILOV, 530, 2, Lp.LAYOUT_STATE_BLOCKED, 1)]
    total = sum(npts * ns for _, _, npts, ns, _, _ in spec)
    u0 = np.zeros(total)
    blocks = (Lp.CellBlock * 3)()
    off = 0
    for i, (mid, omid, npts, ns, lay, phi) in enumerate(spec):
        prm = O.default_params(omid)
        st0 = O.default_initial_state(omid)
        pts = np.tile(st0, (npts, 1)) + (0.0 if omid == O.PCG2019 else 0.3 * rng.random((npts, ns)))
        if omid == O.PCG2019:
            pts[:, 0] += 60.0 * rng.random(npts)
        u0[off:off + npts * ns] = pts.ravel() if lay == Lp.LAYOUT_POINT_BLOCKED else pts.T.ravel()
        blocks[i].offset, blocks[i].npoints, blocks[i].model, blocks[i].layout, blocks[i].nparams = off, npts, mid, lay, prm.size
        for j, v in enumerate(prm):
            blocks[i].params[j] = float(v)
        off += npts * ns
    for sub in (1, 10):
        u = tb.B200Vector.from_host(dev, u0, 1)
        uo, du = u0.copy(), np.zeros(total)
        R = C.c_double()
        for s in range(5):
            Lp.call("tb_cell_step_blocks", dev.h, blocks, 3, u.h, 0.01 * s, 0.01, sub, 0.1, C.byref(R))
            off, Ro = 0, -np.inf
            for mid, omid, npts, ns, lay, phi in spec:
                ps, ss = (ns, 1) if lay == Lp.LAYOUT_POINT_BLOCKED else (1, npts)
                seg, dseg = uo[off:off + npts * ns], du[off:off + npts * ns]
                L_.orc_cell_step_strided(omid, O.default_params(omid), seg, dseg, npts, ps, ss, 0.01 * s, 0.01, sub, 0.1, phi)
                dphi = dseg[phi::ns] if lay == Lp.LAYOUT_POINT_BLOCKED else dseg[phi * npts:(phi + 1) * npts]
                Ro = max(Ro, dphi.max())
                off += npts * ns
            assert abs(R.value - Ro) <= 1e-12 * abs(Ro)
        assert np.abs(u.to_host() - uo).max() <= 1e-11 * np.abs(uo).max()
        u.free()


@pytest.mark.gpu
def test_subdomains_with_different_cell_models(tb, dev, oracle):
    """test/test_solution_variables.jl:162-215 ("Subdomains with different cell models": FHN with 2 states next to PCG2019
    with 7) -- the index contract -- and then 5 split steps of that mixed problem against the oracle's composition."""
    import ctypes as C
    O = oracle
    g = O.generate_grid(O.QUAD4, (8, 8), (-1.0, -1.0), (1.0, 1.0))
    xc = g.coords[g.conn]
    fast = np.flatnonzero((np.abs(xc).max(axis=2) <= 0.5).all(axis=1))
    slow = np.setdiff1d(np.arange(g.ncells), fast)
    grid = tb.insert_interfaces(tb.Quadrilateral, g.conn, g.coords, {"Fast": fast, "Slow": slow}, ("Fast", "Slow"))
    coeff = tb.ConstantCoefficient(tb.SymmetricTensor(2, (1.0e-4, 0, 1.0e-4)))
    one = tb.ConstantCoefficient(1.0)
    models = {"Fast": tb.MonodomainModel(one, one, coeff, tb.NoStimulationProtocol(), tb.FHNModel(), "φₘ", "sfast"),
              "Slow": tb.MonodomainModel(one, one, coeff, tb.NoStimulationProtocol(), tb.PCG2019(), "φₘ", "sslow"),
              "interfaces": tb.InterfaceDiffusionModel(tb.ConstantCoefficient(1.0), "φₘ", "φₘi")}
    form = tb.semidiscretize(tb.ReactionDiffusionSplit(models),
                             tb.FiniteElementDiscretization({"φₘ": tb.LagrangeCollection(1), "φₘi": tb.LagrangeCollection(1)}), grid)
    heat, ionic = form.functions
    n = heat.ndofs
    md = tb.multidomain
    iφ = md.solution_indices(form, "φₘ", tb.api)
    assert iφ.size == n and np.array_equal(iφ, np.asarray(form.solution_indices[0]))
    nf, ns_ = ionic.functions[0].block.npoints, ionic.functions[1].block.npoints
    assert tb.solution_size(form) == 2 * nf + 7 * ns_ and nf + ns_ == n
    ifast, islow = md.solution_indices(form, "sfast", tb.api), md.solution_indices(form, "sslow", tb.api)
    assert ifast.size == nf and islow.size == 6 * ns_
    assert np.intersect1d(ifast, islow).size == 0 and np.intersect1d(ifast, iφ).size == 0 and np.intersect1d(islow, iφ).size == 0
    assert np.array_equal(np.sort(np.concatenate([iφ, ifast, islow])), np.arange(1, tb.solution_size(form) + 1))
    u0 = tb.create_initial_condition(form)
    exp = tb.default_initial_state(tb.PCG2019())
    assert np.allclose(u0[islow - 1].reshape(-1, 6), exp[1:])             # point major: (ncomponents per point)
    before = u0[islow - 1].copy()
    u0[ifast - 1] = -1.0                                                   # writing one subdomain's state ...
    assert np.array_equal(u0[islow - 1], before)                           # ... does not touch the other's
    u0[ifast - 1] = 0.0
    # a depolarised patch inside the FHN block, PCG2019 tissue at rest around it
    xd = heat.dof_coords
    blk0 = ionic.functions[0]
    u0[iφ[blk0.subdofs] - 1] = np.maximum(1.0 - 2.0 * np.linalg.norm(xd[blk0.subdofs], axis=1), 0.0)
    tol = dict(atol=1e-14, rtol=1e-13)
    integ = tb.init(tb.OperatorSplittingProblem(form, u0.copy(), (0.0, 0.05)),
                    tb.LieTrotterGodunov((tb.BackwardEulerSolver(inner_solver=tb.B200CG(**tol)), tb.AdaptiveForwardEulerSubstepper())), dt=0.01)
    rowptr, colidx = heat.rowptr, heat.colidx
    D2 = np.array([[1.0e-4, 0.0], [0.0, 1.0e-4]])
    Mo, Ko = _oracle_operators(O, grid, ["Fast", "Slow"], heat.celldofs, n, rowptr, colidx, D2, 1.0)
    name, d, xh, xt, G, q = heat.interfaces[0]
    assert O.lib().orc_assemble_interface_diffusion(2, 2, q, d.shape[0], np.ascontiguousarray(d), np.ascontiguousarray(xh),
                                                    np.ascontiguousarray(xt), G, rowptr, colidx, Ko) == 0
    L_ = O.lib()
    f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L_.orc_cell_step_strided.argtypes = [C.c_int, f64, f64, f64, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int]
    uo, du = u0.copy(), np.zeros_like(u0)
    Ao = O.axpby_values(Mo, Ko, 0.01)
    for step in range(5):
        assert tb.step_(integ)
        x, ito, rno, convo = O.cg(rowptr, colidx, Ao, O.spmv(rowptr, colidx, Mo, uo[iφ - 1].copy()), **tol)
        uo[iφ - 1] = x
        for f, om in zip(ionic.functions, (O.FHN, O.PCG2019)):
            b = f.block
            seg, dseg = uo[b.offset:b.offset + b.nstates * b.npoints], du[b.offset:b.offset + b.nstates * b.npoints]
            L_.orc_cell_step_strided(om, O.default_params(om), seg, dseg, b.npoints, b.nstates, 1, 0.01 * step, 0.01, 10, 0.1, 0)
    h = integ.u.to_host()
    scale = np.maximum(np.abs(uo), 1.0)
    assert (np.abs(h - uo) / scale).max() <= 1e-9


@pytest.mark.gpu
def test_gather_scatter_properties_of_the_reference(tb, dev):
    """tb_vec_gather / tb_vec_scatter against the properties the reference asserts of its own gather! / scatter!
    (test/test_solution_vector_mapping.jl:82-110): gather then scatter is a round trip on the wired entries and leaves
    zeros elsewhere in a wiped source; scatter touches nothing outside the wiring."""
    import ctypes as C
    from thunderbolt_jl_b200 import _lib as L
    rng = np.random.default_rng(12)
    nsource, ntarget = 1000, 317
    wiring = np.sort(rng.choice(nsource, ntarget, replace=False)).astype(np.int64) + 1          # 1-based like the reference's dof ids
    h = C.c_void_p()
    L.call("tb_index_create", dev.h, L.ptr(wiring), int(wiring.size), 1, C.byref(h))
    try:
        original = np.arange(1.0, nsource + 1)
        source, target = tb.B200Vector.from_host(dev, original), tb.B200Vector(dev, ntarget)
        L.call("tb_vec_gather", target.h, 0, source.h, 0, h)
        assert np.array_equal(target.to_host(), original[wiring - 1])
        source.fill(0.0)                                                                          # a scatter that failed to write shows as a zero
        L.call("tb_vec_scatter", source.h, 0, h, target.h, 0)
        got = source.to_host()
        assert np.array_equal(got[wiring - 1], original[wiring - 1])
        untouched = np.setdiff1d(np.arange(nsource), wiring - 1)
        assert np.all(got[untouched] == 0.0)
        source.upload(original)
        target.fill(-1.0)
        L.call("tb_vec_scatter", source.h, 0, h, target.h, 0)
        got = source.to_host()
        assert np.array_equal(got[untouched], original[untouched]) and np.all(got[wiring - 1] == -1.0)
        source.free(); target.free()
    finally:
        L.call("tb_index_destroy", h)
