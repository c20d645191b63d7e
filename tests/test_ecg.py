"""Pseudo-ECG: Plonsey1964ECGGaussCache (src/modeling/electrophysiology/ecg.jl:55-160).
The reference's own integration test (test/integration/test_ecg.jl:5-60,91-310) gives real known answers for it --
equilibrium, idempotence, the sign and size of a planar x^3 wave, symmetry of radially symmetric stimuli -- so this
is one part of the path whose oracle IS pinned by the reference.  CPU: the oracle against those; GPU: the fused
sweep (tb_ecg_plonsey) against the oracle and against the same known answers."""
import numpy as np
import pytest

SIZE = 2.0
ELECTRODES = np.array([[0, 0, 0], [-SIZE, 0, 0], [SIZE, 0, 0], [0, -SIZE, 0], [0, SIZE, 0], [0, 0, -SIZE], [0, 0, SIZE]], float)


def _heart(O, ct):
    """generate_mesh(geo, (6,6,6)) on [-1,1]^3, then x -> sign(x) x^2 (test_ecg.jl:22-23)."""
    m = O.generate_grid(ct, (6, 6, 6), (-1, -1, -1), (1, 1, 1))
    m.coords[:] = np.sign(m.coords) * m.coords ** 2
    return m


def _reference_properties(ev, x):
    """ev(phi, points) -> phi_e; the assertions of test/integration/test_ecg.jl for the Plonsey cache."""
    signal = 0.04
    assert np.all(ev(np.zeros(x.shape[0]), ELECTRODES) == 0.0)                              # Equilibrium (:91-96)
    rng = np.random.default_rng(0)
    u = rng.standard_normal(x.shape[0])
    assert np.array_equal(ev(u, ELECTRODES), ev(u, ELECTRODES))                              # Idempotence (:114-121)
    for dim in range(3):                                                                    # Planar wave (:138-167,193-221)
        for sgn in (1.0, -1.0):
            phi = sgn * x[:, dim] ** 3
            plus, minus = np.zeros(3), np.zeros(3)
            plus[dim], minus[dim] = SIZE, -SIZE
            vp, vm = ev(phi, plus[None])[0], ev(phi, minus[None])[0]
            if sgn > 0:
                assert vp > signal and vm < signal
            else:
                assert vp < signal and vm > signal
            for d2 in range(3):
                if d2 != dim:
                    for s2 in (SIZE, -SIZE):
                        p = np.zeros(3)
                        p[d2] = s2
                        assert abs(ev(phi, p[None])[0]) <= 1e-4
    v = ev(np.sqrt(3) - np.linalg.norm(x, axis=1), ELECTRODES[1:])                          # Symmetric stimuli (:253-266)
    assert np.abs(v - v[0]).max() <= 1e-2
    v = ev(x[:, 0] ** 2, ELECTRODES[1:])                                                    # x1^2 (:286-297)
    assert abs(v[0] - v[1]) <= 1e-2 and abs(v[2] - v[3]) <= 1e-2 and abs(v[3] - v[4]) <= 1e-2 and abs(v[4] - v[5]) <= 1e-2


@pytest.mark.parametrize("name", ["HEX8", "TET4"])
def test_oracle_reproduces_the_reference_ecg_tests(oracle, name):
    O = oracle
    m = _heart(O, getattr(O, name))
    D = np.eye(3).ravel()
    _reference_properties(lambda phi, pts: O.ecg_plonsey(m, 2, O.D_TENSOR, D, phi, pts, 1.0), m.dof_coords)
    # linearity in phi and 1/kappa_t scaling
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal(m.ndofs), rng.standard_normal(m.ndofs)
    ea, eb = (O.ecg_plonsey(m, 2, O.D_TENSOR, D, v, ELECTRODES[1:], 1.0) for v in (a, b))
    eab = O.ecg_plonsey(m, 2, O.D_TENSOR, D, 2 * a - 3 * b, ELECTRODES[1:], 4.0)
    assert np.allclose(eab, (2 * ea - 3 * eb) / 4.0, rtol=1e-12, atol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["HEX8", "TET4"])
def test_gpu_ecg_vs_oracle_and_reference_properties(tb, dev, oracle, name):
    O = oracle
    ct = getattr(O, name)
    m = _heart(O, ct)
    md = tb.DeviceMesh.from_host(dev, ct, m.conn, m.coords, m.celldofs, m.ndofs)
    κ = tb.ConstantCoefficient(np.eye(3))
    op = tb.setup_operator(tb.ElementAssemblyStrategy(dev), tb.BilinearDiffusionIntegrator(κ, tb.QuadratureRuleCollection(2), "φₘ"),
                           None, md)
    cache = tb.Plonsey1964ECGGaussCache(op, np.zeros(m.ndofs))

    def ev(phi, pts):
        tb.update_ecg_(cache, phi)
        return tb.evaluate_ecg(cache, pts, 1.0)
    _reference_properties(ev, m.dof_coords)
    # against the oracle on random data, an anisotropic tensor, many electrodes (more than one launch), Cm*chi != 1
    rng = np.random.default_rng(2)
    phi = rng.standard_normal(m.ndofs)
    B = rng.standard_normal((3, 3))
    D = B @ B.T + np.eye(3)
    pts = rng.uniform(1.5, 3.0, (19, 3)) * rng.choice([-1.0, 1.0], (19, 3))
    want = O.ecg_plonsey(m, 2, O.D_TENSOR, D.ravel(), phi, pts, 0.7, cmchi=2.5)
    pv = tb.B200Vector.from_host(dev, phi)
    got = tb.core.ecg_plonsey(dev, md, 2, tb._lib.D_TENSOR, D, pv, pts, 0.7, cm_chi=2.5)
    assert np.allclose(got, want, rtol=1e-11, atol=1e-13 * np.abs(want).max())
    assert isinstance(tb.evaluate_ecg(cache, pts[0], 1.0), float)
    # spectral (fibre) coefficient, order-3 quadrature
    data = np.concatenate([[0.3, 0.1, 0.05], rng.standard_normal((m.ncells, m.nv, 9)).ravel()])
    want = O.ecg_plonsey(m, 3 if name == "HEX8" else 2, O.D_SPECTRAL, data, phi, pts[:3], 1.0)
    got = tb.core.ecg_plonsey(dev, md, 3 if name == "HEX8" else 2, tb._lib.D_SPECTRAL, data, pv, pts[:3], 1.0)
    assert np.allclose(got, want, rtol=1e-11, atol=1e-13 * np.abs(want).max())
    with pytest.raises(ValueError):
        tb.core.ecg_plonsey(dev, md, 2, tb._lib.D_SCALAR, [1.0], pv, np.zeros((1, 2)), 1.0)
    pv.free(); md.free()


@pytest.mark.gpu
def test_gpu_ecg_large_mesh_properties(tb, dev):
    """Size-independent properties on a 2 M element slab (no oracle): exact zero at rest, determinism, and the mirror
    symmetry phi(x) -> phi(-x) flips the sign seen by mirrored electrodes."""
    md = tb.generate_mesh(tb.Hexahedron, (160, 128, 96), (-1.0, -0.8, -0.6), (1.0, 0.8, 0.6), device=dev)
    x = md.dof_coords()
    D = np.diag([0.1334, 0.0176, 0.0176])
    pts = np.array([[3.0, 0.2, 0.1], [-3.0, -0.2, -0.1]])
    pv = tb.B200Vector.from_host(dev, np.zeros(md.ndofs))
    assert np.all(tb.core.ecg_plonsey(dev, md, 2, tb._lib.D_TENSOR, D, pv, pts, 1.0) == 0.0)
    pv.upload(np.tanh(4 * x[:, 0]) + 0.3 * x[:, 1])                 # odd in x: phi(-x) = -phi(x)
    a = tb.core.ecg_plonsey(dev, md, 2, tb._lib.D_TENSOR, D, pv, pts, 1.0)
    b = tb.core.ecg_plonsey(dev, md, 2, tb._lib.D_TENSOR, D, pv, pts, 1.0)
    assert np.array_equal(a, b) and abs(a[0]) > 1e-4
    assert a[0] == pytest.approx(-a[1], rel=1e-9)
    pv.free(); md.free()
