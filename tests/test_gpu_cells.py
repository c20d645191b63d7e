"""-m gpu: ionic cell sweep through the C ABI vs the oracle.
FHN is polynomial: with -fmad=false the GPU result is BITWISE the oracle's.  PCG2019 differs only
through exp (CUDA's exp vs glibc's, <= 1 ulp each): relative tolerance 1e-13 per step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _fhn_state(n, seed=0):
    rng = np.random.default_rng(seed)
    return np.concatenate([rng.uniform(-0.3, 1.2, n), rng.uniform(-0.1, 0.3, n)])


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 1000, 66049])
@pytest.mark.parametrize("sub", [1, 10])
def test_fhn_bitwise(tb, dev, oracle, n, sub):
    O = oracle
    prm = O.default_params(O.FHN)
    u = _fhn_state(n)
    ud = tb.B200Vector.from_host(dev, u, 2)
    t = 0.0
    for step in range(3):
        O.cell_step(O.FHN, prm, u, n, t, 1.0, substeps=sub, threshold=0.1)
        tb.core.cell_step(dev, tb._lib.FHN, prm, ud, t, 1.0, substeps=sub, threshold=0.1)
        t += 1.0
    assert np.array_equal(ud.to_host(), u)
    ud.free()


@pytest.mark.parametrize("n", [1, 33, 4097])
@pytest.mark.parametrize("sub", [1, 10])
def test_aliev_panfilov_bitwise(tb, dev, oracle, n, sub):
    """The one model with phi_m NOT in state 0 (aliev-panfilov.jl:13-14: state_symbols = (s, phi_m))."""
    O = oracle
    prm = O.default_params(O.ALIEV_PANFILOV)
    rng = np.random.default_rng(5)
    u = np.concatenate([rng.uniform(0.0, 2.5, n), rng.uniform(-0.1, 1.1, n)])      # column 0 = s, column 1 = phi
    ud = tb.B200Vector.from_host(dev, u, 2)
    t = 0.0
    for step in range(3):
        du = O.cell_step(O.ALIEV_PANFILOV, prm, u, n, t, 0.5, substeps=sub, threshold=0.1, phi_idx=1)
        R = tb.core.cell_step(dev, tb._lib.ALIEV_PANFILOV, prm, ud, t, 0.5, substeps=sub, threshold=0.1, phi_idx=1, want_max=True)
        assert R == du[n:].max()                                                   # reaction tangent on the phi column
        t += 0.5
    assert np.array_equal(ud.to_host(), u)
    with pytest.raises(tb.TBError):
        tb.core.cell_step(dev, tb._lib.ALIEV_PANFILOV, prm, ud, t, 0.5, phi_idx=0)
    ud.free()


def test_fhn_adaptive_branch_is_exercised(tb, dev, oracle):
    O = oracle
    prm = O.default_params(O.FHN)
    u = _fhn_state(4096, 3)
    a, b = tb.B200Vector.from_host(dev, u, 2), tb.B200Vector.from_host(dev, u, 2)
    tb.core.cell_step(dev, 0, prm, a, 0.0, 1.0, substeps=1)
    tb.core.cell_step(dev, 0, prm, b, 0.0, 1.0, substeps=10, threshold=0.1)
    ha, hb = a.to_host(), b.to_host()
    du0 = np.abs(u[:4096] * (1 - u[:4096]) * (u[:4096] - 0.1) - u[4096:])
    calm = du0 < 0.1
    assert calm.any() and (~calm).any()
    assert np.array_equal(ha[:4096][calm], hb[:4096][calm])          # below threshold: bitwise a plain FE step
    assert not np.allclose(ha[:4096][~calm], hb[:4096][~calm], rtol=1e-8, atol=0)
    a.free(); b.free()


@pytest.mark.parametrize("n", [1, 7, 64, 549153 // 16])
@pytest.mark.parametrize("sub", [1, 10])
def test_pcg2019(tb, dev, oracle, n, sub):
    O = oracle
    prm = O.default_params(O.PCG2019)
    rng = np.random.default_rng(n)
    u0 = O.default_initial_state(O.PCG2019)
    u = np.repeat(u0, n)
    u[:n] = rng.uniform(-90.0, 40.0, n)
    for s in range(1, 7):
        u[s * n:(s + 1) * n] = np.clip(u[s * n:(s + 1) * n] + rng.uniform(-0.2, 0.2, n), 0.0, 1.0)
    ud = tb.B200Vector.from_host(dev, u, 7)
    O.cell_step(O.PCG2019, prm, u, n, 0.0, 0.01, substeps=sub, threshold=0.1)
    tb.core.cell_step(dev, tb._lib.PCG2019, prm, ud, 0.0, 0.01, substeps=sub, threshold=0.1)
    h = ud.to_host()
    scale = np.repeat(np.maximum(np.abs(u.reshape(7, n)).max(axis=1), 1e-3), n)
    assert np.abs(h - u).max() <= 1e-12 and np.all(np.abs(h - u) <= 1e-13 * scale + 1e-15)
    ud.free()


def test_pcg2019_resting_state_stays(tb, dev, oracle):
    """default_initial_state is the resting state: gates do not move, phi drifts by the tiny net current."""
    O = oracle
    n = 1000
    u = np.repeat(O.default_initial_state(O.PCG2019), n)
    ud = tb.B200Vector.from_host(dev, u, 7)
    tb.core.cell_step(dev, 1, O.default_params(O.PCG2019), ud, 0.0, 0.01)
    h = ud.to_host()
    assert np.allclose(h[n:], u[n:], rtol=0, atol=1e-14)
    ref = u.copy()
    O.cell_step(O.PCG2019, O.default_params(O.PCG2019), ref, n, 0.0, 0.01)
    assert np.allclose(h, ref, rtol=1e-14, atol=1e-15)
    ud.free()


def test_max_dphi_for_reaction_tangent_controller(tb, dev, oracle):
    """rtc.jl:64-67 reads maximum(du[:, phi]) (not abs) left behind by the sweep."""
    O = oracle
    prm = O.default_params(O.FHN)
    n = 5001
    u = _fhn_state(n, 5)
    ud = tb.B200Vector.from_host(dev, u, 2)
    mx = tb.core.cell_step(dev, 0, prm, ud, 0.0, 0.5, want_max=True)
    du = O.cell_step(O.FHN, prm, u, n, 0.0, 0.5)
    assert mx == du[:n].max()
    assert np.array_equal(ud.to_host(), u)
    ud.free()


def test_argument_validation(tb, dev, oracle):
    ud = tb.B200Vector(dev, 10, 2)
    with pytest.raises(tb.TBError):
        tb.core.cell_step(dev, 1, oracle.default_params(oracle.PCG2019), ud, 0.0, 0.1)     # needs 7 columns
    with pytest.raises(tb.TBError):
        tb.core.cell_step(dev, 0, np.zeros(5), ud, 0.0, 0.1)                                 # wrong parameter count
    with pytest.raises(tb.TBError):
        tb.core.cell_step(dev, 7, np.zeros(6), ud, 0.0, 0.1)                                 # unknown model
    ud.free()
