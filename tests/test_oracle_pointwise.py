"""The oracle against the known answers of the reference's own unit tests (no GPU, no reference
checkout needed): LatticeTests.cc:21-200, KernelTests.cc:114-143,296-370, BoundaryTests.cc:31-51,
and the spot values recorded from the compiled reference headers in SURVEY.md Appendix B."""
import math

import numpy as np
import pytest

import oracle as O

LATTICES = (15, 19, 27)


@pytest.fixture(scope="module", autouse=True)
def _built(oracle_built):
    return oracle_built


@pytest.mark.parametrize("Q", LATTICES)
def test_lattice_properties(Q):
    """LatticeTests.cc:21-105: vectors in {-1,0,1}, unique, inverse table, weights sum to 1."""
    c, w, inv = O.lattice(Q)
    assert set(np.unique(c)) <= {-1, 0, 1}
    assert len({tuple(v) for v in c}) == Q
    assert (c[0] == 0).all()
    for i in range(Q):
        assert (c[inv[i]] == -c[i]).all()
    assert list(inv[:5]) == [0, 2, 1, 4, 3]
    assert abs(w.sum() - 1.0) < 1e-15
    # second moment of the weights is Cs2 * identity
    m2 = np.einsum("i,ia,ib->ab", w, c, c)
    assert np.allclose(m2, np.eye(3) / 3.0, atol=1e-15)


@pytest.mark.parametrize("Q", LATTICES)
def test_density_momentum_feq_against_naive(Q):
    """KernelTests.cc:114-143 with LbTestsHelper::CalculateLBGKEqmF (f[i] = (i+1)/10)."""
    c, w, _ = O.lattice(Q)
    f = (np.arange(Q) + 1) / 10.0
    tau = 0.62
    r = O.collide(Q, "LBGK", tau, f)
    rho = f.sum()
    m = (c * f[:, None]).sum(0)
    assert r["rho"] == pytest.approx(rho, abs=1e-10)
    assert np.allclose(r["m"], m, atol=1e-10)
    assert np.allclose(r["u"], m / rho, atol=1e-10)
    mde = c @ m
    feq = w * (rho - 1.5 * (m @ m) / rho + 4.5 * mde * mde / rho + 3.0 * mde)
    assert np.allclose(r["feq"], feq, atol=1e-10)
    assert np.allclose(r["fpost"], f - (f - feq) / tau, atol=1e-10)
    if Q == 15:
        assert r["rho"] == pytest.approx(12.0, abs=1e-10)  # KernelTests.cc:126


def test_recorded_reference_values_d3q19():
    """values printed by the compiled reference headers (scalar path)."""
    f = (np.arange(19) + 1) / 10.0
    r = O.collide(19, "LBGK", 0.62, f)
    assert r["rho"] == 18.999999999999996
    assert tuple(r["m"]) == (-0.50000000000000022, -0.2999999999999996, -0.099999999999999867)
    assert r["feq"][7] == 0.46455409356725141
    assert r["fpost"][7] == 0.25895821543105058
    assert O.collide(19, "MRT", 0.62, f)["fpost"][7] == 0.62150759290699908


def test_mrt_basis_norms():
    """KernelTests.cc:296-370 golden BASIS_TIMES_BASIS_TRANSPOSED arrays."""
    n, r = np.zeros(15), np.zeros(15)
    L = O.oracle_lib()
    K = L.hlbo_mrt_basis(15, O.C.c_double(0.8), O._d(n), O._d(r))
    assert K == 11 and list(n[:11]) == [18, 360, 40, 40, 40, 12, 4, 8, 8, 8, 8]
    assert list(r[:11]) == [1.6, 1.2, 1.6, 1.6, 1.6, 1.25, 1.25, 1.25, 1.25, 1.25, 1.2]
    K = L.hlbo_mrt_basis(19, O.C.c_double(0.8), O._d(n), O._d(r))
    assert K == 15 and list(n) == [2394, 252, 40, 40, 40, 36, 72, 12, 24, 4, 4, 4, 8, 8, 8]
    assert list(r) == [1.19, 1.4, 1.2, 1.2, 1.2, 1.25, 1.4, 1.25, 1.4, 1.25, 1.25, 1.25, 1.98, 1.98, 1.98]


@pytest.mark.parametrize("Q", (15, 19))
def test_mrt_with_equal_rates_is_lbgk(Q):
    """KernelTests.cc:296-370: MRT with every relaxation rate = 1/tau behaves as LBGK."""
    tau = 0.62
    K = 11 if Q == 15 else 15
    f = (np.arange(Q) + 1) / 10.0
    out = np.zeros(Q)
    rates = np.full(K, 1.0 / tau)
    O.oracle_lib().hlbo_collide_mrt_rates(Q, O.C.c_double(tau), O._d(rates), O._d(f), O._d(out))
    assert np.allclose(out, O.collide(Q, "LBGK", tau, f)["fpost"], atol=1e-10)


@pytest.mark.parametrize("Q", LATTICES)
def test_trt_conserves_and_reduces_to_lbgk(Q):
    """TRT has no compilable reference (TRT.h:42-90); pin its derivation: mass and momentum are
    conserved and with tau_minus = tau_plus (Lambda = (tau-1/2)^2) it is LBGK."""
    c, w, _ = O.lattice(Q)
    rng = np.random.default_rng(3)
    f = w * (1 + 0.05 * rng.uniform(-1, 1, Q))
    r = O.collide(Q, "TRT", 0.8, f)
    assert abs(r["fpost"].sum() - f.sum()) < 1e-15
    assert np.abs((c * r["fpost"][:, None]).sum(0) - (c * f[:, None]).sum(0)).max() < 1e-15
    tau = 0.5 + math.sqrt(3.0 / 16.0)  # tau_minus == tau_plus
    a = O.collide(Q, "TRT", tau, f)["fpost"]
    b = O.collide(Q, "LBGK", tau, f)["fpost"]
    assert np.allclose(a, b, atol=1e-15)


@pytest.mark.parametrize("Q", LATTICES)
@pytest.mark.parametrize("tau", [0.51, 0.62, 0.8, 1.3])
def test_trt_against_the_published_two_relaxation_time_scheme(Q, tau):
    """An anchor outside the reference (whose TRT.h does not compile): Ginzburg's two-relaxation-time
    collision written independently -- split every population into its symmetric and antisymmetric
    parts over the pair (i, i-bar), relax them with omega+ = 1/tau and omega- fixed by the magic
    parameter Lambda = (1/omega+ - 1/2)(1/omega- - 1/2) -- must give the oracle's post-collision
    populations, and the oracle's antisymmetric rate must put Lambda at 3/16 for every tau
    (TRT.h:100-107), the value for which a half-way bounce-back wall sits exactly mid-link in
    Poiseuille flow whatever the viscosity."""
    c, w, inv = O.lattice(Q)
    rng = np.random.default_rng(11)
    f = w * (1 + 0.08 * rng.uniform(-1, 1, Q))
    r = O.collide(Q, "TRT", tau, f)
    feq = r["feq"]
    lam = 3.0 / 16.0
    om_p = 1.0 / tau
    om_m = 1.0 / (0.5 + lam / (tau - 0.5))
    fs, fa = 0.5 * (f + f[inv]), 0.5 * (f - f[inv])
    es, ea = 0.5 * (feq + feq[inv]), 0.5 * (feq - feq[inv])
    want = f - om_p * (fs - es) - om_m * (fa - ea)
    assert np.abs(r["fpost"] - want).max() <= 4e-16
    # the antisymmetric rate, measured: perturb one pair antisymmetrically around equilibrium
    i = int(np.nonzero(inv != np.arange(Q))[0][0])
    base = O.collide(Q, "TRT", tau, feq)["fpost"]          # equilibrium is a fixed point
    assert np.abs(base - feq).max() <= 4e-16
    g = feq.copy()
    eps = 1e-3 * feq[i]
    g[i] += eps
    g[inv[i]] -= eps
    out = O.collide(Q, "TRT", tau, g)
    # (mass is unchanged, momentum changes: measure the pair's antisymmetric part against its own f_eq)
    na_in = 0.5 * ((g[i] - out["feq"][i]) - (g[inv[i]] - out["feq"][inv[i]]))
    na_out = 0.5 * ((out["fpost"][i] - out["feq"][i]) - (out["fpost"][inv[i]] - out["feq"][inv[i]]))
    measured = 1.0 - na_out / na_in
    assert abs((tau - 0.5) * (1.0 / measured - 0.5) - lam) <= 1e-9


def test_tau_of_four_cube_xml():
    """dt = 0.0857 s, dx = 0.01 m => tau ~ 0.510284 (LbmParameters.h:35)."""
    tau = O.oracle_lib().hlbo_tau(O.C.c_double(0.0857), O.C.c_double(0.01), O.C.c_double(0.004), O.C.c_double(1000.0))
    assert tau == pytest.approx(0.510284, abs=1e-6)


def test_cosine_iolet_density():
    """BoundaryTests.cc:31-51 shape: mean - amp at t = 0 with phase pi, mean + amp half a period
    later, back after a full period."""
    L = O.oracle_lib()

    def rho(t):
        return L.hlbo_cosine_density(O.C.c_double(1.01), O.C.c_double(0.004), O.C.c_double(math.pi),
                                     O.C.c_double(100.0), O.C.c_double(0.0), O.C.c_double(1.0), O.C.c_uint64(t))
    assert rho(0) == pytest.approx(1.01 - 0.004, abs=1e-12)
    assert rho(50) == pytest.approx(1.01 + 0.004, abs=1e-12)
    assert rho(100) == pytest.approx(1.01 - 0.004, abs=1e-12)


def test_parabolic_velocity_profile():
    """InOutLetTests.cc:140-173 shape: v_max on the axis, zero at r = radius, along the normal."""
    L = O.oracle_lib()
    v = np.zeros(3)
    n = np.array([0.0, 0.0, 1.0])
    pos = np.array([5.0, 5.0, 1.0])
    L.hlbo_parabolic_velocity(O._d(n), O._d(pos), O.C.c_double(4.0), O.C.c_double(0.1), O.C.c_double(0.0),
                              O._d(np.array([5.0, 5.0, 1.5])), O.C.c_uint64(1), O._d(v))
    assert tuple(v) == (0.0, 0.0, 0.1)
    L.hlbo_parabolic_velocity(O._d(n), O._d(pos), O.C.c_double(4.0), O.C.c_double(0.1), O.C.c_double(0.0),
                              O._d(np.array([9.0, 5.0, 1.0])), O.C.c_uint64(1), O._d(v))
    assert abs(v[2]) < 1e-17
    L.hlbo_parabolic_velocity(O._d(n), O._d(pos), O.C.c_double(4.0), O.C.c_double(0.1), O.C.c_double(0.0),
                              O._d(np.array([7.0, 5.0, 1.0])), O.C.c_uint64(1), O._d(v))
    assert v[2] == pytest.approx(0.1 * 0.75, abs=1e-15)


@pytest.mark.parametrize("Q", LATTICES)
def test_stress_identities(Q):
    """LatticeTests.cc:229-289: the stress tensor of an equilibrium f_neq = 0 is the pressure on
    the diagonal; traction = sigma . n; tangential part is orthogonal to n."""
    s = O.stress_functions(Q, 1.03, 0.8, np.zeros(Q), np.array([0.0, 1.0, 0.0]))
    assert np.allclose(s["stress_tensor"].reshape(3, 3), np.eye(3) * 0.03 / 3.0, atol=1e-16)
    assert s["von_mises"] == 0 and s["shear_rate"] == 0
    rng = np.random.default_rng(1)
    fneq = 1e-3 * rng.uniform(-1, 1, Q)
    n = np.array([1.0, 2.0, -2.0]) / 3.0
    s = O.stress_functions(Q, 1.01, 0.9, fneq, n)
    sig = s["stress_tensor"].reshape(3, 3)
    assert np.allclose(sig, sig.T)
    assert np.allclose(s["traction"], sig @ n, atol=1e-16)
    assert abs(s["tangential_traction"] @ n) < 1e-16


def test_one_site_bfl_known_answer():
    """one-site D3Q19 LBGK BFL run of the compiled reference: all 18 links wall,
    q = 0.8 for d = 1,2 and 0.3 otherwise, tau = 0.62, f_old[i] = (i+1)/10."""
    from hemelb_b200.geometry import Geometry, NEIGHBOURHOOD
    from hemelb_b200.domain import gmy_link_of_direction
    bt = np.ones((1, 26), np.uint8)
    bd = np.full((1, 26), 0.3, np.float32)
    lk = gmy_link_of_direction(19)
    bd[0, lk[1]] = 0.8
    bd[0, lk[2]] = 0.8
    g = Geometry(np.array([1, 1, 1], np.int32), 4, np.array([[1, 1, 1]], np.int32), np.array([0], np.int64), bt,
                 np.full((1, 26), -1, np.int32), bd, np.array([1], np.uint8), np.array([[0, 0, 1]], np.float32))
    dom = O.OracleDomains(g, 19)
    sim = O.OracleSim(dom, "LBGK", "BFL", tau=0.62)
    f = np.zeros(20)
    f[:19] = (np.arange(19) + 1) / 10.0
    sim.set_f(f)
    sim.stream_and_collide(1, 0, 1)
    sim.post_step(1, 0, 1)
    out = sim.get_f(which=1)
    assert out[1] == 1.6558762497641955
    assert out[2] == 1.4483493680437654
    assert out[7] == 0.41272165629126545


def _duct_profile(kernel, tau, W=6, L=24, drho=1e-5, wall="SBB"):
    """Steady pressure-driven flow along a lattice-aligned square duct with half-way bounce-back walls:
    u_z over the middle cross-section times nu / (-dp/dz), i.e. in units that do not depend on the viscosity."""
    from hemelb_b200.capi import iolet_record
    from hemelb_b200.lbm import prepare_boundary_objects
    from tests.cases import square_duct
    geom = square_duct(W, L)
    inl, outl = geom.meta["inlets"][0], geom.meta["outlets"][0]
    ins = [iolet_record(0, tuple(inl.normal), tuple(inl.position), radius=W, density_mean=1 + drho / 2)]
    outs = [iolet_record(0, tuple(outl.normal), tuple(outl.position), radius=W, density_mean=1 - drho / 2)]
    prepare_boundary_objects(ins, outs)
    dom = O.OracleDomains(geom, 19)
    sim = O.OracleSim(dom, kernel, wall, "NASH", "NASH", tau=tau, inlets=ins, outlets=outs)
    sim.set_equilibrium(1.0)
    nu = (tau - 0.5) / 3.0
    sim.step(int(12 * W * W / nu) + 2000)  # a dozen viscous diffusion times across the duct
    sim.set_cache_mask(3)
    sim.step(1)
    c = dom.tables(0)["globalCoords"].reshape(-1, 3)
    u, rho = sim.get_cache("velocity").reshape(-1, 3), sim.get_cache("density")
    zs = np.arange(1 + L // 4, 1 + 3 * L // 4)
    gradient = np.polyfit(zs, [rho[c[:, 2] == z].mean() for z in zs], 1)[0]
    mid = c[:, 2] == 1 + L // 2
    order = np.lexsort((c[mid, 1], c[mid, 0]))
    return u[mid, 2][order] * nu / (-gradient / 3.0)


def test_trt_wall_location_does_not_depend_on_viscosity():
    """The defining property of the two-relaxation-time collision with Lambda = 3/16 (the value of
    TRT.h:100-107), and an anchor for the whole TRT path -- collision, streaming, bounce-back -- that owes
    nothing to the reference (whose TRT.h does not compile): with half-way bounce-back walls the steady
    duct flow, in units of (-dp/dz) / nu, is the same whatever the viscosity, because the wall sits
    mid-link for every tau.  With the single-relaxation-time collision the apparent wall moves with tau.
    Measured: TRT 9e-6 between tau = 0.6 and 1.2, LBGK 8e-2; both against the Fourier-series solution of
    the square duct at the node positions."""
    W = 6
    trt = [_duct_profile("TRT", tau) for tau in (0.7, 1.4)]
    bgk = [_duct_profile("LBGK", tau) for tau in (0.7, 1.4)]
    assert np.abs(trt[0] - trt[1]).max() / trt[0].max() < 5e-5
    assert np.abs(bgk[0] - bgk[1]).max() / bgk[0].max() > 2e-2
    a = W / 2.0
    xs = np.arange(W) - (W - 1) / 2.0
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    series = sum((-1) ** ((n - 1) // 2) / n ** 3 * (1 - np.cosh(n * np.pi * Y / (2 * a)) / np.cosh(n * np.pi / 2))
                 * np.cos(n * np.pi * X / (2 * a)) for n in range(1, 200, 2))
    exact = (16 * a * a / np.pi ** 3 * series).ravel()
    for p in trt:
        assert np.abs(p - exact).max() / exact.max() < 0.012  # (second-order scheme on six nodes across)


def test_mrt_with_guo_zheng_shi_walls_gives_the_duct_flow():
    """MRT + GuoZhengShi has no reference anchor either (GuoZhengShi.h:269-282 collides a HydroVars whose m_neq
    was never set; the oracle projects the wall node's f_neq into moment space first).
    What the oracle's reading gives must at least be the flow: steady duct flow within 3 % of the
    Fourier-series solution on six nodes (measured 2.2 %; LBGK + GZS 1.4 %, MRT + BFL 3.1 %) and within
    1.5 % of LBGK with the same walls."""
    W, L = 6, 16
    mrt = _duct_profile("MRT", 0.8, W, L, wall="GZS")
    bgk = _duct_profile("LBGK", 0.8, W, L, wall="GZS")
    a = W / 2.0
    xs = np.arange(W) - (W - 1) / 2.0
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    series = sum((-1) ** ((n - 1) // 2) / n ** 3 * (1 - np.cosh(n * np.pi * Y / (2 * a)) / np.cosh(n * np.pi / 2))
                 * np.cos(n * np.pi * X / (2 * a)) for n in range(1, 200, 2))
    exact = (16 * a * a / np.pi ** 3 * series).ravel()
    assert np.isfinite(mrt).all()
    assert np.abs(mrt - exact).max() / exact.max() < 0.03
    assert np.abs(mrt - bgk).max() / exact.max() < 0.015
