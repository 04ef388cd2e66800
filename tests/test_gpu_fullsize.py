"""Checks the oracle cannot reach: the analytic Poiseuille profile (configs[1] "checked against the
analytic profile") and size-independent properties at BASELINE.json's full benchmark size
(1.0e8 sites, the geometry bench.py times): the rest state is a fixed point, and the result does
not depend on the internal site numbering or on the kernel schedule -- three engines with different
device layouts / launch orders must give bit-identical density and velocity fields."""
import numpy as np
import pytest

from hemelb_b200.capi import iolet_record
from hemelb_b200.lbm import prepare_boundary_objects

pytestmark = pytest.mark.gpu
TAU = 0.8


def _cylinder_engine(radius, length, wall="BFL", reorder=True, drho=1e-3):
    from hemelb_b200.devdomain import DeviceDomain, cylinder_shape
    from hemelb_b200.lbm import GpuLBM
    caps, iolets, shape = cylinder_shape(radius, length)
    dom = DeviceDomain.from_shape(caps, iolets, shape, 19, 8, None, 0, 1, 0)
    inl, outl = dom.meta["inlets"][0], dom.meta["outlets"][0]
    ins = [iolet_record(0, tuple(inl.normal), tuple(inl.position), radius=radius, density_mean=1.0 + drho / 2)]
    outs = [iolet_record(0, tuple(outl.normal), tuple(outl.position), radius=radius, density_mean=1.0 - drho / 2)]
    prepare_boundary_objects(ins, outs)
    gpu = GpuLBM.from_device_domain(dom, "LBGK", wall, "NASH", "NASH", tau=TAU, inlets=ins, outlets=outs, reorder=reorder)
    gpu.set_equilibrium(1.0, (0.0, 0.0, 0.0))
    return dom, gpu


@pytest.mark.parametrize("wall,tol", [("BFL", 0.01), ("SBB", 0.06)])
def test_poiseuille_profile(wall, tol):
    """Steady pressure-driven flow in a straight cylinder (R = 24.3, L = 192 voxels, ~3.6e5 sites):
    u_z(r) = -(dp/dz) (R^2 - r^2) / (4 rho nu), nu = cs^2 (tau - 1/2), with dp/dz = cs^2 d(rho)/dz
    taken from the simulated density field in the middle half of the pipe (the Nash iolets place
    their pressure planes ~1.4 sites inside the ends: the nominal gradient is 4-5 % off).
    Calibrated with the CPU oracle at R = 10.3: BFL 0.4 %, simple bounce-back 4.5 % (staircase)."""
    R, L, steps = 24.3, 192, 14000
    dom, gpu = _cylinder_engine(R, L, wall)
    gpu.step(steps)
    gpu.set_cache_mask(3)
    gpu.step(1)
    mon = gpu.monitor()
    assert mon["min_f"] > 0
    c = dom.global_coords()
    u = gpu.get_cache("velocity").reshape(-1, 3)
    rho = gpu.get_cache("density")
    n = int(np.ceil(2 * R)) + 2 * 2 + 1
    cx = cy = (n - 1) / 2.0
    z = c[:, 2] - 2
    # linear density drop along the pipe
    zs = np.arange(L // 4, 3 * L // 4)
    counts = np.bincount(z, minlength=L)
    rz = (np.bincount(z, weights=rho, minlength=L) / counts)[zs]
    g, c0 = np.polyfit(zs, rz, 1)
    assert np.abs(np.polyval([g, c0], zs) - rz).max() < 1e-5 * 1e-3
    assert abs(g / (-1e-3 / L) - 1.0) < 0.07
    mid = z == L // 2
    r = np.hypot(c[mid, 0] - cx, c[mid, 1] - cy)
    ua = (-g) * (R * R - r * r) / (4.0 * rho[mid].mean() * (TAU - 0.5))
    err = np.abs(u[mid, 2] - ua) / ua.max()
    assert err.max() < tol, err.max()
    # no swirl, no radial flow
    assert np.abs(u[mid, :2]).max() < 2e-3 * ua.max()
    if wall == "BFL":  # the fitted parabola vanishes at the true wall radius
        A = np.stack([np.ones_like(r), -r * r], 1)
        (a, b), *_ = np.linalg.lstsq(A, u[mid, 2], rcond=None)
        assert abs(np.sqrt(a / b) / R - 1.0) < 0.004
    gpu.close()
    dom.close()


def _fields(gpu, steps):
    gpu.step(steps)
    gpu.set_cache_mask(3)
    gpu.step(1)
    gpu.set_cache_mask(0)
    return gpu.get_cache("density"), gpu.get_cache("velocity"), gpu.monitor()


def test_full_size_rest_state_and_layout_independence():
    """bench.py's workload (cylinder r = 146, l = 1500: 1.004e8 sites, D3Q19 LBGK + BFL)."""
    R, L = 146.0, 1500
    # (1) equal iolet densities, fluid at rest: nothing may move
    dom, gpu = _cylinder_engine(R, L, drho=0.0)
    assert dom.N == 100417500
    gpu.set_cache_mask(256)
    gpu.step(40)
    mon = gpu.monitor()
    assert abs(mon["min_density"] - 1.0) < 1e-13 and abs(mon["max_density"] - 1.0) < 1e-13
    assert mon["max_speed"] < 1e-12
    gpu.close()
    dom.close()
    # (2) the driven flow: renumbered + overlapped (default), renumbered + serial, reference order
    ref = None
    for reorder, overlap in ((True, True), (True, False), (False, True)):
        dom, gpu = _cylinder_engine(R, L, reorder=reorder)
        gpu.set_overlap(overlap)
        rho, u, mon = _fields(gpu, 24)
        gpu.close()
        dom.close()
        assert 0.999 < mon["min_density"] <= mon["max_density"] < 1.001 and mon["min_f"] > 0
        if ref is None:
            ref = (rho, u)
            assert float(np.abs(u).max()) > 0  # the pressure wave has started the fluid
        else:
            assert np.array_equal(rho, ref[0]), (reorder, overlap)
            assert np.array_equal(u, ref[1]), (reorder, overlap)
