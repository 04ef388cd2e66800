"""Golden vectors recorded from the compiled reference headers (tests/golden/make_golden.py):
the oracle must reproduce them bit-for-bit wherever it runs (no reference checkout needed), and
on a GPU the CUDA path must too."""
import os

import numpy as np
import pytest

import oracle as O
from hemelb_b200 import geometry as G
from tests.cases import anisotropic_f, geometry, iolets_for

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# ref_vectors.npz: unmodified reference code.  ref_vectors_trt.npz: the reference's streamers around the reference's
# TRT::Collide, whose bit-rotted header is compiled through three build-time substitutions (make_golden_trt.py);
# ref_vectors_mrt_patched.npz: MRT + GuoZhengShi / MRT + Nash from the reference with its one missing line and its
# stale CalculateFeq repaired at build time (same script).
GOLD = {}
for _name in ("ref_vectors.npz", "ref_vectors_trt.npz", "ref_vectors_mrt_patched.npz"):
    if os.path.exists(os.path.join(HERE, _name)):
        with np.load(os.path.join(HERE, _name)) as _z:
            GOLD.update({k: _z[k] for k in _z.files})
KEYS = sorted(k[:-4] for k in GOLD if k.endswith("_tau"))
STEPS = 5


def _parse(key):
    gname, R, Q, k, w, i, o = key.rsplit("_", 6)
    return gname, int(R[1:]), int(Q[1:]), k, w, i, o


@pytest.mark.parametrize("key", KEYS)
def test_oracle_reproduces_reference_vectors(key):
    gname, R, Q, k, w, i, o = _parse(key)
    geom = geometry(gname)
    rank = None if R == 1 else G.slab_decomposition(geom, R)
    inlets, outlets = iolets_for(geom, i, o)
    dom = O.OracleDomains(geom, Q, rank, R)
    sim = O.OracleSim(dom, k, w, i, o, tau=float(GOLD[key + "_tau"][0]), inlets=inlets, outlets=outlets)
    for r in range(R):
        t = dom.tables(r)
        sim.set_f(anisotropic_f(t["N"], Q, t["totalSharedFs"], site_offset=3 * r), r)
    sim.set_cache_mask(3)
    sim.step(STEPS)
    for r in range(R):
        n = dom.tables(r)["N"] * Q
        assert np.array_equal(sim.get_f(r)[:n], GOLD["%s_f%d" % (key, r)])
        assert np.array_equal(sim.get_cache("density", r), GOLD["%s_rho%d" % (key, r)])
        assert np.array_equal(sim.get_cache("velocity", r), GOLD["%s_u%d" % (key, r)])


@pytest.mark.gpu
# (the four-cube bundles here; the larger single-rank goldens -- configs[3]'s and configs[4]'s bundles on the tree and
# the sac -- are held against the GPU in tests/test_zzgpu_vs_ref_lbm.py)
@pytest.mark.parametrize("key", [k for k in KEYS if k.startswith("four_cube_R1_")])
def test_gpu_reproduces_reference_vectors(key):
    from hemelb_b200.domain import build_domains
    from hemelb_b200.lbm import GpuLBM
    gname, R, Q, k, w, i, o = _parse(key)
    geom = geometry(gname)
    inlets, outlets = iolets_for(geom, i, o)
    dom = build_domains(geom, Q)[0]
    gpu = GpuLBM(dom, k, w, i, o, tau=float(GOLD[key + "_tau"][0]), inlets=inlets, outlets=outlets)
    gpu.set_f(anisotropic_f(dom.N, Q, 0))
    gpu.set_cache_mask(3)
    gpu.step(STEPS)
    f = gpu.get_f()[:dom.N * Q]
    assert np.abs(f - GOLD[key + "_f0"]).max() <= 1e-13
    assert np.array_equal(f, GOLD[key + "_f0"])
    assert np.array_equal(gpu.get_cache("density"), GOLD[key + "_rho0"])
    assert np.array_equal(gpu.get_cache("velocity"), GOLD[key + "_u0"])
