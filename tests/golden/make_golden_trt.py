"""Generate tests/golden/ref_vectors_trt.npz: TRT bundles stepped by the reference's own streamers around the
reference's own TRT::Collide.

Run HERE (where /root/reference exists): `python tests/golden/make_golden_trt.py`.  lb/kernels/TRT.h has bit-rotted
in the reference; oracle/Makefile compiles it through three substitutions made at build time (`iBar >= i` ->
`iBar > i`, `f_neq.f[` -> `f_neq[`, `f_eq.f[` -> `f_eq[`) and oracle/ref_driver.cc wraps its Collide in a kernel
whose f_eq / f_neq are computed as LBGK.h computes them (TRT.h:62-92 states the same) -- see DESIGN.md section 2.
Kept apart from ref_vectors.npz, which holds unmodified reference code only.

Also writes tests/golden/ref_vectors_mrt_patched.npz from oracle/_ref/libhemelb_ref_mrtgzs.so: MRT + GuoZhengShi with
the one missing m_neq projection inserted (GuoZhengShi.h:279 is undefined behaviour as written) and MRT + Nash iolets
with MRT::CalculateFeq brought to the form of its neighbour (MRT.h:73-86 does not compile) -- configs[3]'s bundle.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle as O  # noqa: E402
from hemelb_b200 import geometry as G  # noqa: E402
from tests.cases import anisotropic_f, geometry, iolets_for  # noqa: E402

DT, DX, RHO, ETA = 1e-4, 1e-4, 1000.0, 0.004
STEPS = 5


def cases():
    for Q in (15, 19, 27):
        for w in ("SBB", "BFL", "GZS"):
            for (i, o) in (("NASH", "NASH"), ("LADD", "NASH"), ("LADD", "LADD")):
                yield ("four_cube", 1, Q, "TRT", w, i, o)
    yield ("sac", 1, 27, "TRT", "BFL", "NASH", "NASH")       # configs[4]'s bundle
    yield ("sac", 2, 27, "TRT", "BFL", "NASH", "NASH")
    yield ("cylinder", 2, 19, "TRT", "GZS", "LADD", "NASH")


def mrt_cases():
    yield ("four_cube", 1, 19, "MRT", "GZS", "LADD", "NASH")
    yield ("four_cube", 1, 19, "MRT", "GZS", "LADD", "LADD")
    yield ("four_cube", 1, 15, "MRT", "GZS", "NASH", "NASH")
    yield ("four_cube", 1, 15, "MRT", "BFL", "NASH", "NASH")
    yield ("four_cube", 1, 19, "MRT", "SBB", "LADD", "NASH")
    yield ("tree", 1, 19, "MRT", "GZS", "LADD", "NASH")       # configs[3]'s bundle
    yield ("tree", 2, 19, "MRT", "GZS", "LADD", "NASH")


def main():
    O.build()
    assert O.ref_lib() is not None, "oracle/_ref not built (needs /root/reference)"
    write(cases(), False, "ref_vectors_trt.npz")
    assert O.ref_lib("mrtgzs") is not None
    write(mrt_cases(), "mrtgzs", "ref_vectors_mrt_patched.npz")


def write(case_list, lib, filename):
    out = {}
    for (gname, R, Q, k, w, i, o) in case_list:
        geom = geometry(gname)
        rank = None if R == 1 else G.slab_decomposition(geom, R)
        inlets, outlets = iolets_for(geom, i, o)
        dom = O.OracleDomains(geom, Q, rank, R)
        T = [dom.tables(r) for r in range(R)]
        ref = O.RefSim(T, Q, k, w, i, o, dt=DT, dx=DX, rho=RHO, eta=ETA, inlets=inlets, outlets=outlets, sse3=lib)
        for r in range(R):
            ref.set_f(anisotropic_f(T[r]["N"], Q, T[r]["totalSharedFs"], site_offset=3 * r), r)
        ref.set_cache_mask(3)
        ref.step(STEPS)
        key = "%s_R%d_Q%d_%s_%s_%s_%s" % (gname, R, Q, k, w, i, o)
        out[key + "_tau"] = np.array([ref.tau])
        for r in range(R):
            n = T[r]["N"] * Q
            out["%s_f%d" % (key, r)] = ref.get_f(r)[:n]
            out["%s_rho%d" % (key, r)] = ref.get_cache("density", r)
            out["%s_u%d" % (key, r)] = ref.get_cache("velocity", r)
    path = os.path.join(ROOT, "tests", "golden", filename)
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
