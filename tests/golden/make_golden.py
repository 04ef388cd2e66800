"""Generate tests/golden/ref_vectors.npz from the compiled reference (oracle/_ref).

Run HERE (where /root/reference exists): `python tests/golden/make_golden.py`.  Every case is the
UNMODIFIED reference lattice / kernel / streamer code (oracle/ref_driver.cc) stepping a small
geometry from LbTestsHelper's anisotropic initial data; the stored outputs pin the oracle -- and
through it the CUDA path -- on machines where the reference sources are absent (the GPU box).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle as O  # noqa: E402
from hemelb_b200 import geometry as G  # noqa: E402
from tests.cases import anisotropic_f, geometry, iolets_for, valid_combo  # noqa: E402

DT, DX, RHO, ETA = 1e-4, 1e-4, 1000.0, 0.004
STEPS = 5


def cases():
    for Q in (15, 19, 27):
        for k in ("LBGK", "MRT"):
            for w in ("SBB", "BFL", "GZS"):
                for (i, o) in (("NASH", "NASH"), ("LADD", "NASH"), ("LADD", "LADD")):
                    if valid_combo(Q, k, w, i, o, need_ref=True):
                        yield ("four_cube", 1, Q, k, w, i, o)
    for (Q, k, w, i, o) in [(19, "LBGK", "BFL", "NASH", "NASH"), (19, "LBGK", "GZS", "LADD", "NASH"),
                            (15, "MRT", "BFL", "LADD", "LADD"), (27, "LBGK", "SBB", "NASH", "NASH")]:
        yield ("cylinder", 2, Q, k, w, i, o)


def main():
    O.build()
    assert O.ref_lib() is not None, "oracle/_ref not built (needs /root/reference)"
    out = {}
    for (gname, R, Q, k, w, i, o) in cases():
        geom = geometry(gname)
        rank = None if R == 1 else G.slab_decomposition(geom, R)
        inlets, outlets = iolets_for(geom, i, o)
        dom = O.OracleDomains(geom, Q, rank, R)
        T = [dom.tables(r) for r in range(R)]
        ref = O.RefSim(T, Q, k, w, i, o, dt=DT, dx=DX, rho=RHO, eta=ETA, inlets=inlets, outlets=outlets)
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
    path = os.path.join(ROOT, "tests", "golden", "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
