"""A/B of the two CPU forms of the reference on this machine's cores (no GPU needed; needs oracle/_ref):

  arm   the timed `bench.py --impl reference` arm: the reference's streamers driven by oracle/ref_driver.cc's
        href_sim_step_mt (one thread per emulated rank, neighbour flags, one barrier per step)
  lbm   the reference's whole lb::LBM over its own net::Net and StepManager (oracle/ref_lbm_driver.cc)

on a cylinder (one inlet, one outlet: the reference's local-iolet lookup is safe on any decomposition) over
BasicDecomposition, scalar and SSE3 builds, interleaved repetitions.  Output: profiles/r02_reference_arm_ab.txt.
`--json [--sse3-only] [--reps N]`: one JSON object on the last line (bench.py --impl reference runs it that way, in a
child process, and puts the result into its line as `cross_check`)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle as O  # noqa: E402
from bench import TAU, pressure_iolets, usable_cores  # noqa: E402
from hemelb_b200 import geometry as G  # noqa: E402
from hemelb_b200.domain import build_domains  # noqa: E402


def main():
    Q, steps, reps = 19, 20, 3
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("radius", nargs="?", type=float, default=40.0)
    ap.add_argument("length", nargs="?", type=int, default=200)
    ap.add_argument("--json", action="store_true")
    ap.add_argument("--sse3-only", action="store_true")
    ap.add_argument("--reps", type=int, default=reps)
    args = ap.parse_args()
    as_json, sse3_only, reps, radius, length = args.json, args.sse3_only, args.reps, args.radius, args.length
    builds = (True,) if sse3_only else (False, True)
    result = {"workload": "cylinder r=%g l=%d, D3Q19 LBGK+BFL+Nash, BasicDecomposition" % (radius, length), "steps": steps}
    geom = G.cylinder(radius, length)
    R = usable_cores()
    rank_of = G.basic_decomposition(geom, R) if R > 1 else None
    doms = build_domains(geom, Q, rank_of, R)
    tables = [d.tables() for d in doms]
    n = sum(d.N for d in doms)
    inlets, outlets = pressure_iolets(geom.meta)
    dt = (TAU - 0.5) / 3.0 * 1000.0 / 0.004
    w = O.lattice(Q)[1]
    f0 = []
    for t in tables:
        f = np.zeros(t["N"] * Q + 1 + t["totalSharedFs"])
        f[:t["N"] * Q] = np.tile(w, t["N"])
        f0.append(f)
    sims = {}
    for sse3 in builds:
        sim = O.RefSim(tables, Q, "LBGK", "BFL", "NASH", "NASH", dt=dt, dx=1.0, rho=1000.0, eta=0.004, inlets=inlets,
                       outlets=outlets, sse3=sse3)
        for r, f in enumerate(f0):
            sim.set_f(f, r)
            sim.set_f(f, r, 1)
        sim.step_mt(2)
        sims[sse3] = sim
    print("cylinder r=%g l=%d: %d sites, %d emulated ranks = threads, %d steps per measurement" % (radius, length, n, R, steps))
    result.update(sites=int(n), threads=int(R))
    for rep in range(reps):
        for sse3 in builds:
            t0 = time.perf_counter()
            sims[sse3].step_mt(steps)
            arm = n * steps / (time.perf_counter() - t0) / 1e6
            tm = []
            O.ref_lbm_run(geom, Q, "BFL", "NASH", inlets, outlets, dt, 1.0, steps, [d.N for d in doms], rank_of, R, f0=f0,
                          sse3=sse3, timing=tm)
            tag = "sse3" if sse3 else "scalar"
            result.setdefault("arm_mlups_" + tag, []).append(round(arm, 2))
            result.setdefault("reference_lbm_mlups_" + tag, []).append(round(n * steps / tm[0] / 1e6, 2))
            print("rep %d  %-6s  arm %6.1f MLUPS   lbm %6.1f MLUPS   arm / lbm %.2f" % (
                rep, "SSE3" if sse3 else "scalar", arm, n * steps / tm[0] / 1e6, arm / (n * steps / tm[0] / 1e6)), flush=True)
    if as_json:
        print(json.dumps(result), flush=True)


if __name__ == "__main__":
    main()
