#!/usr/bin/env python
"""End to end through the C++ policy classes: tests/_build/host_lbm_run (hemelb_b200/host's Gpu streamers
and device-backed FieldData driven in lb::LBM's phase order, with an lb::StabilityTester assessing every
step) timed wall-clock by its own HLB_HOST_TIMING, next to the same steps driven from Python over
ctypes and to the device-resident figure, on one geometry.

  python bench_host_cxx.py [--radius 100 --length 200] [--steps 200]

The harness reads its Domain from a case file (N*Q int64 neighbour indices and so on), which bounds the
size: the default cylinder has 6.3e6 sites (a 2 GB file under $TMPDIR).  A secondary measurement; the
contract line is bench.py.  One JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import sysconfig
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--radius", type=float, default=100.0)
    ap.add_argument("--length", type=int, default=200)
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    from hemelb_b200 import geometry as G
    from hemelb_b200.capi import iolet_record
    from hemelb_b200.domain import build_domains
    from hemelb_b200.lbm import GpuLBM, prepare_boundary_objects
    from tests.host_build import BUILD
    from tests.test_host_lbm import physical_dt, write_case
    exe = os.path.join(BUILD, "host_lbm_run")
    if not os.path.exists(exe):
        print(json.dumps({"unavailable": "tests/_build/host_lbm_run was not prebuilt (needs the reference headers)"}))
        return 0
    Q, tau = 19, 0.8
    t0 = time.time()
    geom = G.cylinder_extruded(args.radius, args.length)
    dom = build_domains(geom, Q)[0]
    inl, outl = geom.meta["inlets"][0], geom.meta["outlets"][0]
    inlets = [iolet_record(0, tuple(inl.normal), tuple(inl.position), radius=args.radius, density_mean=1.0005, period=1000.0)]
    outlets = [iolet_record(0, tuple(outl.normal), tuple(outl.position), radius=args.radius, density_mean=0.9995, period=1000.0)]
    prepare_boundary_objects(inlets, outlets)
    import oracle as O
    _, w, _ = O.lattice(Q)
    f0 = np.zeros(dom.N * Q + 1)
    f0[:dom.N * Q] = np.tile(w, dom.N)
    setup = time.time() - t0
    line = {"geometry": "cylinder r=%g l=%d" % (args.radius, args.length), "sites": int(dom.N), "steps": args.steps,
            "policies": "D3Q19 LBGK + BFL + Nash", "setup_seconds": setup}
    with tempfile.TemporaryDirectory() as tmp:
        case = os.path.join(tmp, "case.bin")
        write_case(case, dom, "LBGK", "BFL", "NASH", "NASH", inlets, outlets, f0, args.steps + 1, 0, physical_dt(tau))
        extra = ["/usr/local/cuda/lib64", os.path.join(sysconfig.get_paths()["purelib"], "nvidia", "cuda_runtime", "lib")]
        for label, stab in (("cxx", None), ("cxx_with_stability_tester", "1")):
            env = dict(os.environ, LD_LIBRARY_PATH=":".join([os.environ.get("LD_LIBRARY_PATH", "")] + extra).strip(":"),
                       HLB_HOST_TIMING="1")
            if stab:
                env["HLB_HOST_STABILITY"] = stab
            r = subprocess.run([exe, case, os.path.join(tmp, "out.bin")], capture_output=True, text=True, env=env, timeout=1800)
            m = re.search(r"timed steps, ([0-9.]+) s, ([0-9.]+) MLUPS", r.stderr)
            line[label + "_mlups"] = float(m.group(2)) if m else None
            if not m:
                line[label + "_error"] = r.stderr[-300:]
    gpu = GpuLBM(dom, "LBGK", "BFL", "NASH", "NASH", tau=tau, inlets=inlets, outlets=outlets)
    gpu.set_f(f0)
    gpu.step(10)
    ms = gpu.time_steps(args.steps)
    line["device_resident_mlups"] = dom.N * args.steps / (ms * 1e-3) / 1e6
    gpu.set_cache_mask(256)
    gpu.do_time_step()
    gpu.monitor()
    gpu.sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        gpu.do_time_step()
        gpu.monitor()
    gpu.sync()
    line["python_ctypes_mlups"] = dom.N * args.steps / (time.perf_counter() - t0) / 1e6
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
