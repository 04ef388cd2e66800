#!/usr/bin/env python
"""Secondary measurements: the other BASELINE.json configs at sizes the numpy voxeliser reaches in
a minute or two (the headline metric and contract line live in bench.py).

  python bench_configs.py [--steps 50] [--only tree_lbgk_bfl,...]

One JSON line per config: MLUPS on one GPU, the roofline fraction of the whole step and of the
mid-fluid kernel at B(Q) = 20*Q bytes per site update, the boundary-site fraction, and the extra
boundary bytes that B(Q) deliberately leaves out.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from hemelb_b200 import geometry as G  # noqa: E402
from hemelb_b200.capi import iolet_record  # noqa: E402
from hemelb_b200.domain import build_domains  # noqa: E402
from hemelb_b200.lbm import GpuLBM, prepare_boundary_objects  # noqa: E402


def iolets(geom, inlet_bc, outlet_bc):
    def rec(p, bc, k, inlet):
        if bc == "LADD":
            return iolet_record(1, tuple(p.normal), tuple(p.position), radius=p.radius - 1.0, max_speed=0.01)
        return iolet_record(0, tuple(p.normal), tuple(p.position), radius=p.radius,
                            density_mean=1.0005 if inlet else 0.9995, density_amp=0.0, period=1000.0)
    ins = [rec(p, inlet_bc, k, True) for k, p in enumerate(geom.meta["inlets"])]
    outs = [rec(p, outlet_bc, k, False) for k, p in enumerate(geom.meta["outlets"])]
    prepare_boundary_objects(ins, outs)
    return ins, outs


CONFIGS = {
    # name: (geometry factory, Q, kernel, wall, inlet, outlet, which BASELINE config it stands for)
    "cylinder_lbgk_bfl": (lambda: G.cylinder_extruded(146, 300), 19, "LBGK", "BFL", "NASH", "NASH", "configs[1] at 1/5 length"),
    "tree_lbgk_bfl": (lambda: G.capsule_tree(5, 36.0, 150.0), 19, "LBGK", "BFL", "NASH", "NASH", "configs[2], 5 generations"),
    "tree_mrt_gzs_ladd": (lambda: G.capsule_tree(5, 36.0, 150.0), 19, "MRT", "GZS", "LADD", "NASH", "configs[3]"),
    "sac_trt_bfl_q27": (lambda: G.sac(110, 30, 40, roughness=3.0), 27, "TRT", "BFL", "NASH", "NASH", "configs[4]"),
    "cylinder_lbgk_sbb_q15": (lambda: G.cylinder_extruded(146, 300), 15, "LBGK", "SBB", "NASH", "NASH", "configs[0] policies at scale"),
    "cylinder_mrt_bfl_q19": (lambda: G.cylinder_extruded(146, 300), 19, "MRT", "BFL", "NASH", "NASH", "MRT bulk cost"),
    "cylinder_lbgk_bfl_q27": (lambda: G.cylinder_extruded(146, 300), 27, "LBGK", "BFL", "NASH", "NASH", "D3Q27 bulk"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    names = [n for n in CONFIGS if not args.only or n in args.only.split(",")]
    geoms = {}
    for name in names:
        factory, Q, kernel, wall, inlet, outlet, what = CONFIGS[name]
        key = factory.__code__.co_code + repr(factory.__code__.co_consts).encode()
        t0 = time.time()
        if key not in geoms:
            geoms[key] = factory()
        geom = geoms[key]
        dom = build_domains(geom, Q)[0]
        ins, outs = iolets(geom, inlet, outlet)
        gpu = GpuLBM(dom, kernel, wall, inlet, outlet, tau=0.8, inlets=ins, outlets=outs)
        gpu.set_equilibrium(1.0, (0.0, 0.0, 0.0))
        setup = time.time() - t0
        gpu.step(args.warmup)
        gpu.sync()
        ms, fused_ms, fused_sites = gpu.time_steps_detail(args.steps)  # product schedule: fused mid-domain kernel
        gpu.sync()
        gpu.set_overlap(False)           # second region, kernels back to back: the bulk kernel's own duration
        gpu.step(2)
        gpu.sync()
        serial_ms, bulk_ms, bulk_sites = gpu.time_steps_detail(args.steps)
        gpu.set_overlap(True)
        mon = gpu.monitor()
        B = 20 * Q
        mlups = dom.N * args.steps / (ms * 1e-3) / 1e6
        nb = int(dom.N - dom.mid[0] - dom.edge[0])
        wall_links = int(sum(bin(int(x)).count("1") for x in dom.wallMask[dom.wallMask != 0]))
        line = {"config": name, "stands_for": what, "lattice": "D3Q%d" % Q, "kernel": kernel, "wall": wall,
                "inlet": inlet, "outlet": outlet, "sites": dom.N, "boundary_typed_sites": nb,
                "boundary_fraction": nb / dom.N, "n_inlets": len(ins), "n_outlets": len(outs),
                "MLUPS": mlups, "ms_per_step": ms / args.steps, "bytes_per_site": B,
                "whole_step_frac_of_hbm_roofline": mlups * 1e6 * B / 1e9 / peak,
                "bulk_kernel_GBps": fused_sites * B / 1e9 / (fused_ms * 1e-3) if fused_ms else None,
                "bulk_kernel_frac": (fused_sites * B / 1e9 / (fused_ms * 1e-3)) / peak if fused_ms else None,
                "bulk_kernel_frac_plain_order": (bulk_sites * B / 1e9 / (bulk_ms * 1e-3)) / peak if bulk_ms else None,
                "fused_mid_kernel": bool(fused_sites > bulk_sites),
                
                "bulk_share_of_step": fused_ms / ms if ms else None,
                "serial_ms_per_step": serial_ms / args.steps,
                "uncounted_boundary_bytes_per_step": 16 * nb + 8 * wall_links,
                "peak_GBps": peak, "setup_seconds": setup, "stable": bool(mon["min_f"] > 0), "monitor": mon}
        print(json.dumps(line), flush=True)
        gpu.close()


if __name__ == "__main__":
    main()
