#!/usr/bin/env python
"""Measurement of the extraction / checkpoint row: record bytes per second
through ``hlb_xtr_encode`` on one B200, with the reference's own writer timed beside it.

  python bench_extraction.py [--radius 146 --length 300] [--repeats 5]

One JSON line per output file kind (checkpoint; pressure + velocity; wall surface stresses):
  kernel_GBps    algorithmic bytes of xtr_encode_kernel (record bytes written + the cache /
                 distribution values and the 16 B of site id + coordinates read, per included site)
                 / its CUDA-event time on the engine's stream, and that as a fraction of the
                 measured HBM copy bandwidth (MEASURED_PEAKS.json)
  e2e_GBps       record bytes / wall time of encode + D2H into the handle's pinned buffer, in the
                 chunks GpuLocalPropertyOutput.write uses (no file system in the timed region)
  file_GBps      the same through GpuLocalPropertyOutput.write into a tmpfs file
  cpu_baseline   the reference's LocalPropertyOutput::Write (oracle/_ref, unmodified sources) on
                 the host cores, one emulated rank per core, tmpfs file, on a bounded sample
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

Q = 19
DT, DX, RHO, ETA = 1e-4, 1e-4, 1000.0, 0.004
ORIGIN = (0.0, 0.0, 0.0)
KINDS = {
    # name: (selector, fields, bytes read per included site besides the 16 B list entry)
    "checkpoint": ("whole", [("distributions", "distributions", "double", ())], 8 * Q + 4),
    "pressure_velocity": ("whole", [("Pressure", "pressure", "float", (80.0,)), ("Velocity", "velocity", "float", ())], 32),
    "surface_stresses": ("surface", [("ShearStress", "shearstress", "float", ()), ("Traction", "traction", "float", ()),
                                     ("TangTraction", "tangentialprojectiontraction", "float", ())], 8 + 24 + 24 + 24 + 4),
}


def tmpdir():
    return tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)


def gpu_lines(args, peak):
    from hemelb_b200 import geometry as G
    from hemelb_b200.capi import check, iolet_record, lib
    from hemelb_b200.domain import build_domains
    from hemelb_b200.extraction import GpuLocalPropertyOutput, OutputField, PropertyOutputFile, Units
    from hemelb_b200.lbm import GpuLBM, prepare_boundary_objects
    geom = G.cylinder_extruded(args.radius, args.length)
    dom = build_domains(geom, Q)[0]
    inl, outl = geom.meta["inlets"][0], geom.meta["outlets"][0]
    ins = [iolet_record(0, tuple(inl.normal), tuple(inl.position), radius=args.radius, density_mean=1.0005)]
    outs = [iolet_record(0, tuple(outl.normal), tuple(outl.position), radius=args.radius, density_mean=0.9995)]
    prepare_boundary_objects(ins, outs)
    gpu = GpuLBM(dom, "LBGK", "BFL", "NASH", "NASH", tau=0.8, inlets=ins, outlets=outs)
    gpu.set_equilibrium(1.0, (0.0, 0.0, 0.0))
    gpu.set_cache_mask(255)
    gpu.step(3)
    gpu.sync()
    L = lib()
    out = []
    for name, (selector, fields, read_bytes) in KINDS.items():
        d = tmpdir()
        spec = PropertyOutputFile(os.path.join(d, name + ".xtr"), 1, selector, (),
                                  [OutputField(n, s, t, o) for (n, s, t, o) in fields])
        t0 = time.perf_counter()
        po = GpuLocalPropertyOutput(gpu, spec, Units(DT, DX, ORIGIN, RHO, 80.0), chunk_sites=args.chunk_sites)
        create_s = time.perf_counter() - t0
        n, sl = po.local_site_count, po.site_len
        cap = min(args.chunk_sites, n) * sl
        p = C.c_void_p()
        check(L.hlb_xtr_pinned_buffer(po.x, C.c_uint64(cap), C.byref(p)))
        ms = C.c_float()

        def encode_all():
            k = 0.0
            for s0 in range(0, n, args.chunk_sites):
                m = min(args.chunk_sites, n - s0)
                check(L.hlb_xtr_encode(po.x, C.c_uint64(s0), C.c_uint64(m), p, C.c_uint64(cap)))
                check(L.hlb_xtr_last_encode_ms(po.x, C.byref(ms)))
                k += ms.value
            return k
        encode_all()
        kms, walls = [], []
        for _ in range(args.repeats):
            t0 = time.perf_counter()
            kms.append(encode_all())
            walls.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        po.write(0, 10)
        file_s = time.perf_counter() - t0
        size = os.path.getsize(spec.filename)
        po.close()
        for fn in os.listdir(d):
            os.remove(os.path.join(d, fn))
        os.rmdir(d)
        rec = n * sl
        alg = n * (sl + 16 + read_bytes)
        k = float(np.median(kms)) * 1e-3
        out.append({"output": name, "selector": selector, "fields": [f[0] for f in fields], "sites_written": n,
                    "site_record_bytes": sl, "record_bytes": rec, "file_bytes": size,
                    "kernel_ms": k * 1e3, "kernel_algorithmic_bytes": alg, "kernel_GBps": alg / k / 1e9,
                    "kernel_frac_of_hbm_peak": alg / k / 1e9 / peak, "peak_GBps": peak,
                    "e2e_GBps": rec / float(np.median(walls)) / 1e9, "e2e_ms": float(np.median(walls)) * 1e3,
                    "file_GBps": size / file_s / 1e9, "selector_and_site_list_seconds": create_s,
                    "chunk_sites": args.chunk_sites, "workload": "cylinder r=%g l=%d (%d sites), D3Q19 LBGK+BFL" % (
                        args.radius, args.length, dom.N)})
    return out


def cpu_lines(args):
    import oracle as O
    from oracle import xtr as X
    from hemelb_b200 import geometry as G
    from hemelb_b200.capi import iolet_record
    from hemelb_b200.domain import build_domains
    from hemelb_b200.lbm import prepare_boundary_objects
    if O.ref_lib() is None or not hasattr(O.ref_lib(), "href_xtr_open"):
        return {}
    cores = os.cpu_count() or 1
    radius, length = 40.0, 320
    geom = G.cylinder_extruded(radius, length)
    rank = G.slab_decomposition(geom, cores) if cores > 1 else None
    doms = build_domains(geom, Q, rank, cores)
    inl, outl = geom.meta["inlets"][0], geom.meta["outlets"][0]
    ins = [iolet_record(0, tuple(inl.normal), tuple(inl.position), radius=radius, density_mean=1.0005)]
    outs = [iolet_record(0, tuple(outl.normal), tuple(outl.position), radius=radius, density_mean=0.9995)]
    prepare_boundary_objects(ins, outs)
    sim = O.RefSim([d.tables() for d in doms], Q, "LBGK", "BFL", "NASH", "NASH", dt=DT, dx=DX, rho=RHO, eta=ETA,
                   inlets=ins, outlets=outs)
    _, w, _ = O.lattice(Q)
    for r, d in enumerate(doms):
        f = np.zeros(d.N * Q + 1 + d.totalSharedFs)
        f[:d.N * Q] = np.tile(w, d.N)
        sim.set_f(f, r)
        sim.set_f(f, r, 1)
    sim.set_cache_mask(255)
    sim.step_mt(2)
    res = {}
    for name, (selector, fields, _) in KINDS.items():
        d = tmpdir()
        path = os.path.join(d, name + ".xtr")
        s = sim.xtr_open(path, [X.Field(n, sname, t, o) for (n, sname, t, o) in fields], selector, (), dt=DT, dx=DX,
                         origin=ORIGIN, fluid_density=RHO, reference_pressure=80.0)
        sim.xtr_write(s, 0)
        size0 = os.path.getsize(path)
        reps = 0
        t0 = time.perf_counter()
        while reps < 3 or time.perf_counter() - t0 < 2.0:
            reps += 1
            sim.xtr_write(s, reps)
        dtm = time.perf_counter() - t0
        size = os.path.getsize(path)
        sim.xtr_close(s)
        for fn in os.listdir(d):
            os.remove(os.path.join(d, fn))
        os.rmdir(d)
        res[name] = {"value": (size - size0) / dtm / 1e9, "unit": "GB/s of record bytes", "cores": cores, "kind": "reference",
                     "sample": "cylinder r=%g l=%d (%d sites), %d writes, %d emulated ranks (threads), tmpfs file" % (
                         radius, length, geom.n_sites, reps, cores)}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--radius", type=float, default=146.0)
    ap.add_argument("--length", type=int, default=300)
    ap.add_argument("--repeats", type=int, default=5)
    ap.add_argument("--chunk-sites", type=int, default=1 << 21)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    lines = gpu_lines(args, peak)
    base = {} if args.no_cpu_baseline else cpu_lines(args)
    for ln in lines:
        if ln["output"] in base:
            ln["cpu_baseline"] = base[ln["output"]]
        print(json.dumps(ln), flush=True)


if __name__ == "__main__":
    main()
