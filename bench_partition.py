#!/usr/bin/env python
"""Host-side figures of the METIS-free decomposition: cut lattice links, weighted
imbalance, neighbour pairs and halo size per start and stage on the synthetic tree, against the
reference's BasicDecomposition.  No GPU involved; one JSON line per (ranks, method).

  python bench_partition.py [--root-radius 22 --root-length 80 --generations 4] > profiles/r01_partition_quality.jsonl
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--generations", type=int, default=4)
    ap.add_argument("--root-radius", type=float, default=22.0)
    ap.add_argument("--root-length", type=float, default=80.0)
    ap.add_argument("--lattice", type=int, default=19)
    ap.add_argument("--wall", default="BFL")
    args = ap.parse_args()
    from hemelb_b200 import geometry as G
    from hemelb_b200 import partition as P
    from hemelb_b200.domain import build_domains

    geom = G.capsule_tree(args.generations, args.root_radius, args.root_length)
    Q = args.lattice
    t = build_domains(geom, Q)[0].tables()
    wall = np.asarray(t["wallMask"]) != 0
    st = np.asarray(t["siteType"])
    local = np.where(st == 2, np.where(wall, 4, 2), np.where(st == 3, np.where(wall, 5, 3), np.where(wall, 1, 0)))
    types = np.empty(geom.n_sites, np.int64)
    types[np.asarray(t["inputIndex"])] = local
    xadj, adjncy = P.site_graph(geom, Q)
    vw = P.site_weights(args.wall, "NASH", "NASH")[types]
    src = np.repeat(np.arange(geom.n_sites), np.diff(xadj))

    def line(R, method, part, seconds):
        cut = part[src] != part[adjncy]
        pairs = len(set(zip(part[src][cut].tolist(), part[adjncy][cut].tolist()))) // 2
        q = P.site_quality(xadj, adjncy, vw, part, R)
        print(json.dumps({"sites": geom.n_sites, "adjacencies": int(adjncy.size), "lattice": Q, "ranks": R, "method": method,
                          "cut_links": q["edge_cut"], "halo_doubles_total": 2 * q["edge_cut"],
                          "weighted_imbalance": round(q["imbalance"], 5), "neighbour_pairs": pairs,
                          "seconds": round(seconds, 3)}), flush=True)

    for R in (2, 4, 8):
        t0 = time.time()
        line(R, "BasicDecomposition (reference, whole blocks)", G.basic_decomposition(geom, R), time.time() - t0)
        for start in P.STARTS:
            t0 = time.time()
            blocks, _ = P.partition_geometry(geom, types, args.wall, nranks=R, initial=start)
            line(R, "weighted blocks, %s start" % start, blocks, time.time() - t0)
            first = blocks if start == "morton" else P.coordinate_bisection_native(geom.coords, vw, R, start == "inertial")
            t0 = time.time()
            sites, _ = P.refine_sites_native(xadj, adjncy, vw, first, R)
            line(R, "site stage (hlb_part_refine_kway), %s start" % start, sites, time.time() - t0)


if __name__ == "__main__":
    main()
