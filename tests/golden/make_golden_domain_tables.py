"""Writes tests/golden/domain_tables_*.npz: the per-rank index tables that the reference's own geometry::Domain
(oracle/_ref/libhemelb_refdom.so: Domain.cc, LookupTree.cc, BasicDecomposition.cc compiled unmodified, emulated
ranks) builds for the synthetic test geometries.  Run where /root/reference exists; the files are committed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from tests.cases import geometry  # noqa: E402
from tests.test_domain_vs_ref import GOLDEN, KEYS, decomposition  # noqa: E402

if __name__ == "__main__":
    for name, Q, R, kind in GOLDEN:
        geom = geometry(name)
        ref = O.RefDomains(geom, Q, decomposition(geom, R, kind), R)
        out = {}
        for r in range(R):
            t = ref.tables(r)
            for k in KEYS:
                out["r%d_%s" % (r, k)] = t[k]
            out["r%d_N" % r] = np.int64(t["N"])
            out["r%d_totalSharedFs" % r] = np.int64(t["totalSharedFs"])
        path = os.path.join(ROOT, "tests", "golden", "domain_tables_%s_q%d_r%d.npz" % (name, Q, R))
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), "bytes")
