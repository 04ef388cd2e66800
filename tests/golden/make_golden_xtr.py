"""Generate tests/golden/xtr_*.{xtr,off} from the compiled reference (oracle/_ref).

Run HERE (where /root/reference exists): `python tests/golden/make_golden_xtr.py`.  The files are
written by the UNMODIFIED reference extraction sources (LocalPropertyOutput.cc, LbDataSourceIterator.cc,
the selectors, the XDR writers; oracle/ref_xtr.h) from the reference's own streamers' state after
STEPS steps of the four-cube fixture (D3Q15 LBGK, SimpleBounceBack, Nash iolets -- configs[0]) and
of the small cylinder (D3Q19 LBGK + BFL, 2 emulated ranks).  The GPU extraction path has to
reproduce them byte for byte from the same inputs on a box that has no reference checkout.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle as O  # noqa: E402
from tests.xtr_cases import CASES, DT, DX, ORIGIN, REF_PRESSURE, RHO, make_sim, xfields  # noqa: E402


def main():
    O.build()
    assert O.ref_lib() is not None, "oracle/_ref not built (needs /root/reference)"
    gold = os.path.join(ROOT, "tests", "golden")
    for name, case in CASES.items():
        ref, T = make_sim("ref", case)
        base = os.path.join(gold, "xtr_" + name)
        for ext in (".xtr", ".off"):
            if os.path.exists(base + ext):
                os.remove(base + ext)
        s = ref.xtr_open(base + ".xtr", xfields(case["fields"]), case["selector"], case["params"], frequency=case["frequency"],
                         dt=DT, dx=DX, origin=ORIGIN, fluid_density=RHO, reference_pressure=REF_PRESSURE)
        t = case["steps"]
        for more in case["writes"]:
            ref.step(more)
            t += more
            ref.xtr_write(s, t)
        ref.xtr_close(s)
        print(name, os.path.getsize(base + ".xtr"), "bytes")


if __name__ == "__main__":
    main()
