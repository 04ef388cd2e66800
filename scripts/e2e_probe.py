"""Where the end-to-end figure's distance from the device-resident one goes: the phase API driven from Python
(a) without monitors, (b) with the in-kernel gathering switched on but never read, (c) read every step
(blocking), (d) read every step one step behind (hlb_gpu_monitor_begin / _end).  One GPU, bench.py's tree."""
import json
import sys
import time

sys.path.insert(0, ".")
import bench  # noqa: E402
from hemelb_b200.lbm import GpuLBM  # noqa: E402

dom, inlets, outlets, _ = bench.build_tree(1.1e8, 0, 1, 0, "basic", "morton", None)
gpu = GpuLBM.from_device_domain(dom, "LBGK", "BFL", "NASH", "NASH", tau=bench.TAU, inlets=inlets, outlets=outlets)
gpu.set_equilibrium(1.0, (0.0, 0.0, 0.0))
gpu.step(20)
steps = 60
out = {"sites": int(dom.N)}
ms = gpu.time_steps(steps)
out["device_resident"] = dom.N * steps / (ms * 1e-3) / 1e6


def run(label, mask, mode):
    gpu.set_cache_mask(mask)
    gpu.do_time_step()
    gpu.sync()
    t0 = time.perf_counter()
    pending = False
    for _ in range(steps):
        gpu.do_time_step()
        if mode == "blocking":
            gpu.monitor()
        elif mode == "lagged":
            if pending:
                gpu.monitor_end()
            gpu.monitor_begin()
            pending = True
    if pending:
        gpu.monitor_end()
    gpu.sync()
    out[label] = dom.N * steps / (time.perf_counter() - t0) / 1e6


run("phase_api_no_monitor", 0, None)
run("phase_api_gathering_not_read", 256, None)
run("phase_api_blocking_read", 256, "blocking")
run("phase_api_lagged_read", 256, "lagged")
gpu.set_cache_mask(256)
ms = gpu.time_steps(steps)
out["device_resident_with_gathering"] = dom.N * steps / (ms * 1e-3) / 1e6
print(json.dumps(out))
