"""ctypes binding of ``libhemelb_b200.so`` (the C ABI declared in ``include/hemelb_b200.h``).

There is no fallback: if the library is missing, or there is no CUDA device, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HLB_LIB", os.path.join(HERE, "libhemelb_b200.so"))  # HLB_LIB: an experimental build

KERNELS = {"LBGK": 0, "MRT": 1, "TRT": 2}
WALLS = {"SBB": 0, "SIMPLEBOUNCEBACK": 0, "BFL": 1, "GZS": 2}
IOLETS = {"NASH": 0, "NASHZEROTHORDERPRESSUREIOLET": 0, "LADD": 1, "LADDIOLET": 1}
CACHES = {"density": 1, "velocity": 2, "wall_shear_stress": 4, "von_mises": 8, "shear_rate": 16,
          "stress_tensor": 32, "traction": 64, "tangential_traction": 128}
CACHE_WIDTH = {1: 1, 2: 3, 4: 1, 8: 1, 16: 1, 32: 9, 64: 3, 128: 3}

# every symbol include/hemelb_b200.h declares
SYMBOLS = [
    "hlb_gpu_last_error", "hlb_gpu_device_count", "hlb_gpu_create", "hlb_gpu_destroy",
    "hlb_gpu_set_neighbour_indices", "hlb_gpu_set_site_data", "hlb_gpu_set_wall_distances",
    "hlb_gpu_set_wall_normals", "hlb_gpu_set_site_coords", "hlb_gpu_set_neighbours",
    "hlb_gpu_set_streaming_indices", "hlb_gpu_set_iolets", "hlb_gpu_set_gzs_remote", "hlb_gpu_set_gzs_serve", "hlb_gpu_exchange_site_halo", "hlb_gpu_get_gzs_send",
    "hlb_gpu_set_gzs_ghost", "hlb_gpu_finalise",
    "hlb_gpu_comm_unique_id", "hlb_gpu_comm_init", "hlb_gpu_set_f", "hlb_gpu_get_f", "hlb_gpu_get_halo", "hlb_gpu_set_halo",
    "hlb_gpu_set_equilibrium", "hlb_gpu_request_comms", "hlb_gpu_copy_received", "hlb_gpu_swap",
    "hlb_gpu_set_step_scalars", "hlb_gpu_stream_and_collide", "hlb_gpu_post_step", "hlb_gpu_edge_done",
    "hlb_gpu_get_cache", "hlb_gpu_step", "hlb_gpu_get_time_step", "hlb_gpu_sync", "hlb_gpu_time_steps", "hlb_gpu_time_steps_detail",
    "hlb_gpu_monitor", "hlb_gpu_monitor_begin", "hlb_gpu_monitor_end", "hlb_gpu_monitor_global", "hlb_gpu_stability", "hlb_gpu_launch_count", "hlb_gpu_target_runs", "hlb_gpu_get_neighbour_indices", "hlb_gpu_set_overlap",
    # device-side Domain construction
    "hlb_dom_create", "hlb_dom_destroy", "hlb_dom_set_sites", "hlb_dom_set_shape", "hlb_dom_set_roughness", "hlb_dom_gzs_needs", "hlb_dom_lookup_sites", "hlb_dom_set_partition_slabs",
    "hlb_dom_set_partition_blocks", "hlb_dom_count_block_sites", "hlb_dom_count_block_sites_typed", "hlb_dom_build", "hlb_dom_build_seconds",
    "hlb_dom_get_counts", "hlb_dom_get_neighbours", "hlb_dom_get_streaming_indices", "hlb_dom_get_neighbour_indices",
    "hlb_dom_get_site_coords", "hlb_dom_get_input_index", "hlb_dom_get_boundary_tables", "hlb_dom_get_geometry_sizes",
    "hlb_dom_get_geometry", "hlb_gpu_create_from_domain",
    # property extraction and checkpoints
    "hlb_xtr_create", "hlb_xtr_create_from_domain", "hlb_xtr_destroy", "hlb_xtr_sizes", "hlb_xtr_required_caches",
    "hlb_xtr_header", "hlb_xtr_encode", "hlb_xtr_pinned_buffer", "hlb_xtr_last_encode_ms",
    "hlb_gpu_load_distributions", "hlb_gpu_load_distributions_from_domain",
    # METIS-free partition of the site graph (host code)
    "hlb_part_refine_kway", "hlb_part_bisect",
]


class HlbConfig(C.Structure):
    _fields_ = [("lattice", C.c_int), ("kernel", C.c_int), ("wall", C.c_int), ("inlet", C.c_int),
                ("outlet", C.c_int), ("tau", C.c_double), ("device", C.c_int), ("rank", C.c_int),
                ("nranks", C.c_int), ("n_sites", C.c_int64), ("mid_count", C.c_int64 * 6),
                ("edge_count", C.c_int64 * 6), ("total_shared_fs", C.c_int64), ("n_neighbours", C.c_int),
                ("n_inlets", C.c_int), ("n_outlets", C.c_int), ("reorder", C.c_int)]


class HlbError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HlbError("%s is missing: build it with `python -m hemelb_b200.build` (no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.hlb_gpu_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise HlbError(lib().hlb_gpu_last_error().decode())


def ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def iolet_record(kind=0, normal=(0, 0, 1), position=(0, 0, 0), radius=1.0, max_speed=0.0, density_mean=1.0,
                 density_amp=0.0, phase=0.0, period=1000.0, warmup=0.0, min_density=1.0):
    return np.array([kind, *normal, *position, radius, max_speed, density_mean, density_amp, phase, period,
                     warmup, min_density, 0.0], np.float64)
