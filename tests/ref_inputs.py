"""The reference's own test inputs (tests/golden/ref_inputs/, copied by tests/golden/make_reference_inputs.py
from Code/tests/resources) as cases for the parity tests: the .gmy through ``geometry.read_gmy``, the few
numbers of the .xml that reach the collide-and-stream path converted to lattice units as the reference
converts them (Code/util/UnitConverter.cc:14-40, Code/lb/LbmParameters.h:35, SimBuilder.cc:66-69,106-109;
SURVEY.md Appendix A).  Test infrastructure: the XML reader itself (configuration::SimConfig) is out of
scope of the build."""
from __future__ import annotations

import functools
import os
import xml.etree.ElementTree as ET

import numpy as np

from hemelb_b200 import geometry as G
from hemelb_b200.capi import iolet_record
from hemelb_b200.lbm import prepare_boundary_objects

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_inputs")
MMHG_TO_PASCAL = 133.3223874  # Code/constants.h
ETA, RHO = 0.004, 1000.0      # default viscosity (Pa s) and density (kg/m3), Code/constants.h:21-22
CS2 = 1.0 / 3.0


def _vec(text):
    return np.array([float(x) for x in text.strip("()").split(",")])


@functools.lru_cache(maxsize=None)
def load(name: str):
    """(geometry, tau, rho0, inlet records, outlet records) of a reference fixture."""
    geom = G.read_gmy(os.path.join(HERE, name + ".gmy"))
    root = ET.parse(os.path.join(HERE, name + ".xml")).getroot()
    sim = root.find("simulation")
    dt = float(sim.find("step_length").get("value"))
    dx = float(sim.find("voxel_size").get("value"))
    origin = _vec(sim.find("origin").get("value"))
    rho_phys = float(sim.find("fluid_density").get("value")) if sim.find("fluid_density") is not None else RHO
    tau = 0.5 + (dt * ETA / rho_phys) / (CS2 * dx * dx)   # LbmParameters.h:35
    lattice_pressure = rho_phys * dx * dx / (dt * dt)       # UnitConverter: one lattice pressure unit in Pa

    def density(p_mmhg):  # ConvertPressureToLatticeUnits(p) / Cs2, reference pressure 0 mmHg
        return (CS2 + p_mmhg * MMHG_TO_PASCAL / lattice_pressure) / CS2

    def records(tag):
        out = []
        for io in root.find(tag + "s").findall(tag):
            c = io.find("condition")
            assert c.get("type") == "pressure" and c.get("subtype") == "cosine", "fixture uses another iolet kind"
            mean = float(c.find("mean").get("value"))
            amp = float(c.find("amplitude").get("value"))
            period = float(c.find("period").get("value")) / dt
            phase = float(c.find("phase").get("value"))
            normal = _vec(io.find("normal").get("value"))
            position = (_vec(io.find("position").get("value")) - origin) / dx
            out.append(iolet_record(0, tuple(normal), tuple(position), radius=1.0, density_mean=density(mean),
                                    density_amp=amp * MMHG_TO_PASCAL / lattice_pressure / CS2, phase=phase, period=period))
        return out

    ic = root.find("initialconditions")
    p0 = float(ic.find("pressure").find("uniform").get("value")) if ic is not None else 0.0
    inlets, outlets = records("inlet"), records("outlet")
    prepare_boundary_objects(inlets, outlets)
    geom.meta.setdefault("kind", name)
    return geom, tau, density(p0), inlets, outlets
