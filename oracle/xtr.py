"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's property-extraction writer and
checkpoint reader (the ``.xtr`` / ``.off`` path), used as the checker of
``hlb_xtr_*``.  Byte-for-byte the same files as the reference's own sources compiled into
``oracle/_ref`` (``tests/test_xtr_oracle.py``) and the known-answer header of
``Code/tests/extraction/LocalPropertyOutputTests.cc:152-199``.

Follows (all under /root/reference/Code):
  extraction/LocalPropertyOutput.cc:67-131 (offsets, lengths), :133-213 (headers), :262-367 (records),
                                    :369-396 (offset file)
  extraction/LbDataSourceIterator.cc:36-87   (cache -> physical units, float narrowing)
  util/UnitConverter.{h,cc}                   (conversion arithmetic and its operand types)
  extraction/{Whole,GeometrySurface,Plane,StraightLine,SurfacePoint}GeometrySelector.cc, GeometrySelector.cc
  extraction/LocalDistributionInput.cc:40-165 (checkpoint read), :252-308 (offset read)
  io/formats/{formats,extraction,offset}.h, doc/dev/file-formats/{extraction,offset}.md
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import struct

import numpy as np

HEMELB_MAGIC = 0x686C6221
XTR_MAGIC = 0x78747204
XTR_VERSION = 5
OFF_MAGIC = 0x6F666604
OFF_VERSION = 1
MAIN_HEADER_LENGTH = 60
MMHG_TO_PASCAL = 133.3223874
CS2 = 1.0 / 3.0

# extraction::source::Type order (OutputField.h:18-40) and field lengths (LocalPropertyOutput.cc:398-431)
SOURCES = ["pressure", "velocity", "shearstress", "vonmisesstress", "shearrate", "stresstensor", "traction",
           "tangentialprojectiontraction", "distributions", "mpirank"]
# io::formats::extraction::TypeCode
TYPECODES = {"float": 0, "double": 1, "int32": 2, "uint32": 3, "int64": 4, "uint64": 5}
_BE = {0: ">f4", 1: ">f8", 2: ">i4", 3: ">u4", 4: ">i8", 5: ">u8"}
_NATIVE = {0: np.float32, 1: np.float64, 2: np.int32, 3: np.uint32, 4: np.int64, 5: np.uint64}
SELECTORS = {"whole": 0, "surface": 1, "plane": 2, "line": 3, "surfacepoint": 4}


def field_length(source, Q):
    return {"pressure": 1, "velocity": 3, "shearstress": 1, "vonmisesstress": 1, "shearrate": 1, "stresstensor": 6,
            "traction": 3, "tangentialprojectiontraction": 3, "distributions": Q, "mpirank": 1}[source]


class UnitConverter:
    """util/UnitConverter.cc:14-24 -- same operand order."""

    def __init__(self, dt, dx, origin, fluid_density, reference_pressure):
        self.latticeDistance = float(dx)
        self.latticeTime = float(dt)
        self.latticeMass = fluid_density * dx * dx * dx
        self.latticeSpeed = dx / dt
        self.origin = np.asarray(origin, np.float64)
        self.latticePressure = self.latticeMass / (self.latticeDistance * self.latticeTime * self.latticeTime)
        self.reference_pressure_mmHg = float(reference_pressure)


def stored_string_length(s):
    n = len(s)
    if n % 4:
        n += 4 - n % 4
    return n + 4


def field_header_length(name, noffsets, typecode):
    return stored_string_length(name) + 12 + (4 if typecode in (0, 2, 3) else 8) * noffsets


def xdr_string(s):
    b = s.encode()
    return struct.pack(">I", len(b)) + b + b"\0" * ((4 - len(b) % 4) % 4)


class Field:
    def __init__(self, name, source, typecode="float", offsets=()):
        self.name, self.source = name, source
        self.typecode = TYPECODES[typecode] if isinstance(typecode, str) else int(typecode)
        self.offsets = [float(x) for x in offsets]


def site_write_length(fields, Q):
    n = 12
    for f in fields:
        ln = field_length(f.source, Q)
        if len(f.offsets) not in (0, 1, ln):
            raise ValueError("Invalid length of offsets array %d" % len(f.offsets))
        n += ln * np.dtype(_BE[f.typecode]).itemsize
    return n


def header_bytes(fields, Q, conv, global_site_count):
    """LocalPropertyOutput::PrepareHeader."""
    fh = b""
    for f in fields:
        fh += xdr_string(f.name) + struct.pack(">III", field_length(f.source, Q), f.typecode, len(f.offsets))
        fh += np.asarray(f.offsets, np.float64).astype(_NATIVE[f.typecode]).astype(_BE[f.typecode]).tobytes()
    assert len(fh) == sum(field_header_length(f.name, len(f.offsets), f.typecode) for f in fields)
    main = struct.pack(">IIIddddQII", HEMELB_MAGIC, XTR_MAGIC, XTR_VERSION, conv.latticeDistance, *conv.origin,
                       global_site_count, len(fields), len(fh))
    assert len(main) == MAIN_HEADER_LENGTH
    return main + fh


# ------------------------------------------------------------------------------------ selectors
def _f32(x):
    return np.asarray(x, np.float32)


def lattice_to_physical(coords, conv):
    """GeometrySelector::LatticeToPhysical: Vector3D<float>{location} * float(voxel) + origin.as<float>()."""
    return coords.astype(np.float32) * np.float32(conv.latticeDistance) + conv.origin.astype(np.float32)


def _dot(a, b):  # std::inner_product with a float zero: ((0 + a0 b0) + a1 b1) + a2 b2
    acc = np.zeros(a.shape[:-1], np.float32)
    for k in range(3):
        acc = acc + a[..., k] * b[..., k]
    return acc


def select(kind, params, coords, is_wall, conv):
    """GeometrySelector::Include for sites that are valid and local: the five IsWithinGeometry bodies."""
    kind = SELECTORS[kind] if isinstance(kind, str) else kind
    n = coords.shape[0]
    p = _f32(params)
    voxel = conv.latticeDistance
    if kind == 0:
        return np.ones(n, bool)
    if kind == 1:
        return is_wall.astype(bool).copy()
    x = lattice_to_physical(coords, conv)
    if kind == 2:  # PlaneGeometrySelector.cc:52-74
        point, normal, radius = p[0:3], p[3:6], p[6]
        normal = normal / np.sqrt(_dot(normal, normal))
        perp = _dot(x - point, np.broadcast_to(normal, x.shape))
        inc = ~(np.abs(perp).astype(np.float64) > 0.5 * voxel)
        if radius > 0:
            r = (x - normal[None, :] * perp[:, None]) - point
            r2 = _dot(r, r)
            inc &= r2 <= radius * radius
        return inc
    if kind == 3:  # StraightLineGeometrySelector.cc:34-56
        e1 = p[0:3]
        line = p[3:6] - e1
        length = np.sqrt(_dot(line, line))
        along = _dot(np.broadcast_to(line, x.shape), x - e1) / length
        inc = ~((along.astype(np.float64) < 0.0) | (along > length))
        q = (e1 + line[None, :] * along[:, None] / length) - x
        d2 = _dot(q, q)
        return inc & (d2.astype(np.float64) <= 2.0 * 0.5 * 0.5 * voxel * voxel)
    # SurfacePointSelector.cc:28-44
    d = x - p[0:3]
    dist = np.sqrt(_dot(d, d)).astype(np.float64) / voxel
    return is_wall.astype(bool) & (dist <= np.float64(np.float32(np.sqrt(3.0))))


# ------------------------------------------------------------------------------------ records
def _nan_as_x86(a):
    """NaNs born from invalid operations carry x86's 'real indefinite' pattern (sign bit set)."""
    a = a.copy()
    if a.dtype == np.float64:
        a.view(np.uint64)[np.isnan(a)] = 0xFFF8000000000000
    elif a.dtype == np.float32:
        a.view(np.uint32)[np.isnan(a)] = 0xFFC00000
    return a


def field_values(f, sel, data, conv, Q, rank):
    """Values of one field for the selected sites as the reference hands them to write<FileT>():
    (n, len) array in the C++ type the expression has before the cast to the file type."""
    src = f.source
    with np.errstate(all="ignore"):
        if src == "pressure":  # float GetPressure() - double offset[0]
            p = conv.reference_pressure_mmHg + ((data["density"][sel] * CS2 - CS2) * conv.latticePressure / MMHG_TO_PASCAL)
            v = p.astype(np.float32).astype(np.float64) - f.offsets[0]
            return v[:, None]
        if src == "velocity":  # velocityCache.as<float>() * float(latticeSpeed)
            return data["velocity"].reshape(-1, 3)[sel].astype(np.float32) * np.float32(conv.latticeSpeed)
        if src == "shearstress":
            return (data["wall_shear_stress"][sel] * conv.latticePressure).astype(np.float32)[:, None]
        if src == "vonmisesstress":
            return (data["von_mises"][sel] * conv.latticePressure).astype(np.float32)[:, None]
        if src == "shearrate":
            return (data["shear_rate"][sel] / conv.latticeTime).astype(np.float32)[:, None]
        if src == "stresstensor":  # Matrix3D * latticePressure, addDiagonal(ref * mmHg); upper triangle row-wise
            t = conv.latticePressure * data["stress_tensor"].reshape(-1, 3, 3)[sel]
            for k in range(3):
                t[:, k, k] += conv.reference_pressure_mmHg * MMHG_TO_PASCAL
            return np.stack([t[:, 0, 0], t[:, 0, 1], t[:, 0, 2], t[:, 1, 1], t[:, 1, 2], t[:, 2, 2]], 1)
        if src == "traction":  # traction * latticePressure; += wallNormal * ref * mmHg
            t = data["traction"].reshape(-1, 3)[sel] * conv.latticePressure
            return t + data["wallNormal"].reshape(-1, 3)[sel] * conv.reference_pressure_mmHg * MMHG_TO_PASCAL
        if src == "tangentialprojectiontraction":
            return data["tangential_traction"].reshape(-1, 3)[sel] * conv.latticePressure
        if src == "distributions":
            return data["f"][:data["N"] * Q].reshape(-1, Q)[sel]
        return np.full((int(np.count_nonzero(sel)) if sel.dtype == bool else len(sel), 1), rank, np.int32)


def _cast(v, typecode):
    """FileT(val) as x86-64 g++ does it for the combinations that are defined behaviour."""
    with np.errstate(all="ignore"):
        if typecode in (0, 1):
            return _nan_as_x86(_nan_as_x86(np.asarray(v)).astype(_NATIVE[typecode]))
        return np.asarray(v).astype(_NATIVE[typecode])


def rank_chunk(fields, sel_mask, data, conv, Q, rank, timestep=None):
    """One rank's bytes of one record (LocalPropertyOutput::Write); timestep only on the IO rank."""
    coords = data["globalCoords"].reshape(-1, 3)[sel_mask]
    n = coords.shape[0]
    site_len = site_write_length(fields, Q)
    rec = np.zeros((n, site_len), np.uint8)
    rec[:, 0:12] = coords.astype(np.uint32).astype(">u4").view(np.uint8).reshape(n, 12)
    o = 12
    for f in fields:
        v = _cast(field_values(f, sel_mask, data, conv, Q, rank), f.typecode).astype(_BE[f.typecode])
        w = v.shape[1] * v.dtype.itemsize
        rec[:, o:o + w] = np.ascontiguousarray(v).view(np.uint8).reshape(n, w)
        o += w
    head = b"" if timestep is None else struct.pack(">Q", timestep)
    return head + rec.tobytes()


class PropertyOutput:
    """LocalPropertyOutput over R emulated ranks: layout, header, offset file, records."""

    def __init__(self, fields, selector, sel_params, conv, Q, rank_data):
        self.fields, self.conv, self.Q, self.rank_data = fields, conv, Q, rank_data
        self.masks = [select(selector, sel_params, d["globalCoords"].reshape(-1, 3), d["wallMask"] != 0, conv)
                      for d in rank_data]
        self.local_counts = [int(m.sum()) for m in self.masks]
        self.global_count = sum(self.local_counts)
        self.site_len = site_write_length(fields, Q)
        self.header = header_bytes(fields, Q, conv, self.global_count)
        lens = [c * self.site_len + (8 if r == 0 else 0) for r, c in enumerate(self.local_counts)]
        ends = np.cumsum(lens) + len(self.header)
        self.local_len = lens
        self.local_start = [int(e - l) for e, l in zip(ends, lens)]
        self.global_len = self.site_len * self.global_count + 8

    def offset_file(self):
        R = len(self.rank_data)
        out = struct.pack(">IIIi", HEMELB_MAGIC, OFF_MAGIC, OFF_VERSION, R)
        for r in range(R):
            out += struct.pack(">Q", self.local_start[r])
        return out + struct.pack(">Q", self.local_start[-1] + self.local_len[-1])

    def record(self, timestep):
        return b"".join(rank_chunk(self.fields, self.masks[r], d, self.conv, self.Q, r, timestep if r == 0 else None)
                        for r, d in enumerate(self.rank_data))


# ------------------------------------------------------------------------------------ checkpoint
def read_offsets(off_bytes, R):
    magic, omagic, ver, n = struct.unpack(">IIIi", off_bytes[:16])
    if magic != HEMELB_MAGIC:
        raise ValueError("This file does not start with the HemeLB magic number.")
    if omagic != OFF_MAGIC:
        raise ValueError("This file does not have the offset magic number.")
    if ver != OFF_VERSION:
        raise ValueError("Version number incorrect.")
    if n != R:
        raise ValueError("Offset file has wrong number of MPI ranks.")
    return list(struct.unpack(">%dQ" % (R + 1), off_bytes[16:16 + 8 * (R + 1)]))


def load_checkpoint(xtr_bytes, off_bytes, Q, coords_per_rank, target=None):
    """LocalDistributionInput::LoadDistribution: (timestep, [f (N, Q) per rank])."""
    magic, xmagic, ver = struct.unpack(">III", xtr_bytes[:12])
    if magic != HEMELB_MAGIC:
        raise ValueError("This file does not start with the HemeLB magic number.")
    if xmagic != XTR_MAGIC:
        raise ValueError("This file does not have the extraction magic number.")
    if ver != XTR_VERSION:
        raise ValueError("Version number incorrect.")
    nsites, nfields, fhlen = struct.unpack(">QII", xtr_bytes[44:60])
    if nfields != 1:
        raise ValueError("Checkpoint file must contain exactly one field")
    if fhlen != 32:
        raise ValueError("Checkpoint file's field header must be 32 B long")
    name_len = struct.unpack(">I", xtr_bytes[60:64])[0]
    name = xtr_bytes[64:64 + name_len].decode()
    nel, tc, noff = struct.unpack(">III", xtr_bytes[60 + stored_string_length(name):60 + stored_string_length(name) + 12])
    if name != "distributions":
        raise ValueError("Checkpoint file must contain field named 'distributions'")
    if nel != Q:
        raise ValueError("Checkpoint field distributions contains %d distributions" % nel)
    if tc != 1:
        raise ValueError("Checkpoint contains wrong data type")
    if noff != 0:
        raise ValueError("Checkpoint should not have offsets")
    R = len(coords_per_rank)
    offs = read_offsets(off_bytes, R)
    all_len = offs[-1] - offs[0]
    data_size = len(xtr_bytes) - (MAIN_HEADER_LENGTH + 32)
    if data_size % all_len:
        raise ValueError("Checkpoint file length not consistent with integer number of checkpoints")
    ntimes = data_size // all_len
    times = [struct.unpack(">Q", xtr_bytes[offs[0] + i * all_len:offs[0] + i * all_len + 8])[0] for i in range(ntimes)]
    if target is None:
        its = ntimes - 1
        timestep = times[its]
    else:
        # LocalDistributionInput.cc:79-96, literally: a lower-bound search whose `timestep` is the
        # LAST PROBED record, not the one at iTS -- so a target that the search brackets from the
        # left last (e.g. 6 in [4, 6]) is reported "not found" by the reference.  Kept as is.
        its, length, timestep = 0, ntimes, None
        while length != 0:
            l2 = length // 2
            m = its + l2
            timestep = times[m]
            if timestep < target:
                its = m + 1
                length -= l2 + 1
            else:
                length = l2
        if timestep != target:
            raise ValueError("Target timestep %d not found in checkpoint file." % target)
    out = []
    for r in range(R):
        chunk = xtr_bytes[its * all_len + offs[r]:its * all_len + offs[r + 1]]
        if r == 0:
            chunk = chunk[8:]
        site_len = 12 + 8 * Q
        if len(chunk) % site_len:
            raise ValueError("ragged chunk")
        a = np.frombuffer(chunk, np.uint8).reshape(-1, site_len)
        c = a[:, :12].copy().view(">u4").astype(np.int64)
        want = np.asarray(coords_per_rank[r], np.int64).reshape(-1, 3)
        if c.shape[0] != want.shape[0]:
            raise ValueError("Read %d sites but expected %d" % (c.shape[0], want.shape[0]))
        if not np.array_equal(c, want):
            raise ValueError("Site read at the wrong index or rank")
        out.append(a[:, 12:].copy().view(">f8").astype(np.float64))
    return timestep, out
