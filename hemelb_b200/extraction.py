"""Host-side mirror of the reference's property-extraction and checkpoint classes over the
device encoder (``hlb_xtr_*`` / ``hlb_gpu_load_distributions``; ``csrc/extraction.cu``).

Same names and argument meaning as
  extraction::OutputField / PropertyOutputFile   Code/extraction/OutputField.h, PropertyOutputFile.h
  extraction::LocalPropertyOutput                Code/extraction/LocalPropertyOutput.{h,cc}
  extraction::PropertyWriter / PropertyActor     Code/extraction/PropertyWriter.cc, PropertyActor.cc
  extraction::LocalDistributionInput             Code/extraction/LocalDistributionInput.{h,cc}
The per-site work (selector, unit conversion, XDR encoding, record decoding) happens on the GPU;
what stays here is what the reference does once per file or per write on the host: the
AllReduce / Scan of the per-rank lengths, the header and ``.off`` file, and the positioned file
writes (POSIX pwrite instead of MPI-IO -- same bytes at the same offsets).
"""
from __future__ import annotations

import ctypes as C
import os
import struct
from dataclasses import dataclass, field

import numpy as np

from .capi import HlbError, check, lib, ptr

SOURCES = {"pressure": 0, "velocity": 1, "shearstress": 2, "vonmisesstress": 3, "shearrate": 4, "stresstensor": 5,
           "traction": 6, "tangentialprojectiontraction": 7, "distributions": 8, "mpirank": 9}
TYPECODES = {"float": 0, "double": 1, "int32": 2, "uint32": 3, "int64": 4, "uint64": 5}
SELECTORS = {"whole": 0, "surface": 1, "plane": 2, "line": 3, "surfacepoint": 4}
HEMELB_MAGIC, OFFSET_MAGIC, OFFSET_VERSION, OFFSET_HEADER_LENGTH = 0x686C6221, 0x6F666604, 1, 16
XTR_MAGIC, XTR_VERSION, MAIN_HEADER_LENGTH = 0x78747204, 5, 60
IO_RANK = 0  # net::IOCommunicator::IO_RANK


class XtrField(C.Structure):
    _fields_ = [("name", C.c_char_p), ("source", C.c_int), ("typecode", C.c_int), ("n_offsets", C.c_uint32),
                ("offsets", C.POINTER(C.c_double))]


class XtrSpec(C.Structure):
    _fields_ = [("selector", C.c_int), ("selector_params", C.c_float * 7), ("n_fields", C.c_int),
                ("fields", C.POINTER(XtrField)), ("time_step", C.c_double), ("voxel_size", C.c_double),
                ("origin", C.c_double * 3), ("fluid_density", C.c_double), ("reference_pressure", C.c_double)]


@dataclass
class OutputField:
    name: str
    src: str                      # extraction::source::*  (lower case)
    typecode: str = "float"       # io::formats::extraction::TypeCode
    offset: tuple = ()


@dataclass
class PropertyOutputFile:
    filename: str
    frequency: int
    geometry: str = "whole"       # selector kind
    geometry_params: tuple = ()
    fields: list = field(default_factory=list)
    single_timestep_files: bool = False   # file_timestep_mode


@dataclass
class Units:
    """The util::UnitConverter constructor arguments (Code/util/UnitConverter.cc:14-24)."""
    time_step: float
    voxel_size: float
    origin: tuple = (0.0, 0.0, 0.0)
    fluid_density: float = 1000.0
    reference_pressure: float = 0.0


class SingleComm:
    """One rank."""
    rank, size = 0, 1

    def allreduce_sum(self, v):
        return v

    def scan_sum(self, v):
        return v

    def broadcast(self, v, root=0):
        return v

    def scatter(self, values, root=0):
        return values[0]

    def barrier(self):
        pass


class TorchComm:
    """torch.distributed (gloo or nccl process group) as the reference's net::IOCommunicator."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)

    def _gather(self, v):
        out = [None] * self.size
        self.dist.all_gather_object(out, v, group=self.group)
        return out

    def allreduce_sum(self, v):
        return sum(self._gather(v))

    def scan_sum(self, v):  # inclusive, as MPI_Scan
        return sum(self._gather(v)[:self.rank + 1])

    def broadcast(self, v, root=0):
        box = [v]
        self.dist.broadcast_object_list(box, src=root, group=self.group)
        return box[0]

    def scatter(self, values, root=0):
        return self.broadcast(values, root)[self.rank]

    def barrier(self):
        self.dist.barrier(group=self.group)


def extraction_to_offset(path: str) -> str:
    """io::formats::offset::ExtractionToOffset (Code/io/formats/offset.h:47-52)."""
    i = path.rfind(".")
    if i < 0:
        raise HlbError("Cannot split extension from extraction filename")
    return path[:i] + ".off"


def write_layout(comms, local_site_count: int, site_len: int, header_length: int) -> dict:
    """Where each rank writes (LocalPropertyOutput.cc:96-110): only the IO rank writes the 8-byte
    time stamp; every rank's start = header + exclusive prefix sum of the local lengths."""
    global_site_count = comms.allreduce_sum(local_site_count)                                    # :99
    local_len = local_site_count * site_len + (8 if comms.rank == IO_RANK else 0)                # :104
    global_len = site_len * global_site_count + 8                                                # :106
    local_write_end = comms.scan_sum(local_len) + header_length                                  # :109
    return dict(global_site_count=global_site_count, local_data_write_length=local_len,
                global_data_write_length=global_len, local_write_start=local_write_end - local_len)


def _create_handle(lbm, spec: PropertyOutputFile, units: Units):
    L = lib()
    fields = (XtrField * max(1, len(spec.fields)))()
    keep = []
    for i, f in enumerate(spec.fields):
        offs = np.ascontiguousarray(f.offset, np.float64)
        keep.append(offs)
        fields[i].name = f.name.encode()
        fields[i].source = SOURCES[f.src]
        fields[i].typecode = TYPECODES[f.typecode]
        fields[i].n_offsets = offs.size
        fields[i].offsets = ptr(offs, C.c_double) if offs.size else None
    cs = XtrSpec()
    cs.selector = SELECTORS[spec.geometry]
    for k, v in enumerate(spec.geometry_params):
        cs.selector_params[k] = v
    cs.n_fields = len(spec.fields)
    cs.fields = fields
    cs.time_step, cs.voxel_size = units.time_step, units.voxel_size
    for k in range(3):
        cs.origin[k] = units.origin[k]
    cs.fluid_density, cs.reference_pressure = units.fluid_density, units.reference_pressure
    x = C.c_void_p()
    dom = lbm.domain
    if hasattr(dom, "d"):  # devdomain.DeviceDomain: coordinates are already on the device
        check(L.hlb_xtr_create_from_domain(lbm.h, dom.d, C.byref(cs), C.byref(x)))
    else:
        coords = np.ascontiguousarray(dom.globalCoords, np.int64)
        check(L.hlb_xtr_create(lbm.h, C.byref(cs), ptr(coords, C.c_int64), C.byref(x)))
    return x


class GpuLocalPropertyOutput:
    """extraction::LocalPropertyOutput: stores sufficient information to output property
    information from this rank's GPU (LocalPropertyOutput.h:26-100)."""

    def __init__(self, lbm, output_spec: PropertyOutputFile, units: Units, comms=None, chunk_sites: int = 1 << 20):
        self.L = lib()
        self.lbm, self.spec, self.comms = lbm, output_spec, comms or SingleComm()
        self.chunk_sites = int(chunk_sites)
        self.fd = None
        fn = str(output_spec.filename)
        if output_spec.single_timestep_files:  # LocalPropertyOutput.cc:77-93
            i = fn.find("%d")
            if i < 0:
                raise HlbError("single-timestep file names need a %d")
            self.offset_file_name = extraction_to_offset(fn[:i] + fn[i + 2:])
            self.output_file_pattern = (fn[:i], fn[i + 2:])
        else:
            self.offset_file_name = extraction_to_offset(fn)
            self.output_file_pattern = None
        self.x = _create_handle(lbm, output_spec, units)
        n, sl, hl = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(self.L.hlb_xtr_sizes(self.x, C.byref(n), C.byref(sl), C.byref(hl)))
        self.local_site_count, self.site_len, self.header_length = int(n.value), int(sl.value), int(hl.value)
        c = self.comms
        for k, v in write_layout(c, self.local_site_count, self.site_len, self.header_length).items():
            setattr(self, k, v)
        self.header_data = b""
        if c.rank == IO_RANK:
            buf = (C.c_char * self.header_length)()
            check(self.L.hlb_xtr_header(self.x, C.c_uint64(self.global_site_count), buf, C.c_uint64(self.header_length)))
            self.header_data = bytes(buf)
        m = C.c_uint32()
        check(self.L.hlb_xtr_required_caches(self.x, C.byref(m)))
        self.required_caches = int(m.value)
        self._write_offset_file()
        if not output_spec.single_timestep_files:
            self._start_file(fn)
        self.encode_kernel_ms = 0.0

    # ---- file plumbing (net::MpiFile in the reference) -------------------------------------------
    def _open_collective(self, path):
        """MPI_MODE_WRONLY | MPI_MODE_CREATE | MPI_MODE_EXCL, collectively."""
        c = self.comms
        err = None
        if c.rank == IO_RANK:
            try:
                os.close(os.open(path, os.O_WRONLY | os.O_CREAT | os.O_EXCL, 0o644))
            except OSError as e:
                err = str(e)
        err = c.broadcast(err, IO_RANK)
        if err:
            raise HlbError("cannot create %s: %s" % (path, err))
        return os.open(path, os.O_WRONLY)

    def _start_file(self, fn):  # LocalPropertyOutput::StartFile, :226-238
        self.fd = self._open_collective(fn)
        if self.comms.rank == IO_RANK:
            os.pwrite(self.fd, self.header_data, 0)

    def _write_offset_file(self):  # LocalPropertyOutput::WriteOffsetFile, :369-396
        c = self.comms
        fd = self._open_collective(self.offset_file_name)
        try:
            if c.rank == IO_RANK:
                os.pwrite(fd, struct.pack(">IIIi", HEMELB_MAGIC, OFFSET_MAGIC, OFFSET_VERSION, c.size), 0)
            at = c.rank * 8 + OFFSET_HEADER_LENGTH
            os.pwrite(fd, struct.pack(">Q", self.local_write_start), at)
            if c.rank == c.size - 1:
                os.pwrite(fd, struct.pack(">Q", self.local_write_start + self.local_data_write_length), at + 8)
        finally:
            os.close(fd)
        c.barrier()

    # ---- LocalPropertyOutput interface -----------------------------------------------------------
    def should_write(self, timestep: int) -> bool:
        return timestep % self.spec.frequency == 0

    def get_output_spec(self):
        return self.spec

    def write(self, timestep: int, total_steps: int):
        if not self.should_write(timestep):
            return
        if self.spec.single_timestep_files:  # :250-258
            prec, nxt = 3, 1000
            while total_steps > nxt:
                prec += 1
                nxt *= 10
            self._start_file("%s%*d%s" % (self.output_file_pattern[0], prec, timestep, self.output_file_pattern[1]))
        if self.local_data_write_length > 0:
            at = self.local_write_start
            if self.comms.rank == IO_RANK:
                os.pwrite(self.fd, struct.pack(">Q", timestep), at)
                at += 8
            self.encode_kernel_ms = 0.0
            cap = max(1, min(self.chunk_sites, self.local_site_count)) * self.site_len
            p = C.c_void_p()
            check(self.L.hlb_xtr_pinned_buffer(self.x, C.c_uint64(cap), C.byref(p)))
            view = (C.c_char * cap).from_address(p.value)
            ms = C.c_float()
            for s0 in range(0, self.local_site_count, self.chunk_sites):
                n = min(self.chunk_sites, self.local_site_count - s0)
                check(self.L.hlb_xtr_encode(self.x, C.c_uint64(s0), C.c_uint64(n), p, C.c_uint64(cap)))
                check(self.L.hlb_xtr_last_encode_ms(self.x, C.byref(ms)))
                self.encode_kernel_ms += ms.value
                os.pwrite(self.fd, memoryview(view)[:n * self.site_len], at + s0 * self.site_len)
        if self.spec.single_timestep_files:
            os.close(self.fd)
            self.fd = None
        else:
            self.local_write_start += self.global_data_write_length  # :356-360

    def encode(self) -> bytes:
        """This rank's record bytes of the current state (no file)."""
        out = np.zeros(max(1, self.local_site_count * self.site_len), np.uint8)
        check(self.L.hlb_xtr_encode(self.x, C.c_uint64(0), C.c_uint64(self.local_site_count), ptr(out, C.c_uint8),
                                    C.c_uint64(out.size)))
        return out[:self.local_site_count * self.site_len].tobytes()

    def close(self):
        if self.fd is not None:
            os.close(self.fd)
            self.fd = None
        if self.x:
            self.L.hlb_xtr_destroy(self.x)
            self.x = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GpuPropertyActor:
    """extraction::PropertyActor over a list of outputs (PropertyWriter): before a step, ask the
    engine to refresh the caches the outputs written this iteration need; after it, write."""

    def __init__(self, lbm, property_outputs, total_steps: int):
        self.lbm, self.outputs, self.total_steps = lbm, list(property_outputs), total_steps

    def set_required_properties(self):  # PropertyActor.cc:22-75 (flags were reset by SimulationMaster)
        mask = 0
        for o in self.outputs:
            if o.should_write(self.lbm.state.time_step):
                mask |= o.required_caches
        self.lbm.set_cache_mask(mask)
        return mask

    def end_iteration(self, timestep=None):  # PropertyActor.cc:78-83; call before SimulationState::Increment
        t = self.lbm.state.time_step if timestep is None else timestep
        for o in self.outputs:
            o.write(t, self.total_steps)


class GpuLocalDistributionInput:
    """extraction::LocalDistributionInput: read a checkpoint (a distributions-only double
    extraction file + its offset file) into the engine's f_old and f_new."""

    EXPECTED_FIELD_HEADER_LENGTH = 32

    def __init__(self, data_file_path, maybe_offset_path=None, comms=None):
        self.comms = comms or SingleComm()
        self.file_path = str(data_file_path)
        self.offset_path = str(maybe_offset_path) if maybe_offset_path else extraction_to_offset(self.file_path)
        self.timestep = None

    def _read_extraction_headers(self, fh, Q):  # LocalDistributionInput.cc:167-250
        pre = fh.read(MAIN_HEADER_LENGTH)
        magic, xmagic, version = struct.unpack(">III", pre[:12])
        if magic != HEMELB_MAGIC:
            raise HlbError("This file does not start with the HemeLB magic number. Expected: %d Actual: %d" % (HEMELB_MAGIC, magic))
        if xmagic != XTR_MAGIC:
            raise HlbError("This file does not have the extraction magic number. Expected: %d Actual: %d" % (XTR_MAGIC, xmagic))
        if version != XTR_VERSION:
            raise HlbError("Version number incorrect. Supported: %d Input: %d" % (XTR_VERSION, version))
        _, nfields, fhlen = struct.unpack(">QII", pre[44:60])
        if nfields != 1:
            raise HlbError("Checkpoint file must contain exactly one field, the distributions, but has %d" % nfields)
        if fhlen != self.EXPECTED_FIELD_HEADER_LENGTH:
            raise HlbError("Checkpoint file's field header must be %d B long, but is %d B" % (self.EXPECTED_FIELD_HEADER_LENGTH, fhlen))
        fhb = fh.read(fhlen)
        nlen = struct.unpack(">I", fhb[:4])[0]
        name = fhb[4:4 + nlen].decode()
        o = 4 + nlen + (4 - nlen % 4) % 4
        nel, tc, noff = struct.unpack(">III", fhb[o:o + 12])
        if name != "distributions":
            raise HlbError("Checkpoint file must contain field named 'distributions', but has '%s'" % name)
        if nel != Q:
            raise HlbError("Checkpoint field distributions contains %d distributions but this build of HemeLB requires %d" % (nel, Q))
        if tc != TYPECODES["double"]:
            raise HlbError("Checkpoint contains wrong data type")
        if noff != 0:
            raise HlbError("Checkpoint should not have offsets")

    def _read_offsets(self):  # LocalDistributionInput.cc:252-308
        c = self.comms
        pairs, all_len = None, None
        if c.rank == IO_RANK:
            with open(self.offset_path, "rb") as fh:
                b = fh.read()
            magic, omagic, version, nranks = struct.unpack(">IIIi", b[:16])
            if magic != HEMELB_MAGIC:
                raise HlbError("This file does not start with the HemeLB magic number. Expected: %d Actual: %d" % (HEMELB_MAGIC, magic))
            if omagic != OFFSET_MAGIC:
                raise HlbError("This file does not have the offset magic number. Expected: %d Actual: %d" % (OFFSET_MAGIC, omagic))
            if version != OFFSET_VERSION:
                raise HlbError("Version number incorrect. Supported: %d Input: %d" % (OFFSET_VERSION, version))
            if nranks != c.size:
                raise HlbError("Offset file has wrong number of MPI ranks. Running with: %d Input: %d" % (c.size, nranks))
            offs = struct.unpack(">%dQ" % (nranks + 1), b[16:16 + 8 * (nranks + 1)])
            pairs = [(offs[r], offs[r + 1]) for r in range(nranks)]
            all_len = offs[nranks] - offs[0]
        all_len = c.broadcast(all_len, IO_RANK)
        start, stop = c.scatter(pairs, IO_RANK)
        return all_len, start, stop

    def load_distribution(self, lbm, target_time=None):
        """Returns the time step read (the last one in the file when target_time is None)."""
        c = self.comms
        L = lib()
        Q = lbm.Q
        total_header = MAIN_HEADER_LENGTH + self.EXPECTED_FIELD_HEADER_LENGTH
        with open(self.file_path, "rb") as fh:
            if c.rank == IO_RANK:
                self._read_extraction_headers(fh, Q)
            all_len, local_start, local_stop = self._read_offsets()
            data_size = os.fstat(fh.fileno()).st_size - total_header
            if data_size % all_len:
                raise HlbError("Checkpoint file length not consistent with integer number of checkpoints")
            n_times = data_size // all_len

            def read_time_by_index(i):
                return struct.unpack(">Q", os.pread(fh.fileno(), 8, local_start + i * all_len))[0]

            its = timestep = None
            if c.rank == IO_RANK:
                if target_time is not None:
                    # the reference's search, literally (LocalDistributionInput.cc:79-96): `timestep`
                    # ends as the LAST PROBED record, so some present targets are reported missing
                    its, length = 0, n_times
                    while length != 0:
                        l2 = length // 2
                        m = its + l2
                        timestep = read_time_by_index(m)
                        if timestep < target_time:
                            its = m + 1
                            length -= l2 + 1
                        else:
                            length = l2
                    if timestep != target_time:
                        raise HlbError("Target timestep %d not found in checkpoint file." % target_time)
                else:
                    its = n_times - 1
                    timestep = read_time_by_index(its)
            timestep = c.broadcast(timestep, IO_RANK)
            its = c.broadcast(its, IO_RANK)
            chunk = os.pread(fh.fileno(), local_stop - local_start, its * all_len + local_start)
        if c.rank == IO_RANK:
            chunk = chunk[8:]
        buf = np.frombuffer(chunk, np.uint8)
        dom = lbm.domain
        if hasattr(dom, "d"):
            check(L.hlb_gpu_load_distributions_from_domain(lbm.h, dom.d, ptr(buf, C.c_uint8) if buf.size else None,
                                                           C.c_uint64(buf.size)))
        else:
            coords = np.ascontiguousarray(dom.globalCoords, np.int64)
            check(L.hlb_gpu_load_distributions(lbm.h, ptr(buf, C.c_uint8) if buf.size else None, C.c_uint64(buf.size),
                                               ptr(coords, C.c_int64)))
        self.timestep = timestep
        return timestep
