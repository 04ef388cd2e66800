"""Builds the runnable C++ host-side check (tests/host_lbm_run.cc) into tests/_build/:

  host_lbm_run_mock   linked against the recording C-ABI stand-in (tests/host_mock_abi.cc), CPU
  host_lbm_run        linked against hemelb_b200/libhemelb_b200.so, GPU

Both compile the reference's own LbmParameters / SimulationState / InOutLet / SiteData /
MacroscopicPropertyCache sources where they lie under /root/reference (never copied), so they can
only be built where the reference exists; the binaries travel to the GPU box with the snapshot
(tests/_build/ is git-ignored, not gpurun-ignored).  Test infrastructure only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")
REF = "/root/reference/Code"
REF_SRCS = ["net/IteratedAction.cc", "geometry/neighbouring/RequiredSiteInformation.cc",
            "lb/MacroscopicPropertyCache.cc", "lb/SimulationState.cc", "geometry/SiteDataBare.cc", "util/Matrix3D.cc",
            "util/Vector3D.cc", "lb/kernels/DHumieresD3Q19MRTBasis.cc", "lb/iolets/InOutLet.cc",
            "lb/iolets/InOutLetCosine.cc", "lb/iolets/InOutLetVelocity.cc", "lb/iolets/InOutLetParabolicVelocity.cc"]
# the real lb::LBM over the real geometry::Domain, net::Net and lb::BoundaryValues, ranks as threads
# (tests/host_lbm_real.cc): every reference source it links, unmodified
REAL_LBM_REF_SRCS = [
    "geometry/Domain.cc", "geometry/LookupTree.cc", "geometry/GmyReadResult.cc", "geometry/Block.cc",
    "geometry/BlockTraverser.cc", "geometry/SiteTraverser.cc", "geometry/VolumeTraverser.cc", "geometry/SiteDataBare.cc",
    "geometry/neighbouring/NeighbouringDomain.cc", "geometry/neighbouring/RequiredSiteInformation.cc",
    "geometry/decomposition/BasicDecomposition.cc",
    "net/MpiCommunicator.cc", "net/MpiGroup.cc", "net/IOCommunicator.cc", "net/MpiError.cc", "net/BaseNet.cc",
    "net/IteratedAction.cc", "net/mixins/StoringNet.cc", "net/mixins/pointpoint/SeparatedPointPoint.cc",
    "net/mixins/alltoall/SeparatedAllToAll.cc", "net/mixins/gathers/SeparatedGathers.cc",
    "util/Vector3D.cc", "util/Vector3DHemeLb.cc", "util/UnitConverter.cc", "util/utilityFunctions.cc", "util/Matrix3D.cc",
    "lb/iolets/BoundaryValues.cc", "lb/iolets/BoundaryComms.cc", "lb/iolets/BoundaryCommunicator.cc",
    "lb/iolets/InOutLet.cc", "lb/iolets/InOutLetCosine.cc", "lb/iolets/InOutLetVelocity.cc",
    "lb/iolets/InOutLetParabolicVelocity.cc", "lb/SimulationState.cc", "lb/MacroscopicPropertyCache.cc",
    "reporting/Timers.cc"]
XTR_REF_SRCS = ["util/Vector3D.cc", "extraction/GeometrySelector.cc", "extraction/WholeGeometrySelector.cc",
                "extraction/GeometrySurfaceSelector.cc", "extraction/PlaneGeometrySelector.cc",
                "extraction/StraightLineGeometrySelector.cc", "extraction/SurfacePointSelector.cc"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def build_host_binaries(verbose=False):
    """No-op when the reference is absent (prebuilt binaries are used) or everything is up to date."""
    if not os.path.isdir(REF):
        return False
    os.makedirs(BUILD, exist_ok=True)
    host = os.path.join(ROOT, "hemelb_b200", "host")
    deps = [os.path.join(ROOT, "tests", "host_lbm_run.cc"), os.path.join(ROOT, "include", "hemelb_b200.h"),
            os.path.join(host, "geometry", "FieldData.h"), os.path.join(host, "lb", "streamers", "GpuStreamers.h"),
            os.path.join(host, "lb", "StabilityTester.h"), os.path.join(host, "lb", "IncompressibilityChecker.h"),
            os.path.join(host, "geometry", "neighbouring", "NeighbouringDataManager.h"),
            os.path.join(ROOT, "tests", "host_shim", "net", "mixins", "InterfaceDelegationNet.h"), os.path.join(ROOT, "tests", "host_shim", "net", "PhasedBroadcastRegular.h"),
            os.path.join(ROOT, "tests", "host_shim", "reporting", "Timers.h"),
            os.path.join(ROOT, "tests", "host_shim", "geometry", "Domain.h"), os.path.abspath(__file__)]
    common = ["g++", "-std=c++20", "-O1", "-w", "-I" + host, "-I" + os.path.join(ROOT, "include"),
              "-I" + os.path.join(ROOT, "tests", "host_shim"), "-I" + os.path.join(ROOT, "oracle", "ref_shim"), "-I" + REF]
    obj = os.path.join(BUILD, "host_lbm_run.o")
    refobj = os.path.join(BUILD, "host_ref_objs.o")
    if _stale(obj, deps):
        subprocess.run(common + ["-c", deps[0], "-o", obj], check=True)
    if _stale(refobj, [os.path.abspath(__file__)]):
        objs = []
        for i, s in enumerate(REF_SRCS):
            o = os.path.join(BUILD, "ref_%d.o" % i)
            subprocess.run(common + ["-c", os.path.join(REF, s), "-o", o], check=True)
            objs.append(o)
        subprocess.run(["ld", "-r", "-o", refobj] + objs, check=True)
        for o in objs:
            os.remove(o)
    mock = os.path.join(BUILD, "host_lbm_run_mock")
    mock_src = os.path.join(ROOT, "tests", "host_mock_abi.cc")
    if _stale(mock, [obj, refobj, mock_src]):
        subprocess.run(common + [obj, refobj, mock_src, "-o", mock], check=True)
    lib = os.path.join(ROOT, "hemelb_b200", "libhemelb_b200.so")
    real = os.path.join(BUILD, "host_lbm_run")
    if os.path.exists(lib) and _stale(real, [obj, refobj, lib]):
        # rpath relative to the binary: the snapshot is unpacked at another path on the GPU box
        subprocess.run(["g++", obj, refobj, "-L" + os.path.dirname(lib), "-lhemelb_b200",
                        "-Wl,-rpath,$ORIGIN/../../hemelb_b200", "-Wl,--allow-shlib-undefined", "-o", real], check=True)
    # the extraction face (extraction/GpuPropertyEncoder.h) against the recording ABI
    xtr = os.path.join(BUILD, "host_xtr_run_mock")
    xtr_src = os.path.join(ROOT, "tests", "host_xtr_run.cc")
    if _stale(xtr, [xtr_src, mock_src, os.path.join(host, "extraction", "GpuPropertyEncoder.h"), deps[1], os.path.abspath(__file__)]):
        subprocess.run(common + [xtr_src, mock_src] + [os.path.join(REF, s) for s in XTR_REF_SRCS] + ["-o", xtr], check=True)
    # the reference's own lb::LBM driving the GPU policy classes (recording ABI, emulated ranks)
    real_lbm = os.path.join(BUILD, "libhost_lbm_real.so")
    real_src = os.path.join(ROOT, "tests", "host_lbm_real.cc")
    shim_lbm = os.path.join(ROOT, "tests", "host_shim_lbm")
    oracle = os.path.join(ROOT, "oracle")
    real_deps = [real_src, mock_src, os.path.join(oracle, "fake_mpi.cc"), os.path.join(oracle, "ref_domain_build.h"),
                 os.path.join(shim_lbm, "Traits.h"), os.path.join(shim_lbm, "build_info.h"), deps[1], deps[2], deps[3],
                 os.path.abspath(__file__)]
    if _stale(real_lbm, real_deps):
        # (-Bsymbolic: oracle/_ref/libhemelb_refdom.so, which a test process may have loaded before, defines the
        # same stand-in MPI and reference symbols with a state of its own)
        subprocess.run(["g++", "-std=c++20", "-O1", "-w", "-fPIC", "-shared", "-pthread", "-Wl,--no-undefined",
                        "-Wl,-Bsymbolic", "-I" + host, "-I" + os.path.join(ROOT, "include"), "-I" + shim_lbm,
                        "-I" + os.path.join(oracle, "ref_shim_dom"), "-I" + oracle, "-I" + REF, "-o", real_lbm,
                        real_src, mock_src, os.path.join(oracle, "fake_mpi.cc")]
                       + [os.path.join(REF, s_) for s_ in REAL_LBM_REF_SRCS], check=True)
    # the same against the product library (GPU, one rank): results for the oracle
    real_gpu = os.path.join(BUILD, "libhost_lbm_real_gpu.so")
    if os.path.exists(lib) and _stale(real_gpu, real_deps + [lib]):
        subprocess.run(["g++", "-std=c++20", "-O1", "-w", "-fPIC", "-shared", "-pthread", "-DHLB_REAL_ENGINE",
                        "-Wl,-Bsymbolic", "-I" + host, "-I" + os.path.join(ROOT, "include"), "-I" + shim_lbm,
                        "-I" + os.path.join(oracle, "ref_shim_dom"), "-I" + oracle, "-I" + REF, "-o", real_gpu,
                        real_src, os.path.join(oracle, "fake_mpi.cc")]
                       + [os.path.join(REF, s_) for s_ in REAL_LBM_REF_SRCS]
                       + ["-L" + os.path.dirname(lib), "-lhemelb_b200", "-Wl,-rpath,$ORIGIN/../../hemelb_b200",
                          "-Wl,--allow-shlib-undefined"], check=True)
    if verbose:
        print("host binaries in", BUILD)
    return True


if __name__ == "__main__":
    build_host_binaries(True)
