"""The C++ host-side drop-in (hemelb_b200/host): Gpu*Streamer classes satisfy the reference's
lb::streamer concept and its wall+iolet combination trait.  Compile-only; needs the reference
headers, so it runs where /root/reference exists."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/Code"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout absent")
def test_gpu_streamers_satisfy_reference_concepts():
    cmd = ["g++", "-std=c++20", "-fsyntax-only", "-w", "-I" + os.path.join(ROOT, "hemelb_b200", "host"),
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "tests", "host_shim"),
           "-I" + os.path.join(ROOT, "oracle", "ref_shim"), "-I" + REF,
           os.path.join(ROOT, "tests", "host_concept_check.cc")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout absent")
def test_gpu_property_encoder_builds_against_reference_extraction_headers():
    cmd = ["g++", "-std=c++20", "-fsyntax-only", "-w", "-I" + os.path.join(ROOT, "hemelb_b200", "host"),
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "tests", "host_shim"),
           "-I" + os.path.join(ROOT, "oracle", "ref_shim"), "-I" + REF,
           os.path.join(ROOT, "tests", "host_extraction_check.cc")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


def test_c_header_is_plain_c():
    """include/hemelb_b200.h must be consumable from C (cgo/JNI/ctypes style binding)."""
    src = '#include "hemelb_b200.h"\nint main(void){ hlb_gpu_config c; (void)c; return 0; }\n'
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), "-x", "c", "-"],
                       input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
