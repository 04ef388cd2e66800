"""Copies the geometry / configuration fixtures that the reference's own tests hold
(Code/tests/resources) into tests/golden/ref_inputs/, so that the GPU box -- which has no reference
checkout -- runs the parity tests on reference-held inputs with real cut distances and wall normals.
Data files only (no source).  Run once where /root/reference exists; the copies are committed."""
import os
import shutil

SRC = "/root/reference/Code/tests/resources"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_inputs")
NAMES = ["four_cube", "large_cylinder", "fedosov1c", "cyl_l100_r5"]

if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    for n in NAMES:
        for ext in (".gmy", ".xml"):
            shutil.copyfile(os.path.join(SRC, n + ext), os.path.join(DST, n + ext))
            print("copied", n + ext, os.path.getsize(os.path.join(DST, n + ext)), "bytes")
