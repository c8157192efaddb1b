"""Builds tests/_build/libhostgeom.so: the product's pair-geometry header compiled for the HOST (test infrastructure,
see host_geom.cu).  Used by tests/test_host_geom.py and by __graft_entry__.build()."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_geom.cu")
HDR = os.path.join(HERE, "..", "..", "jdet_b200", "csrc", "rbox_geom.cuh")
SO = os.path.join(HERE, "..", "_build", "libhostgeom.so")


def build(force=False):
    """-> path of the library, or None when nvcc is not on PATH."""
    if shutil.which("nvcc") is None:
        return None
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off",
                               "-gencode", "arch=compute_100a,code=sm_100a", SRC, "-o", SO])
    return SO
