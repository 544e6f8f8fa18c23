"""Helper: builds tests/cxx/shim_check.cpp against the drop-in C++ headers (fdm_b200/cxx)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cxx", "shim_check.cpp")
LIBDIR = os.path.join(ROOT, "fdm_b200")


def build(out, with_reference_headers=False):
    # the drop-in class headers shadow the reference's same-named ones (lapl_cube.h, ns_cube.h, ...)
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "fdm_b200", "cxx")]
    std = "-std=c++17"
    if with_reference_headers:
        # ... while the reference's own tensor.h / config.h are picked up, like inside the reference tree
        inc += ["-I/root/reference/src", "-I" + os.path.join(ROOT, "oracle", "stub"), "-include", "cmath"]
        std = "-std=c++20"
    cmd = ["/usr/bin/g++", std, "-O1", "-Wall", *inc, SRC, "-o", out]
    if with_reference_headers:
        cmd.append("-c")         # compile only: the reference's config.cpp is not linked here
    else:
        cmd += ["-L" + LIBDIR, "-lfdm_b200", "-Wl,-rpath," + LIBDIR]
    subprocess.run(cmd, check=True)
    return out
