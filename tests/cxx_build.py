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


REPLACED = ("lapl_cube.h", "lapl_rect.h", "lapl_cyl.h", "ns_cube.h", "ns_cyl.h")


def make_overlay(dst, native_plotter=False):
    """The reference's src/ with the five class headers REPLACED by the drop-in ones (what a maintainer does; an extra
    include directory is not enough, because a quoted #include from a reference header looks in its own directory
    first).  Everything else, and test/*.cpp, are symlinks to the untouched reference files.
    native_plotter: velocity_plot.h is replaced too (header-only device-side plotter; src/velocity_plot.cpp is then
    dropped from the build)."""
    ref = "/root/reference"
    os.makedirs(os.path.join(dst, "src"), exist_ok=True)
    os.makedirs(os.path.join(dst, "test"), exist_ok=True)
    for name in os.listdir(os.path.join(ref, "src")):
        tgt = os.path.join(dst, "src", name)
        if name in REPLACED or (native_plotter and name == "velocity_plot.h"):
            src = os.path.join(ROOT, "fdm_b200", "cxx", name)
        else:
            src = os.path.join(ref, "src", name)
        if not os.path.lexists(tgt):
            os.symlink(src, tgt)
    for name in os.listdir(os.path.join(ref, "test")):
        tgt = os.path.join(dst, "test", name)
        if not os.path.lexists(tgt):
            os.symlink(os.path.join(ref, "test", name), tgt)
    return dst


def compile_in_overlay(overlay, rel, out):
    """Compile an UNMODIFIED reference translation unit (path relative to the overlay) against the drop-in headers."""
    cmd = ["/usr/bin/g++", "-std=c++20", "-O1", "-c", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(overlay, "src"), "-I" + os.path.join(ROOT, "oracle", "stub"), "-include", "cmath",
           os.path.join(overlay, rel), "-o", out]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return out


DRIVER = os.path.join(ROOT, "tests", "cxx", "_build", "fdm_ns_cube")


DRIVER_NATIVE_PLOT = os.path.join(ROOT, "tests", "cxx", "_build", "fdm_ns_cube_native_plot")


def build_reference_driver_native_plot(overlay, out=DRIVER_NATIVE_PLOT):
    """The same unmodified driver source with velocity_plot.h replaced as well (overlay made with native_plotter=True):
    no src/velocity_plot.cpp, no plplot symbols -- everything resolves."""
    os.makedirs(os.path.dirname(out), exist_ok=True)
    srcs = [os.path.join(overlay, p) for p in ("test/test_ns_cube.cpp", "src/config.cpp", "src/asp_misc.cpp")]
    cmd = ["/usr/bin/g++", "-std=c++20", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(overlay, "src"),
           "-I" + os.path.join(ROOT, "oracle", "stub"), "-include", "cmath", *srcs, "-o", out,
           "-L" + LIBDIR, "-lfdm_b200", "-Wl,-rpath,$ORIGIN/../../../fdm_b200"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return out


def build_reference_driver(overlay, out=DRIVER):
    """The reference's OWN driver executable fdm_ns_cube (test/test_ns_cube.cpp + src/velocity_plot.cpp + src/config.cpp
    + src/asp_misc.cpp, all unmodified) built against the replaced class headers and linked with libfdm_b200.so.
    plot() needs plplot, which is absent: its symbols stay unresolved and the driver is run with --plot:png=0."""
    os.makedirs(os.path.dirname(out), exist_ok=True)
    srcs = [os.path.join(overlay, p) for p in ("test/test_ns_cube.cpp", "src/velocity_plot.cpp", "src/config.cpp",
                                               "src/asp_misc.cpp")]
    cmd = ["/usr/bin/g++", "-std=c++20", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(overlay, "src"),
           "-I" + os.path.join(ROOT, "oracle", "stub"), "-include", "cmath", *srcs, "-o", out,
           "-L" + LIBDIR, "-lfdm_b200", "-Wl,-rpath,$ORIGIN/../../../fdm_b200", "-Wl,--unresolved-symbols=ignore-all"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return out


def build_example(name, out):
    """examples/<name>.cpp: our own thin drivers with the reference drivers' command lines, against the compat headers."""
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "fdm_b200", "cxx"), os.path.join(ROOT, "examples", name + ".cpp"), "-o", out,
           "-L" + LIBDIR, "-lfdm_b200", "-Wl,-rpath," + LIBDIR]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return out
