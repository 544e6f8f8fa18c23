"""The CUDA tile transforms (fdm_b200/csrc/xform.cuh) compiled for the host with
thread-barrier shims, checked against the oracle for every instantiated length.
This exercises the exact device algebra (fold, radix passes, digit reversal,
untangle, prefix scan) without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import fdm_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "xform_host.cpp")
LIB = os.path.join(HERE, "host_emul", "libxform_host.so")


@pytest.fixture(scope="module")
def emul():
    hdr = os.path.join(HERE, "..", "fdm_b200", "csrc", "xform.cuh")
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++20", "-shared", "-fPIC", "-pthread", SRC, "-o", LIB],
                       check=True)
    L = C.CDLL(LIB)
    L.emul_xform.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_double]
    L.emul_xform_f32.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_float]
    L.emul_dst_fused.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int, C.c_double, C.c_int, C.POINTER(C.c_double), C.c_int]
    L.emul_dst_fused_half.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_double, C.c_int, C.POINTER(C.c_double), C.c_int]
    return L


@pytest.mark.parametrize("N", [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048])
@pytest.mark.parametrize("sj", [1, 5])
def test_tile_transforms(emul, N, sj):
    rng = np.random.default_rng(N + sj)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    x = rng.uniform(-1, 1, N - 1)
    d = np.zeros(N * sj); d[sj::sj][: N - 1] = x; d[0] = 99.0     # slot 0 is scratch
    assert emul.emul_xform(0, N, p(d), sj, 0.37) == 0
    assert O.rel_l2(d[sj::sj][: N - 1], O.sFFT(x, 0.37)) < 2e-14
    x = rng.uniform(-1, 1, N)
    d = np.zeros(N * sj); d[::sj] = x
    assert emul.emul_xform(1, N, p(d), sj, 0.37) == 0
    assert O.rel_l2(d[::sj], O.pFFT_1(x, 0.37)) < 2e-15
    d = np.zeros(N * sj); d[::sj] = x
    assert emul.emul_xform(2, N, p(d), sj, 0.37) == 0
    assert O.rel_l2(d[::sj], O.pFFT(x, 0.37)) < 2e-15
    x = rng.uniform(-1, 1, N + 1)                                 # cFFT: N + 1 values (src/fft.cpp:368-445)
    d = np.zeros((N + 1) * sj); d[::sj] = x
    assert emul.emul_xform(3, N, p(d), sj, 0.37) == 0
    assert O.rel_l2(d[::sj], O.cFFT(x, 0.37)) < 2e-14


def test_roundtrip_identity(emul):
    # ut/ut_fft.cpp:191-228 -- sFFT o sFFT * (2/N) = id
    N = 256
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, N - 1)
    d = np.zeros(N); d[1:] = x
    p = d.ctypes.data_as(C.POINTER(C.c_double))
    emul.emul_xform(0, N, p, 1, 1.0)
    emul.emul_xform(0, N, p, 1, 2.0 / N)
    assert O.rel_l2(d[1:], x) < 1e-14


@pytest.mark.parametrize("N", [32, 64, 128, 256, 512, 1024, 2048])
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("swz", [0, 1])
def test_fused_dst(emul, N, mode, swz):
    """dst_tile_fused on the planar tile (fold + first pass, last pass + untangle in registers, output
    policies); swz = 1 adds the 8-column-tile swizzles (ignored for two-pass plans)."""
    rng = np.random.default_rng(7 * N + mode)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    sj = 3
    x = rng.uniform(-1, 1, N - 1)
    d = np.zeros(N * sj); d[sj::sj][: N - 1] = x; d[0] = 123.0     # slot 0 must be ignored
    sep = np.zeros(N - 1)
    assert emul.emul_dst_fused(N, p(d), sj, 0.37, mode, p(sep), swz) == 0
    got = sep if mode == 1 else d[sj::sj][: N - 1]
    assert O.rel_l2(got, O.sFFT(x, 0.37)) < 2e-14


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("swz", [0, 1])
def test_fused_dst_1024_half_threads(emul, mode, swz):
    """N = 1024 with 32 threads per sequence (PipeCfg<1024>::GC): two first/middle-pass butterflies per
    thread, one last-pass unit per thread."""
    N = 1024
    rng = np.random.default_rng(99 + mode)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    sj = 3
    x = rng.uniform(-1, 1, N - 1)
    d = np.zeros(N * sj); d[sj::sj][: N - 1] = x; d[0] = 123.0
    sep = np.zeros(N - 1)
    assert emul.emul_dst_fused_half(p(d), sj, 0.37, mode, p(sep), swz) == 0
    got = sep if mode == 1 else d[sj::sj][: N - 1]
    assert O.rel_l2(got, O.sFFT(x, 0.37)) < 2e-14


@pytest.mark.parametrize("N", [4, 8, 16, 32, 64, 128, 256, 512, 1024])
def test_tile_transforms_single_precision(emul, N):
    """The same tile transforms instantiated for float (fdm::FFT<float>, src/fft.cpp:481): single-precision accuracy
    against the double oracle, growing like sqrt(log N) as a float FFT does."""
    rng = np.random.default_rng(7 * N)
    pf = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    tol = 4e-7 * (2 + np.log2(N))
    x = rng.uniform(-1, 1, N - 1).astype(np.float32)
    d = np.zeros(N, dtype=np.float32); d[1:] = x
    assert emul.emul_xform_f32(0, N, pf(d), 1, 0.37) == 0
    assert O.rel_l2(d[1:], O.sFFT(x.astype(np.float64), 0.37)) < tol
    x = rng.uniform(-1, 1, N).astype(np.float32)
    d = x.copy()
    assert emul.emul_xform_f32(1, N, pf(d), 1, 0.37) == 0
    assert O.rel_l2(d, O.pFFT_1(x.astype(np.float64), 0.37)) < tol
    d = x.copy()
    assert emul.emul_xform_f32(2, N, pf(d), 1, 0.37) == 0
    assert O.rel_l2(d, O.pFFT(x.astype(np.float64), 0.37)) < tol
    x = rng.uniform(-1, 1, N + 1).astype(np.float32)
    d = x.copy()
    assert emul.emul_xform_f32(3, N, pf(d), 1, 0.37) == 0
    assert O.rel_l2(d, O.cFFT(x.astype(np.float64), 0.37)) < tol
