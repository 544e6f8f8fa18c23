"""Static checks of the built library (no GPU): no kernel spills registers, the sweep kernels really use the TMA /
mbarrier path (UTMALDG, UBLKCP, SYNCS in the SASS), the PM deposit uses fp64 reductions, and nothing on this path went
to the tensor cores (none of these steps is a dense contraction).  See profiles/r01k_static.md."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "fdm_b200", "csrc")
LIB = os.path.join(ROOT, "fdm_b200", "libfdm_b200.so")


def test_no_register_spills():
    logs = [f for f in os.listdir(CSRC) if f.endswith(".ptxas.log")]
    if not logs:
        pytest.skip("ptxas logs not present (library built elsewhere)")
    kernels = 0
    for log in logs:
        text = open(os.path.join(CSRC, log)).read()
        kernels += len(re.findall(r"Compiling entry function", text))
        for m in re.finditer(r"(\d+) bytes spill stores, (\d+) bytes spill loads", text):
            assert m.group(1) == "0" and m.group(2) == "0", (log, m.group(0))
    assert kernels >= 100          # every transform length x kind x sweep flavour is instantiated


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_sass_uses_tma_and_no_tensor_cores():
    if not os.path.exists(LIB):
        pytest.skip("library not built")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    assert "sm_100a" in sass or "sm_100" in sass
    count = lambda mn: len(re.findall(r"\b" + mn + r"\b", sass))      # noqa: E731
    assert count("UTMALDG") > 100        # tensor-map tile loads of the column sweeps
    assert count("UBLKCP") > 10          # 1-D bulk copies of the row sweeps
    assert count("SYNCS") > 100          # mbarrier traffic around them
    assert len(re.findall(r"REDG\.E\.ADD\.F64", sass)) >= 8       # the PM deposit's fp64 atomics
    assert count("DFMA") > 10000
    for mma in ("HMMA", "IMMA", "DMMA", "UTCHMMA", "UTCQMMA"):
        assert count(mma) == 0, mma


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_sass_dependent_launch_and_marching_fgh():
    """Programmatic dependent launch is compiled into every sweep / NS kernel (griddepcontrol.wait = ACQBULK,
    launch_dependents = PREEXIT), and the fused FGH + divergence sweep is the barrier-free register-marching kernel
    DESIGN 4.3 describes: warp shuffles for F[j-1], L2 prefetches, no CTA barrier, no shared memory."""
    if not os.path.exists(LIB):
        pytest.skip("library not built")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    count = lambda mn, text=sass: len(re.findall(r"\b" + mn + r"\b", text))      # noqa: E731
    assert count("ACQBULK") >= 100 and count("PREEXIT") >= 100
    # the body of k_fgh_div: from its "Function :" header to the next one
    m = re.search(r"Function : (\S*k_fgh_div\S*)(.*?)(?=\n\s*Function : |\Z)", sass, re.S)
    assert m, "k_fgh_div not in the library"
    body = m.group(2)
    assert count("ACQBULK", body) == 1 and count("PREEXIT", body) == 1
    assert len(re.findall(r"SHFL\.UP", body)) >= 2            # F[j-1] from the lane to the left (two 32-bit halves)
    assert len(re.findall(r"CCTL\.E\.PF2", body)) >= 3        # prefetch.global.L2 of the first-touch rows
    assert len(re.findall(r"BAR\.SYNC", body)) == 0 and count("LDS", body) == 0 and count("STS", body) == 0
    assert count("DFMA", body) > 100
