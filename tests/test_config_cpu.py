"""The stand-in Config of fdm_b200/cxx/fdm_compat_config.h (used by the drop-in headers and the example drivers
outside the reference tree) reads INI files and --section:key=value overrides exactly like the reference's Config
(src/config.cpp:62-150,255-297): the same program is built against both and their outputs compared."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cxx", "config_check.cpp")

INI = """early = 5
# a comment line
; another one
[ns]
nx = 31
nz=63
Re = 250.5 ; inline comment
dt\t=\t0.01
tabbed\t1.5
spaced   =   12   trailing tokens are dropped
x1 = -3.25
label = lid-driven

[plot]
interval = 10
vtk = 1
[solver]
datatype = float
[st]
input = input.nc ; eigenvectors
[pre]
early = 9
"""


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")
@pytest.mark.parametrize("argv", [[], ["--ns:nx=127", "--ns:Re=1e3", "--other:check=1", "--solver:datatype=double",
                                       "--bad", "-ns:nz=1", "--ns:steps=40", "--nosep", "--ns:label=a=b"]])
def test_compat_config_matches_the_reference(tmp_path, argv):
    ini = tmp_path / "run.ini"
    ini.write_text(INI)
    ref_exe, our_exe = str(tmp_path / "cfg_ref"), str(tmp_path / "cfg_ours")
    subprocess.run(["/usr/bin/g++", "-std=c++20", "-O1", "-DUSE_REFERENCE_CONFIG", "-I/root/reference/src", SRC,
                    "/root/reference/src/config.cpp", "-o", ref_exe], check=True, capture_output=True)
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "fdm_b200", "cxx"), SRC, "-o", our_exe],
                   check=True, capture_output=True)
    a = subprocess.run([ref_exe, str(ini), *argv], capture_output=True, text=True, check=True).stdout
    b = subprocess.run([our_exe, str(ini), *argv], capture_output=True, text=True, check=True).stdout
    assert a == b and "int ns:nx" in a
    # and a missing file is silently fine in both
    a = subprocess.run([ref_exe, str(tmp_path / "absent.ini"), *argv], capture_output=True, text=True, check=True).stdout
    b = subprocess.run([our_exe, str(tmp_path / "absent.ini"), *argv], capture_output=True, text=True, check=True).stdout
    assert a == b
