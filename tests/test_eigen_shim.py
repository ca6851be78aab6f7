"""oracle/eigen_shim (the Eigen stand-in the reference's liquid solver is compiled over) keeps the
evaluation semantics its header promises: coefficient-wise left-to-right sums, dense x sparse summed over the sorted
column from zero, duplicate triplets added, diagonal and array products.  The inputs make the ORDER of the
floating-point operations visible in the result."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_eigen_shim_semantics(tmp_path):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not cxx:
        pytest.skip("no C++ compiler")
    exe = str(tmp_path / "selftest")
    shim = os.path.join(ROOT, "oracle", "eigen_shim")
    r = subprocess.run([cxx, "-std=c++20", "-O1", "-ffp-contract=off", "-fno-fast-math", "-I", shim, os.path.join(shim, "selftest.cpp"), "-o", exe],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "selftest ok" in r.stdout, r.stdout + r.stderr
