"""The reference-side binding INTEGRATION.md advertises (include/spaND_b200.hpp) is compiled here with plain
g++ -std=c++14 against the C-ABI library, exactly as a maintainer of the reference would, and run."""
import os
import shutil
import subprocess

import pytest

import spand_public_b200 as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cxx_facade_compiles_and_runs(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "facade_smoke")
    libdir = os.path.dirname(S.LIB_PATH)
    subprocess.check_call([gxx, "-std=c++14", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cxx", "facade_smoke.cpp"), "-o", exe, "-L", libdir,
                           "-lspand_b200", f"-Wl,-rpath,{libdir}"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "FACADE_OK top_separator_dofs=5" in r.stdout  # PartitionTest.Square: the middle line of the 5 x 5 grid
    assert "FACADE_GPU" in r.stdout or "FACADE_NOGPU" in r.stdout
