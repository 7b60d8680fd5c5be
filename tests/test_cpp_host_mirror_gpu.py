"""Runs the C++ host mirror's test program (tests/cpp/host_mirror_test.cpp over
phylogenetics_b200/host/phylo_b200.hpp) on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "host_mirror_test")


def build_cpp_test():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "phylogenetics_b200", "host"),
                    os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "-o", EXE,
                    "-L" + os.path.join(ROOT, "phylogenetics_b200"), "-lphylo_b200",
                    "-Wl,-rpath,$ORIGIN/../../../phylogenetics_b200"], check=True)
    return EXE


def test_cpp_host_mirror_compiles():
    """CPU-side: the header and its test program compile and link against the C ABI."""
    build_cpp_test()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_host_mirror_runs():
    if not os.path.exists(EXE):
        build_cpp_test()
    out = subprocess.run([EXE], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all ok" in out.stdout and "FAIL" not in out.stdout
