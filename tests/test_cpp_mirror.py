"""The C++ host mirror (include/scann_b200.hpp) compiles against the C ABI (CPU check) and passes the
reference's unit-test sequences on a GPU (-m gpu)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_hpp_mirror.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_hpp_mirror.bin")


def _build(pkg):
    lib = pkg.build_lib.build()
    libdir = os.path.dirname(lib)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE, "-L", libdir,
           "-lscann_b200", f"-Wl,-rpath,{libdir}"]
    subprocess.check_call(cmd)
    return EXE


def test_cpp_mirror_compiles_and_links(pkg):
    assert os.path.exists(_build(pkg))


def test_c_header_is_plain_c(pkg):
    # the ABI header must be consumable from C (no C++-isms)
    probe = os.path.join(ROOT, "tests", "cpp", "_probe.c")
    with open(probe, "w") as f:
        f.write('#include "scann_b200.h"\nint main(void){ return scann_version() > 0 ? 0 : 1; }\n')
    try:
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", probe,
                               "-o", probe + ".o"])
    finally:
        for p in (probe, probe + ".o"):
            if os.path.exists(p):
                os.remove(p)


@pytest.mark.gpu
def test_cpp_mirror_runs_reference_unit_tests(gpu_lib):
    exe = _build(gpu_lib)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "hpp mirror ok" in out.stdout
