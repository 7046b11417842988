"""The C++17 drop-in headers (include/pcpx/pcp.hpp): they compile with strict warnings on a
CPU-only box, and — on a GPU — the reference's own test scenarios, written with the reference's
call signatures, hold when run through libpcpx.so (tests/cpp/dropin_test.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "dropin_test.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "dropin_test")
LIBDIR = os.path.join(ROOT, "point-cloud-processing_b200", "lib")


def compile_dropin(pcpx):
    pcpx.lib()  # the library must exist to link against
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror",
                    "-I", os.path.join(ROOT, "include"), SRC, "-L", LIBDIR, "-lpcpx",
                    "-Wl,-rpath," + LIBDIR, "-o", EXE], check=True)


def test_dropin_headers_compile(pcpx):
    compile_dropin(pcpx)
    if pcpx.device_count() == 0:
        r = subprocess.run([EXE], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU path" in r.stderr  # fails loudly without a GPU


@pytest.mark.gpu
def test_reference_scenarios_through_the_dropin(pcpx):
    compile_dropin(pcpx)
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all scenarios hold" in r.stdout
