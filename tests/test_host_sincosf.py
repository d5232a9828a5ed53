"""lego_loam_b200/csrc/glibc_sincosf.cuh compiled for the host against the C library's sinf / cosf, the functions the
reference calls on float arguments: identical bits on 2e7 arguments in (-100, 100) on a host whose glibc selects its FMA
variant (every x86-64 CPU with FMA); the test also records how often libm differs from the correctly rounded value."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cpu_has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read()
    except OSError:
        return False


@pytest.mark.skipif(not cpu_has_fma(), reason="glibc selects its non-FMA sinf/cosf on this CPU")
def test_sincosf_restatement_matches_libm(tmp_path):
    so = str(tmp_path / "libhost_sincosf.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "host_sincosf_test.cpp"), "-lm"])
    L = ctypes.CDLL(so)
    out = (ctypes.c_long * 4)()
    L.host_sincosf_mismatches(ctypes.c_long(20_000_000), ctypes.c_uint(3), out)
    assert out[0] == 0 and out[1] == 0, list(out)
    assert out[2] > 100000 and out[3] > 50000        # libm is not the correctly rounded value in ~1 % of the cases
