"""lego_loam_b200/csrc/glibc_atan2f.cuh (the feature-extraction kernel's atan2f) compiled for the host against the C
library's atan2f - the function the reference calls at FA:504: identical bits on 2e7 arguments incl. the special cases.
Compiled without FMA contraction, as the CUDA build is (-fmad=false)."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_atan2f_restatement_matches_libm(tmp_path):
    so = str(tmp_path / "libhost_atan2f.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "host_atan2f_test.cpp"), "-lm"])
    L = ctypes.CDLL(so)
    L.host_atan2f_mismatches.restype = ctypes.c_long
    bad4 = np.zeros(4, np.float32)
    bad = L.host_atan2f_mismatches(ctypes.c_long(20_000_000), ctypes.c_uint(7), bad4.ctypes.data_as(ctypes.c_void_p))
    assert bad == 0, (bad, bad4.tolist())
