"""The product's restatement of libstdc++ std::sort (lego_loam_b200/csrc/std_sort.cuh, used by the feature-extraction
kernel) compiled for the host: same order of equal-curvature records as the oracle restatement - which
tests/test_oracle_features.py pins against libstdc++ itself - and as libstdc++ when oracle/_ref is present."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from oracle import ref_harness as rh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_sort(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("stdsort") / "libhost_std_sort.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "host_std_sort_test.cpp")])
    L = ctypes.CDLL(so)

    def run(v, ind, depth=-1, closed=0):
        v = np.ascontiguousarray(v, np.float32).copy(); i = np.ascontiguousarray(ind, np.uint32).copy()
        L.host_std_sort(v.ctypes.data_as(ctypes.c_void_p), i.ctypes.data_as(ctypes.c_void_p), v.shape[0], int(depth), int(closed))
        return v, i
    return run


@pytest.mark.parametrize("n", [0, 1, 2, 15, 16, 17, 33, 100, 301, 1000, 1800])
@pytest.mark.parametrize("depth", [-1, 0, 2, 5])
def test_device_std_sort_on_host(host_sort, n, depth):
    rng = np.random.default_rng(n + 100 * (depth + 1))
    for levels in (2, 5, 40, 5000, 0):
        for shape in ("random", "sorted", "reverse", "pipe"):
            v = (rng.integers(0, levels, n).astype(np.float32) * np.float32(0.125)) if levels else rng.random(n).astype(np.float32)
            if shape == "sorted": v = np.sort(v)
            if shape == "reverse": v = np.sort(v)[::-1].copy()
            if shape == "pipe": v = np.concatenate([np.sort(v[: n // 2]), np.sort(v[n // 2:])[::-1]])
            ind = np.arange(n)
            gv, gi = host_sort(v, ind, depth)
            ov, oi = oracle.std_sort_by_value(v, ind, depth)
            assert np.array_equal(gv, ov) and np.array_equal(gi, oi)
            cv, ci = host_sort(v, ind, depth, closed=1)          # the data-parallel formulation the kernel executes
            assert np.array_equal(cv, ov) and np.array_equal(ci, oi)
            if rh.available():
                rv, ri = rh.std_sort(v, ind, depth)
                assert np.array_equal(gv, rv) and np.array_equal(gi.astype(np.int64), ri.astype(np.int64))
