"""The device linear-algebra header (lego_loam_b200/csrc/linalg.cuh) is __host__ __device__: compile it
for the host with nvcc and check it bit-for-bit against the oracle's OpenCV restatement (no GPU needed)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="no nvcc")
def test_device_linalg_matches_oracle_on_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "host_linalg_test")
    subprocess.check_call([nvcc, "-O2", "-fmad=false", "-std=c++17", "-Xcompiler", "-ffp-contract=off",
                           "-o", exe, os.path.join(ROOT, "tests", "host_linalg_test.cu"),
                           os.path.join(ROOT, "oracle", "llo_linalg.c"), "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
