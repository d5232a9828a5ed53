"""The C-ABI library loads without a GPU and exports every symbol include/llb200.h declares; creating a
context without a device fails loudly (there is no CPU fallback in the product path)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "llb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(llb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from lego_loam_b200 import api
    api.build()
    L = ctypes.CDLL(api.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert set(api.EXPORTS) == set(syms)
    assert L.llb_abi_version() == 1


def test_point_layout_is_pcl_xyzi():
    from lego_loam_b200 import api
    import numpy as np
    p = api.to_pcl(np.array([[1, 2, 3, 4]], np.float32))
    assert p.nbytes == 32 and p[0, 3] == 1.0 and p[0, 4] == 4.0       # x0 y4 z8 (1.0f) intensity16
    assert np.array_equal(api.from_pcl(p), np.array([[1, 2, 3, 4]], np.float32))


def test_params_default_are_reference_constants():
    from lego_loam_b200 import api
    p = api.default_params()
    assert (round(p.corner_leaf, 6), round(p.surf_leaf, 6), round(p.outlier_leaf, 6)) == (0.2, 0.4, 0.4)   # MO:249-251
    assert p.knn_max_sqdist == 1.0 and p.s2m_max_iterations == 10 and p.s2m_min_correspondences == 50
    assert p.s2m_degeneracy_thresh == 100.0 and p.corner_map_min == 10 and p.surf_map_min == 100
    assert p.odom_nearest_sqdist == 25.0 and p.odom_max_iterations == 25 and p.odom_degeneracy_thresh == 10.0


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from lego_loam_b200 import api
    with pytest.raises(api.LlbError) as e:
        api.Context(0)
    assert e.value.status == api.LLB_ERR_NO_DEVICE


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lego_loam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "llo.h" not in txt, f


def test_batch_engine_needs_a_device_too():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from lego_loam_b200 import api
    with pytest.raises(api.LlbError) as e:
        api.Batch(0, 4, 8192, 100000)
    assert e.value.status == api.LLB_ERR_NO_DEVICE
