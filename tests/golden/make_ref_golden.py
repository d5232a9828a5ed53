"""Generate tests/golden/ref_*.npz from the reference harness (oracle/_ref = the unmodified reference
sources compiled against shims).  Needs /root/reference; run in the build container:
    python tests/golden/make_ref_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402
from tests import data  # noqa: E402
from tests.test_gpu_parity import _odom_case  # noqa: E402

assert ref_harness.build(), "reference harness not built"
out = {}
seeds = (11, 12)
for k, seed in enumerate(seeds):
    case = data.mapping_case(seed, 6000, 30000)
    mo = ref_harness.MapOptimization()
    mo.set_map_raw(case["map_corner_raw"], case["map_surf_raw"])
    mo.set_scan(case["corner"], case["surf"], case["outlier"])
    mo.downsampleCurrentScan()
    mo.transformTobeMapped = case["init"]
    out[f"map_corner_ds_{k}"] = mo.map_ds(0); out[f"map_surf_ds_{k}"] = mo.map_ds(1)
    out[f"corner_{k}"] = case["corner"]; out[f"surf_{k}"] = case["surf"]; out[f"outlier_{k}"] = case["outlier"]
    out[f"corner_ds_{k}"] = mo.scan_ds(0); out[f"surf_total_ds_{k}"] = mo.scan_ds(3)
    out[f"init_{k}"] = case["init"]
    mo.scan2MapOptimization()
    out[f"pose_{k}"] = mo.transformTobeMapped
out["n_cases"] = np.int32(len(seeds))
p = os.path.join(ROOT, "tests", "golden", "ref_scan2map_golden.npz")
np.savez_compressed(p, **out); print("wrote", p, os.path.getsize(p))

out = {}
for k, seed in enumerate((11, 12)):
    od = _odom_case(seed)
    fa = ref_harness.FeatureAssociation()
    fa.set_last(od.corner_last, od.surf_last, force=True)
    fa.set_features(od.corner_sharp, od.surf_flat)
    fa.transformCur = np.zeros(6, np.float32)
    fa.updateTransformation()
    out[f"corner_last_{k}"] = od.corner_last; out[f"surf_last_{k}"] = od.surf_last
    out[f"sharp_{k}"] = od.corner_sharp; out[f"flat_{k}"] = od.surf_flat; out[f"cur_{k}"] = fa.transformCur
out["n_cases"] = np.int32(2)
p = os.path.join(ROOT, "tests", "golden", "ref_odometry_golden.npz")
np.savez_compressed(p, **out); print("wrote", p, os.path.getsize(p))

# ---- feature extraction (FA:491-784) + TransformToEnd (FA:885-953): a 2-sweep sequence through ONE reference object (state
# ---- survives between sweeps), ranges quantised to 2 cm so that equal curvatures (std::sort ties) are common
import dataclasses  # noqa: E402
from lego_loam_b200 import synth  # noqa: E402

out = {}
w = synth.make_world()
fa = ref_harness.FeatureAssociation()
T = np.array([0.002, 0.015, -0.001, 0.01, 0.005, -0.15], np.float32)
for k in range(2):
    sw = synth.make_segmented_sweep(w, synth.VLP16, [0.002 * k, 0.05 + 0.01 * k, 0, 3 + 0.4 * k, 0, 5], 21 + k)
    sw = dataclasses.replace(sw, range=(np.round(sw.range / 0.02) * 0.02).astype(np.float32))
    fa.set_segmented(sw); fa.extract_features()
    for name in ("cloud", "start_ring", "end_ring", "ground", "col", "range"):
        out[f"in_{name}_{k}"] = getattr(sw, name)
    out[f"in_ori_{k}"] = np.array([sw.start_ori, sw.end_ori, sw.ori_diff], np.float32)
    for which, name in enumerate(("sharp", "less_sharp", "flat", "less_flat")):
        out[f"out_{name}_{k}"] = fa.feature_cloud(which)
    out[f"out_adjusted_intensity_{k}"] = fa.feature_cloud(4)[:, 3].copy()      # x, y, z of the adjusted cloud = input y, z, x
    curv, picked, label = fa.point_state()
    out[f"out_label_{k}"] = label.astype(np.int8); out[f"out_picked_{k}"] = picked.astype(np.int8)
    fa.transformCur = T; fa.publishCloudsLast()
    out[f"out_corner_last_{k}"] = fa.feature_cloud(5); out[f"out_surf_last_{k}"] = fa.feature_cloud(6)
out["T"] = T; out["n_sweeps"] = np.int32(2)
p = os.path.join(ROOT, "tests", "golden", "ref_features_golden.npz")
np.savez_compressed(p, **out); print("wrote", p, os.path.getsize(p))
