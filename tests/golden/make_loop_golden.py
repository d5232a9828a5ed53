"""Generate tests/golden/ref_loop_golden.npz: the loop-closure clouds the UNMODIFIED reference builds
(detectLoopClosure MO:814-872 of oracle/_ref, on the key-frame store of tests.data.loop_closure_case) and the alignment its
performLoopClosure ran (pcl::IterativeClosestPoint = the restated PCL 1.8 algorithm of oracle/llo_loop.c behind the shim).
Needs /root/reference; run in the build container:  python tests/golden/make_loop_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import ref_harness  # noqa: E402
from tests import data  # noqa: E402

assert ref_harness.build(), "reference harness not built"
case = data.loop_closure_case()
mo = case["mo"]
last = mo.keypose6d(case["n_kf"] - 1)
mo.set_robot_pos(last[3], last[4], last[5]); mo.set_time(case["t_last"])
assert mo.detectLoopClosure()
closest, latest = mo.loop_ids()
src, tgt = mo.loop_cloud(0), mo.loop_cloud(2)
mo.performLoopClosure()
rec = mo.icp_last()
step = oracle.icp_step(src, tgt)
out = dict(source=src, target_ds=tgt, closest=np.int32(closest), latest=np.int32(latest), T=rec["T"],
           iterations=np.int32(rec["iterations"]), state=np.int32(rec["state"]), converged=np.int32(rec["converged"]),
           fitness=np.float64(rec["fitness"]), first_step_n=np.int32(step["n"]), first_step_Rt=step["Rt"],
           first_step_nn=step["nn"].astype(np.int32), first_step_mse=np.float64(step["mse"]))
p = os.path.join(ROOT, "tests", "golden", "ref_loop_golden.npz")
np.savez_compressed(p, **out); print("wrote", p, os.path.getsize(p))
