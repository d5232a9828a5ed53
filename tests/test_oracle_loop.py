"""Loop closure on the CPU side: the restatement of pcl::IterativeClosestPoint (oracle/llo_loop.c,
"parity unpinned": PCL / Eigen are absent offline) recovers known rigid motions, and the UNMODIFIED reference
(oracle/_ref: detectLoopClosure MO:814-872, performLoopClosure MO:875-945, publishGlobalMap MO:758-800) runs through
it on a key-frame store built by its own node logic."""
import numpy as np
import pytest

import oracle
from oracle import ref_harness
from tests import data


def _rigid(yaw, t):
    R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
    return T


@pytest.mark.parametrize("seed,yaw,t", [(0, 0.02, (0.2, 0.05, -0.15)), (1, -0.05, (-0.4, 0.0, 0.3)), (2, 0.0, (0.0, 0.0, 0.0))])
def test_icp_recovers_known_motion(seed, yaw, t):
    rng = np.random.default_rng(seed)
    tgt = np.zeros((20000, 4), np.float32); tgt[:, :3] = rng.uniform(-30, 30, (20000, 3)) * [1, 0.1, 1]
    src = tgt[rng.choice(20000, 3000, replace=False)].copy()
    M = _rigid(yaw, np.array(t))
    s2 = src.copy(); s2[:, :3] = (src[:, :3] @ M[:3, :3].T + M[:3, 3]).astype(np.float32)
    r = oracle.icp_align(s2, tgt)
    assert r["converged"] and r["state"] in (2, 3, 4) and r["iterations"] < 20
    assert np.abs(r["T"] - np.linalg.inv(M)).max() < 5e-6 and r["fitness"] < 1e-9


def test_icp_convergence_criteria_and_guards():
    rng = np.random.default_rng(5)
    tgt = np.zeros((5000, 4), np.float32); tgt[:, :3] = rng.uniform(-20, 20, (5000, 3))
    src = tgt[:800].copy(); src[:, :3] += np.float32(0.3)
    one = oracle.icp_align(src, tgt, max_iterations=1)
    assert one["converged"] and one["state"] == 1 and one["iterations"] == 1          # CONVERGENCE_CRITERIA_ITERATIONS
    st = oracle.icp_step(src, tgt)
    assert st["n"] == 800 and np.abs(one["T"] - st["Rt"]).max() == 0.0                 # first step = umeyama of the 1-NN pairs
    far = oracle.icp_align(src + np.float32(1000.0), tgt, max_corr_dist=1.0)          # nothing within the gate
    assert not far["converged"] and far["state"] == 0 and far["iterations"] == 0 and np.array_equal(far["T"], np.eye(4, dtype=np.float32))
    nn = oracle.knn_bruteforce(tgt, src, 1)[0][:, 0]
    assert np.array_equal(st["nn"], nn)


@pytest.mark.skipif(not ref_harness.available(), reason="oracle/_ref not built")
def test_reference_loop_closure_runs_through_the_restated_icp():
    case = data.loop_closure_case()
    mo = case["mo"]
    assert case["n_kf"] >= 40
    last = mo.keypose6d(case["n_kf"] - 1)
    mo.set_robot_pos(last[3], last[4], last[5]); mo.set_time(case["t_last"])
    assert mo.detectLoopClosure()
    closest, latest = mo.loop_ids()
    assert latest == case["n_kf"] - 1 and 0 <= closest < latest - 10
    src, hist, hist_ds = mo.loop_cloud(0), mo.loop_cloud(1), mo.loop_cloud(2)
    assert src.shape[0] > 500 and hist_ds.shape[0] > 2000 and hist.shape[0] > hist_ds.shape[0]
    want_ds, _ = oracle.voxel_grid(hist, 0.4)
    assert np.array_equal(want_ds.view(np.uint32), hist_ds.view(np.uint32))            # downSizeFilterHistoryKeyFrames MO:860
    calls = mo.icp_last()["calls"]
    closed = mo.performLoopClosure()
    rec = mo.icp_last()
    assert rec["calls"] == calls + 1 and rec["converged"] and 1 <= rec["iterations"] <= 100
    assert closed == (rec["fitness"] <= 0.3)                                            # historyKeyframeFitnessScore UT:134
    direct = oracle.icp_align(src, hist_ds)
    assert np.array_equal(direct["T"], rec["T"]) and direct["iterations"] == rec["iterations"]
    # the alignment undoes (most of) the odometry drift of the latest key-frame (0.25 / 0.2 m in x / z): its corrected
    # position lands near the true one
    drifted = np.array([last[3], last[4], last[5], 1.0])
    fixed = rec["T"].astype(np.float64) @ drifted
    err_before = np.linalg.norm(drifted[:3] - case["truth"][-1][3:6]); err_after = np.linalg.norm(fixed[:3] - case["truth"][-1][3:6])
    assert rec["fitness"] < 0.5 and err_before > 0.25 and err_after < 0.6 * err_before, (err_before, err_after, rec)


@pytest.mark.skipif(not ref_harness.available(), reason="oracle/_ref not built")
def test_reference_global_map():
    case = data.loop_closure_case()
    mo = case["mo"]
    last = mo.keypose6d(case["n_kf"] - 1)
    mo.set_robot_pos(last[3], last[4], last[5])
    ids = mo.publishGlobalMap()
    assert ids.shape[0] >= 10 and len(set(ids.tolist())) == ids.shape[0]
    g = mo.loop_cloud(3)
    assert g.shape[0] > 5000


def test_restated_icp_against_committed_golden_vectors():
    """tests/golden/ref_loop_golden.npz (made by tests/golden/make_loop_golden.py from oracle/_ref): the clouds the
    unmodified detectLoopClosure built and the alignment performLoopClosure ran.  Pins the restatement against
    regressions on machines without /root/reference."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_loop_golden.npz"))
    src, tgt = g["source"], g["target_ds"]
    st = oracle.icp_step(src, tgt)
    assert st["n"] == int(g["first_step_n"]) and np.array_equal(st["nn"], g["first_step_nn"])
    assert np.array_equal(st["Rt"], g["first_step_Rt"]) and st["mse"] == float(g["first_step_mse"])
    r = oracle.icp_align(src, tgt)
    assert r["iterations"] == int(g["iterations"]) and r["state"] == int(g["state"]) and r["converged"] == bool(g["converged"])
    assert np.array_equal(r["T"], g["T"]) and r["fitness"] == float(g["fitness"])
    # exact 1-NN of the first pass against the brute-force definition
    nn = oracle.knn_bruteforce(tgt, src, 1)[0][:, 0]
    assert np.array_equal(st["nn"], nn)
