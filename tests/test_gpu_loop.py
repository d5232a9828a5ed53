"""Loop closure + global map on the device against the UNMODIFIED reference (oracle/_ref): the cloud
part of detectLoopClosure (MO:838-861) and of publishGlobalMap (MO:780-788) from the device key-frame store must be
bit-identical; the ICP of performLoopClosure (MO:892-904) is compared with the restatement of PCL 1.8's algorithm the
reference ran through (oracle/llo_loop.c - PCL itself is absent, "parity unpinned"): nearest neighbours bit-exact, sums to
fp64 rounding, final transformation within 1e-4, same iteration count and convergence state."""
import numpy as np
import pytest

import oracle
from lego_loam_b200 import api
from oracle import ref_harness
from tests import data

pytestmark = pytest.mark.gpu


def _store(ctx, mo, n_kf):
    ctx.keyframe_clear()
    for i in range(n_kf):
        assert ctx.keyframe_add_clouds(mo.keyframe_cloud(i, 0), mo.keyframe_cloud(i, 1), mo.keyframe_cloud(i, 2)) == i


def _apply(T, pts):
    """IterativeClosestPoint::transformCloud in float32, columns added left to right"""
    T = T.astype(np.float32); x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    out = pts.copy()
    for r in range(3):
        out[:, r] = ((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3]
    return out


@pytest.mark.skipif(not ref_harness.available(), reason="oracle/_ref not built")
def test_loop_closure_matches_reference():
    case = data.loop_closure_case()
    mo, n_kf = case["mo"], case["n_kf"]
    ctx = api.Context(0)
    _store(ctx, mo, n_kf)
    last = mo.keypose6d(n_kf - 1)
    mo.set_robot_pos(last[3], last[4], last[5]); mo.set_time(case["t_last"])
    assert mo.detectLoopClosure()
    closest, latest = mo.loop_ids()
    hist = [j for j in range(closest - 25, closest + 26) if 0 <= j <= latest]              # MO:853-855
    n_src, n_hist = ctx.loop_set_clouds(latest, mo.keypose6d(latest), hist, np.stack([mo.keypose6d(j) for j in hist]))
    ref_src, ref_hist, ref_hist_ds = mo.loop_cloud(0), mo.loop_cloud(1), mo.loop_cloud(2)
    assert n_src == ref_src.shape[0] and n_hist == ref_hist_ds.shape[0]
    assert np.array_equal(ctx.loop_get_cloud(0).view(np.uint32), ref_src.view(np.uint32))
    assert np.array_equal(ctx.loop_get_cloud(1).view(np.uint32), ref_hist.view(np.uint32))
    assert np.array_equal(ctx.loop_get_cloud(2).view(np.uint32), ref_hist_ds.view(np.uint32))
    # ---- one iteration: correspondences and sums of the first pass
    p1 = api.default_loop_params(); p1.max_iterations = 1
    r1 = ctx.loop_icp(p1)
    st = oracle.icp_step(ref_src, ref_hist_ds)
    assert r1.iterations == 1 and r1.convergence_state == 1 and r1.has_converged == 1 and r1.n_correspondences == st["n"]
    assert np.abs(np.array(r1.T).reshape(4, 4) - st["Rt"]).max() < 2e-7
    idx, d2 = ctx.loop_get_nn()                                  # the fitness pass: source through T of that iteration
    moved = _apply(np.array(r1.T, np.float32).reshape(4, 4), ref_src)
    want_idx, want_d2 = oracle.knn_bruteforce(ref_hist_ds, moved, 1)
    assert np.array_equal(idx, want_idx[:, 0]) and np.array_equal(d2.view(np.uint32), want_d2[:, 0].view(np.uint32))
    assert abs(r1.fitness_score - float(np.mean(want_d2[:, 0].astype(np.float64)))) < 1e-12 * max(1.0, r1.fitness_score)
    # ---- the whole alignment as performLoopClosure configures it
    mo.performLoopClosure()
    rec = mo.icp_last()
    r = ctx.loop_icp()
    T = np.array(r.T, np.float32).reshape(4, 4)
    assert r.n_source == ref_src.shape[0] and r.n_target == ref_hist_ds.shape[0]
    assert r.has_converged == int(rec["converged"]) and r.iterations == rec["iterations"] and r.convergence_state == rec["state"], (r.iterations, rec)
    assert np.abs(T - rec["T"]).max() < 1e-4, np.abs(T - rec["T"]).max()
    assert abs(r.fitness_score - rec["fitness"]) < 1e-4 * rec["fitness"]
    ctx.close()


@pytest.mark.skipif(not ref_harness.available(), reason="oracle/_ref not built")
def test_global_map_matches_reference():
    case = data.loop_closure_case()
    mo, n_kf = case["mo"], case["n_kf"]
    ctx = api.Context(0)
    _store(ctx, mo, n_kf)
    last = mo.keypose6d(n_kf - 1)
    mo.set_robot_pos(last[3], last[4], last[5])
    ids = mo.publishGlobalMap()                                  # host part of MO:766-778 (radius search + 1 m filter of the poses)
    want = mo.loop_cloud(3)
    n = ctx.global_map_assemble(ids, np.stack([mo.keypose6d(i) for i in ids]), 0.4)
    got = ctx.loop_get_cloud(3)
    assert n == want.shape[0] and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    ctx.close()


def test_icp_guards_and_known_motion():
    rng = np.random.default_rng(3)
    tgt = np.zeros((30000, 4), np.float32); tgt[:, :3] = rng.uniform(-30, 30, (30000, 3)) * [1, 0.1, 1]
    src = tgt[rng.choice(30000, 4000, replace=False)].copy()
    a = 0.03; R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]); t = np.array([0.3, -0.05, 0.2])
    s2 = src.copy(); s2[:, :3] = (src[:, :3] @ R.T + t).astype(np.float32)
    ctx = api.Context(0)
    ctx.loop_set_clouds_host(s2, tgt)
    r = ctx.loop_icp()
    want = oracle.icp_align(s2, tgt)
    T = np.array(r.T, np.float32).reshape(4, 4)
    Minv = np.eye(4); Minv[:3, :3] = R.T; Minv[:3, 3] = -R.T @ t
    assert r.has_converged and r.iterations == want["iterations"] and r.convergence_state == want["state"]
    assert np.abs(T - want["T"]).max() < 1e-6 and np.abs(T - Minv).max() < 5e-6 and r.fitness_score < 1e-9
    # nothing inside the correspondence gate: hasConverged() false, identity, like PCL's "not enough correspondences"
    p = api.default_loop_params(); p.max_correspondence_distance = 0.5
    ctx.loop_set_clouds_host(s2 + np.float32([500, 0, 0, 0]), tgt)
    far = ctx.loop_icp(p)
    assert far.has_converged == 0 and far.iterations == 0 and far.convergence_state == 0
    assert np.array_equal(np.array(far.T, np.float32).reshape(4, 4), np.eye(4, dtype=np.float32))
    # empty clouds and the call order
    ctx.loop_set_clouds_host(np.zeros((0, 4), np.float32), tgt)
    e = ctx.loop_icp()
    assert e.has_converged == 0 and e.iterations == 0
    fresh = api.Context(0)
    with pytest.raises(api.LlbError):
        fresh.loop_icp()
    fresh.close(); ctx.close()


def test_loop_argument_errors_and_empty_selection():
    ctx = api.Context(0)
    pose = np.zeros(6, np.float32)
    with pytest.raises(api.LlbError) as e:
        ctx.loop_set_clouds(0, pose, [0], [pose])             # empty key-frame store: id 0 does not exist
    assert e.value.status == api.LLB_ERR_INVALID
    rng = np.random.default_rng(1)
    cl = [np.concatenate([rng.uniform(-5, 5, (n, 3)), rng.uniform(0, 15, (n, 1))], 1).astype(np.float32) for n in (300, 900, 500)]
    assert ctx.keyframe_add_clouds(*cl) == 0
    with pytest.raises(api.LlbError):
        ctx.loop_set_clouds(0, pose, [0, 1], [pose, pose])    # id 1 does not exist
    with pytest.raises(api.LlbError):
        ctx.global_map_assemble([0], [pose], leaf=0.0)
    assert ctx.global_map_assemble([], np.zeros((0, 6), np.float32)) == 0 and ctx.loop_get_cloud(3).shape[0] == 0
    # one key-frame at the identity pose: the global map is the voxel filter of its three clouds, the latest cloud is
    # corner + surf, the history cloud its VoxelGrid(0.4)
    n = ctx.global_map_assemble([0], [pose])
    want, _ = oracle.voxel_grid(np.concatenate(cl), 0.4)
    assert n == want.shape[0] and np.array_equal(ctx.loop_get_cloud(3).view(np.uint32), want.view(np.uint32))
    n_src, n_hist = ctx.loop_set_clouds(0, pose, [0], [pose])
    assert n_src == 1200 and np.array_equal(ctx.loop_get_cloud(0).view(np.uint32), np.concatenate(cl[:2]).view(np.uint32))
    want_h, _ = oracle.voxel_grid(np.concatenate(cl[:2]), 0.4)
    assert n_hist == want_h.shape[0] and np.array_equal(ctx.loop_get_cloud(2).view(np.uint32), want_h.view(np.uint32))
    # a negative intensity is dropped from the latest cloud (MO:845-849)
    cl2 = [c.copy() for c in cl]; cl2[0][5, 3] = -2.0; cl2[1][7, 3] = -1.0
    assert ctx.keyframe_add_clouds(*cl2) == 1
    n_src, _ = ctx.loop_set_clouds(1, pose, [0], [pose])
    keep = np.concatenate(cl2[:2]); keep = keep[keep[:, 3].astype(np.int32) >= 0]
    assert n_src == 1198 and np.array_equal(ctx.loop_get_cloud(0).view(np.uint32), keep.view(np.uint32))
    r = ctx.loop_icp()
    assert r.has_converged == 1 and r.n_source == 1198
    ctx.close()


def test_icp_against_committed_golden_vectors():
    """tests/golden/ref_loop_golden.npz: the clouds of the reference's own detectLoopClosure and the alignment its
    performLoopClosure ran (restated PCL behind the shim) - available on the GPU box without /root/reference."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_loop_golden.npz"))
    ctx = api.Context(0)
    ctx.loop_set_clouds_host(g["source"], g["target_ds"])
    r = ctx.loop_icp()
    assert r.has_converged == int(g["converged"]) and r.iterations == int(g["iterations"]) and r.convergence_state == int(g["state"])
    assert np.abs(np.array(r.T, np.float32).reshape(4, 4) - g["T"]).max() < 1e-4
    assert abs(r.fitness_score - float(g["fitness"])) < 1e-4 * float(g["fitness"])
    p1 = api.default_loop_params(); p1.max_iterations = 1
    r1 = ctx.loop_icp(p1)
    assert r1.n_correspondences == int(g["first_step_n"]) and np.abs(np.array(r1.T, np.float32).reshape(4, 4) - g["first_step_Rt"]).max() < 2e-7
    ctx.close()
