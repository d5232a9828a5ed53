"""Batched multi-registration engine (llb_batch_*, BASELINE config 5) through the C ABI: every slot of a
batch must reproduce what the CPU oracle and the single-registration path give for the same inputs.
Bars: DS scan clouds bit-exact; poses identical to the oracle in correctly-rounded-trig mode and within the
north-star tolerance (1e-4 m / 1e-4 rad) in libm mode; iteration counts, convergence flags and row counts equal.
"""
import os

import numpy as np
import pytest

import oracle
from lego_loam_b200 import api, synth
from tests import data

pytestmark = pytest.mark.gpu

POSE_TOL = 1e-4


def _oracle_registration(case, mc_ds, ms_ds, trig_mode):
    oracle.set_trig_mode(trig_mode)
    mo = oracle.MapOptimization()
    mo.set_map_ds(mc_ds, ms_ds)
    mo.set_scan(case["corner"], case["surf"], case["outlier"])
    mo.downsampleCurrentScan()
    mo.transformTobeMapped = case["init"]
    iters = mo.scan2MapOptimization()
    ds = [mo.scan_ds(i) for i in range(4)]
    oracle.set_trig_mode(0)
    return mo.transformTobeMapped.copy(), iters, ds


@pytest.fixture(scope="module")
def cases(ctx):
    out = []
    for seed in (1, 2, 3, 4, 5):
        c = data.mapping_case(seed, 15000, 90000)
        ctx.map_set_raw(c["map_corner_raw"], c["map_surf_raw"])
        out.append(dict(c, mc_ds=ctx.map_get_ds(0), ms_ds=ctx.map_get_ds(1)))
    return out


def test_batch_matches_oracle_and_single(ctx, cases):
    B = len(cases)
    b = api.Batch(0, B, 8192, 120000)
    for s, c in enumerate(cases):
        b.scan_set(s, c["corner"], c["surf"], c["outlier"])
        b.map_set_ds(s, c["mc_ds"], c["ms_ds"])
    T, st = b.register(np.stack([c["init"] for c in cases]))
    for s, c in enumerate(cases):
        Tr, iters, ds = _oracle_registration(c, c["mc_ds"], c["ms_ds"], 0)        # the reference's arithmetic (libm sinf / cosf)
        for which in range(4):
            got = b.scan_get_ds(s, which)
            assert got.shape == ds[which].shape
            assert np.array_equal(got.view(np.uint32), ds[which].view(np.uint32)), f"slot {s} DS cloud {which}"
        assert st[s].iterations == iters, (s, st[s].as_dict(), iters)
        assert st[s].skipped == 0 and st[s].n_corner_ds == ds[0].shape[0] and st[s].n_surf_ds == ds[3].shape[0]
        assert np.array_equal(T[s].view(np.uint32), Tr.astype(np.float32).view(np.uint32)), (s, T[s], Tr)
        assert np.max(np.abs(T[s] - Tr)) < POSE_TOL                                # north-star bar (implied by the line above)
        # the single-registration path on the same inputs
        ctx.map_set_ds(c["mc_ds"], c["ms_ds"]); ctx.scan_set(c["corner"], c["surf"], c["outlier"])
        ctx.downsample_current_scan()
        Ts, sts = ctx.s2m_optimize(c["init"])
        assert np.array_equal(T[s].view(np.uint32), Ts.view(np.uint32))
        assert sts.iterations == st[s].iterations and sts.n_correspondences == st[s].n_correspondences
    b.close()


def test_batch_resident_map_and_repeat(cases):
    """A slot whose map is not set again keeps its index; a second step with new scans and device-resident
    inputs gives the same poses as the host-cloud path."""
    import torch
    B = 3
    b = api.Batch(0, B, 8192, 120000)
    for s in range(B):
        c = cases[s]
        b.scan_set(s, c["corner"], c["surf"], c["outlier"]); b.map_set_ds(s, c["mc_ds"], c["ms_ds"])
    T0, st0 = b.register(np.stack([cases[s]["init"] for s in range(B)]))
    l0 = b.launch_count()
    # step 2: same scans, maps untouched (no index build), inputs from device memory
    dev = torch.device("cuda", 0)
    keep = []
    for s in range(B):
        c = cases[s]
        t = [torch.from_numpy(np.ascontiguousarray(c[k], np.float32)).to(dev) for k in ("corner", "surf", "outlier")]
        keep.append(t)
        b.scan_set_dev(s, t[0].data_ptr(), t[0].shape[0], t[1].data_ptr(), t[1].shape[0], t[2].data_ptr(), t[2].shape[0])
    torch.cuda.synchronize()
    T1, st1 = b.register(np.stack([cases[s]["init"] for s in range(B)]))
    assert np.array_equal(T0.view(np.uint32), T1.view(np.uint32))
    assert [x.iterations for x in st0] == [x.iterations for x in st1]
    # launches of a step do not depend on the slot count and exclude the 5 index-build launches here
    # 2 voxel launches (downsampleCurrentScan), prepare, query ordering, ONE persistent registration kernel, collect
    assert b.launch_count() - l0 == 2 + 2 + 1 + 1
    b.close()


def test_batch_guard_and_errors(cases):
    b = api.Batch(0, 2, 8192, 120000)
    c = cases[0]
    with pytest.raises(api.LlbError) as e:
        b.register(np.zeros((2, 6), np.float32))              # nothing set yet
    assert e.value.status == api.LLB_ERR_STATE
    with pytest.raises(api.LlbError) as e:
        b.scan_set(0, np.zeros((9000, 4), np.float32), c["surf"], c["outlier"])          # beyond max_scan_points
    assert e.value.status == api.LLB_ERR_CAPACITY
    # slot 1 gets a map below the reference's guard (MO:1331: > 10 corner and > 100 surf points): skipped,
    # pose untouched; slot 0 registers normally
    b.scan_set(0, c["corner"], c["surf"], c["outlier"]); b.map_set_ds(0, c["mc_ds"], c["ms_ds"])
    b.scan_set(1, c["corner"], c["surf"], c["outlier"]); b.map_set_ds(1, c["mc_ds"][:10], c["ms_ds"][:500])
    init = np.stack([c["init"], c["init"]]).astype(np.float32)
    T, st = b.register(init)
    assert st[1].skipped == 1 and st[1].iterations == 0 and np.array_equal(T[1], init[1])
    assert st[0].skipped == 0 and st[0].iterations > 0 and not np.array_equal(T[0], init[0])
    b.close()
    with pytest.raises(api.LlbError):
        api.Batch(0, 2, 20000, 1000)                           # scan capacity beyond the shared-memory voxel kernel
    b2 = api.Batch(0, 1, 16384, 1000)
    with pytest.raises(api.LlbError) as e:                     # surf + outlier of one sweep must fit the fourth filter
        b2.scan_set(0, c["corner"], np.zeros((9000, 4), np.float32), np.zeros((9000, 4), np.float32))
    assert e.value.status == api.LLB_ERR_CAPACITY
    b2.close()


def test_batch_ragged_and_empty_clouds(ctx, cases):
    """Slots with very different sizes in one step: an empty outlier cloud, an empty corner cloud (registration runs on
    surf rows only), a tiny sweep (< 50 correspondences: LMOptimization returns early every iteration, MO:1238) - each
    slot must equal the single-registration path on the same inputs."""
    c = cases[0]
    empty = np.zeros((0, 4), np.float32)
    variants = [
        (c["corner"], c["surf"], empty),
        (empty, c["surf"], c["outlier"]),
        (c["corner"][:12], c["surf"][:40], c["outlier"][:5]),
        (c["corner"], c["surf"], c["outlier"]),
    ]
    b = api.Batch(0, len(variants), 8192, 120000)
    for s, (co, su, ou) in enumerate(variants):
        b.scan_set(s, co, su, ou); b.map_set_ds(s, c["mc_ds"], c["ms_ds"])
    init = np.stack([c["init"]] * len(variants)).astype(np.float32)
    T, st = b.register(init)
    for s, (co, su, ou) in enumerate(variants):
        ctx.map_set_ds(c["mc_ds"], c["ms_ds"]); ctx.scan_set(co, su, ou)
        counts = ctx.downsample_current_scan()
        Ts, sts = ctx.s2m_optimize(c["init"])
        assert np.array_equal(T[s].view(np.uint32), Ts.view(np.uint32)), (s, T[s], Ts)
        assert (st[s].iterations, st[s].converged, st[s].n_correspondences) == (sts.iterations, sts.converged, sts.n_correspondences)
        assert (st[s].n_corner_ds, st[s].n_surf_ds) == (counts[0], counts[3])
    assert st[2].iterations == 10 and st[2].converged == 0 and np.array_equal(T[2], init[2])   # too few rows: pose untouched
    b.close()


def test_keyframe_store_edge_cases(ctx, cases):
    c = cases[0]
    ctx.keyframe_clear()
    with pytest.raises(api.LlbError):                          # unknown key-frame id
        ctx.map_assemble([0], np.zeros((1, 6), np.float32))
    k0 = ctx.keyframe_add_clouds(c["corner"], c["surf"], c["outlier"])
    k1 = ctx.keyframe_add_clouds(np.zeros((0, 4), np.float32), c["surf"][:10], np.zeros((0, 4), np.float32))
    assert (k0, k1) == (0, 1) and ctx.keyframe_count() == 2
    pose = np.array([[0.01, 0.5, -0.02, 1.0, 0.2, -3.0], [0.0, 0.0, 0.0, 0.0, 0.0, 0.0]], np.float32)
    ctx.map_assemble([1, 0, 1], pose[[1, 0, 1]])               # repeated ids, an empty corner cloud in the middle
    raw_c, raw_s = ctx.map_get_raw(0), ctx.map_get_raw(1)
    assert raw_c.shape[0] == c["corner"].shape[0] and raw_s.shape[0] == 20 + c["surf"].shape[0] + c["outlier"].shape[0]
    # identity pose: the first surf segment is the stored cloud itself; transformed segment: reference formula in float32
    assert np.array_equal(raw_s[:10].view(np.uint32), np.ascontiguousarray(c["surf"][:10], np.float32).view(np.uint32))
    f = np.float32
    r, p, y = pose[0, :3]
    cr, sr, cp, sp, cy, sy = (f(np.cos(r, dtype=f)), f(np.sin(r, dtype=f)), f(np.cos(p, dtype=f)), f(np.sin(p, dtype=f)),
                              f(np.cos(y, dtype=f)), f(np.sin(y, dtype=f)))
    q = np.ascontiguousarray(c["corner"], f)
    x1 = cy * q[:, 0] - sy * q[:, 1]; y1 = sy * q[:, 0] + cy * q[:, 1]; z1 = q[:, 2]
    y2 = cr * y1 - sr * z1; z2 = sr * y1 + cr * z1
    exp = np.stack([cp * x1 + sp * z2 + pose[0, 3], y2 + pose[0, 4], -sp * x1 + cp * z2 + pose[0, 5], q[:, 3]], 1).astype(f)
    assert np.allclose(raw_c, exp, atol=2e-6, rtol=0)          # numpy's cosf/sinf may differ from libm by an ulp
    ctx.map_assemble([], np.zeros((0, 6), np.float32))         # empty selection: empty maps, guard MO:1331 skips
    assert ctx.map_get_raw(0).shape[0] == 0 and ctx.map_get_ds(1).shape[0] == 0
    ctx.scan_set(c["corner"], c["surf"], c["outlier"]); ctx.downsample_current_scan()
    T, st = ctx.s2m_optimize(c["init"])
    assert st.skipped == 1 and np.array_equal(T, np.asarray(c["init"], np.float32))
    ctx.keyframe_clear()


def test_batch_keyframe_stores(ctx, cases):
    """Per-slot device key-frame stores of the batch engine against the single-context store (which is pinned against
    the reference's node logic in test_gpu_sequence.py): key-frames saved from the slots' own DS clouds, local maps
    assembled in different orders, voxel-filtered by the BATCHED multi-kernel path (the context uses the cluster kernel
    below 16384 points), then a registration against them - raw maps, DS maps and poses bit-identical."""
    B, K = 2, 6
    b = api.Batch(0, B, 8192, 120000)
    with pytest.raises(api.LlbError):
        b.map_assemble(0, [0], np.zeros((1, 6), np.float32))   # not enabled yet
    b.enable_keyframes(400000, 16)
    for k in range(K):
        for s in range(B):
            c = cases[(s + k) % len(cases)]
            b.scan_set(s, c["corner"], c["surf"], c["outlier"]); b.map_set_ds(s, c["mc_ds"], c["ms_ds"])
        b.register(np.stack([cases[(s + k) % len(cases)]["init"] for s in range(B)]))
        for s in range(B):
            assert b.keyframe_add(s) == k
    orders = [list(range(K)), [4, 0, 5, 2, 1, 3]]
    rng = np.random.default_rng(3)
    kposes = [np.concatenate([rng.uniform(-0.03, 0.03, (K, 1)), rng.uniform(-3, 3, (K, 1)), rng.uniform(-0.03, 0.03, (K, 1)),
                              rng.uniform(-15, 15, (K, 1)), rng.uniform(-0.1, 0.1, (K, 1)), rng.uniform(-15, 15, (K, 1))], 1).astype(np.float32)
              for _ in range(B)]
    new = [cases[(s + 2) % len(cases)] for s in range(B)]
    for s in range(B):
        b.map_assemble(s, orders[s], kposes[s][orders[s]])
        b.scan_set(s, new[s]["corner"], new[s]["surf"], new[s]["outlier"])
    init = np.stack([kposes[s][0] for s in range(B)]).astype(np.float32)
    T, st = b.register(init)
    assert b.keyframe_count(0) == K
    for s in range(B):
        ctx.keyframe_clear()
        for k in range(K):
            c = cases[(s + k) % len(cases)]
            ctx.scan_set(c["corner"], c["surf"], c["outlier"]); ctx.downsample_current_scan(); ctx.keyframe_add()
        ctx.map_assemble(orders[s], kposes[s][orders[s]])
        for which in range(2):
            assert np.array_equal(b.map_get(s, which).view(np.uint32), ctx.map_get_raw(which).view(np.uint32)), (s, which)
            assert np.array_equal(b.map_get(s, 2 + which).view(np.uint32), ctx.map_get_ds(which).view(np.uint32)), (s, which)
        assert b.map_get(s, 1).shape[0] > 16384                 # the surf map is beyond the cluster kernel's range
        ctx.scan_set(new[s]["corner"], new[s]["surf"], new[s]["outlier"]); ctx.downsample_current_scan()
        Ts, sts = ctx.s2m_optimize(init[s])
        assert np.array_equal(T[s].view(np.uint32), Ts.view(np.uint32)), (s, T[s], Ts)
        assert (st[s].iterations, st[s].skipped, st[s].n_correspondences) == (sts.iterations, sts.skipped, sts.n_correspondences)
    # the assembled map stays until it is replaced: a second step without a new request registers against it again
    for s in range(B):
        b.scan_set(s, new[s]["corner"], new[s]["surf"], new[s]["outlier"])
    T2, _ = b.register(init)
    assert np.array_equal(T.view(np.uint32), T2.view(np.uint32))
    ctx.keyframe_clear()
    b.close()


def test_batch_odometry_matches_single_and_oracle(ctx):
    """updateTransformation (FA:1666-1695) of several slots in ONE launch against the single-context path and the
    oracle: poses bit-identical (correctly rounded trig), iteration counts equal; a second sweep on the same slots
    re-uses the per-slot state exactly like the context does (isDegenerate / matP, C6)."""
    from tests.test_gpu_parity import _odom_case
    ods = [_odom_case(s) for s in (1, 2, 3)]
    B = len(ods)
    b = api.Batch(0, B, 8192, 1000)
    with pytest.raises(api.LlbError):
        b.odom_optimize(np.zeros((B, 6), np.float32))              # nothing set
    for rep in range(2):
        for s, od in enumerate(ods):
            b.odom_set(s, od.corner_last, od.surf_last, od.corner_sharp, od.surf_flat)
        T, s0, s1 = b.odom_optimize(np.zeros((B, 6), np.float32))
        for s, od in enumerate(ods):
            c1 = api.Context(0)
            for _ in range(rep + 1):                               # same history as the slot
                c1.odom_set_last(od.corner_last, od.surf_last); c1.odom_set_features(od.corner_sharp, od.surf_flat)
                Ts, t0, t1 = c1.odom_optimize(np.zeros(6, np.float32))
            c1.close()
            assert np.array_equal(T[s].view(np.uint32), Ts.view(np.uint32)), (rep, s, T[s], Ts)
            assert (s0[s].iterations, s1[s].iterations) == (t0.iterations, t1.iterations)
            if rep == 0:
                oracle.set_trig_mode(0)
                fa = oracle.FeatureAssociation()
                fa.set_last(od.corner_last, od.surf_last, force=True); fa.set_features(od.corner_sharp, od.surf_flat)
                fa.transformCur = np.zeros(6, np.float32)
                fa.updateTransformation()
                oracle.set_trig_mode(0)
                assert np.array_equal(T[s].view(np.uint32), np.asarray(fa.transformCur, np.float32).view(np.uint32))
    b.close()
