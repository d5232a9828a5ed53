"""GPU parity tests proper: the CUDA path (through the C ABI, via ctypes) against the CPU
oracle on the same seeded inputs.  Bars (BASELINE.json north_star):
  * voxel outputs bit-exact (membership, order, centroids, intensity);
  * kNN index sets bit-exact for every query the reference would use (d2[4] < 1.0);
  * correspondences / normal equations / poses bit-identical to the oracle in the reference's own arithmetic (the host
    libm's sinf / cosf, restated on the device by csrc/glibc_sincosf.cuh) and to the compiled reference itself
    (oracle/_ref, tier B) where it is present; the north-star tolerance 1e-4 m / 1e-4 rad is kept as a second bar.
"""
import numpy as np
import pytest

import oracle
from lego_loam_b200 import synth
from tests import data

pytestmark = pytest.mark.gpu

POSE_TOL_M = 1e-4      # north_star: 1e-4 m per scan
POSE_TOL_RAD = 1e-4    # north_star: 1e-4 rad per scan


def assert_clouds_bitexact(a, b, what=""):
    assert a.shape == b.shape, f"{what}: {a.shape} vs {b.shape}"
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), \
        f"{what}: {np.sum(np.any(a.view(np.uint32) != b.view(np.uint32), axis=1))} rows differ"


# ------------------------------------------------------------------ K1 voxel

@pytest.mark.parametrize("n,leaf,seed", [
    (1, 0.2, 1), (2, 0.4, 2), (31, 0.2, 3), (1000, 0.2, 4), (5000, 0.4, 5),
    (16384, 0.4, 6),           # largest single-CTA case
    (16385, 0.4, 7),           # smallest multi-kernel case
    (60000, 0.2, 8), (300000, 0.4, 9),
])
def test_voxel_bitexact(ctx, n, leaf, seed):
    pts = data.random_cloud(n, seed)
    ref, ovf = oracle.voxel_grid(pts, leaf)
    assert ovf == 0
    out = ctx.voxel_downsample(pts, leaf)
    assert_clouds_bitexact(out, ref, f"voxel n={n}")


def test_voxel_empty_and_duplicates(ctx):
    assert ctx.voxel_downsample(np.zeros((0, 4), np.float32), 0.2).shape == (0, 4)
    pts = np.tile(np.array([[1.05, -2.3, 7.7, 3.0]], np.float32), (500, 1))
    ref, _ = oracle.voxel_grid(pts, 0.4)
    assert_clouds_bitexact(ctx.voxel_downsample(pts, 0.4), ref, "duplicates")


@pytest.mark.parametrize("n", [2000, 40000])
def test_voxel_int32_overflow_passthrough(ctx, n):
    # extent so large that dx*dy*dz > INT32_MAX at leaf 0.2: PCL warns and copies the input (C18)
    pts = data.random_cloud(n, 11, extent=(900.0, 300.0, 900.0), clustered=False)
    ref, ovf = oracle.voxel_grid(pts, 0.2)
    assert ovf == 1 and ref.shape[0] == n
    assert_clouds_bitexact(ctx.voxel_downsample(pts, 0.2), ref, "overflow")


def test_voxel_properties_full_size(ctx):
    """BASELINE-size raw map (2M points): size-independent properties instead of the oracle."""
    n = 2_000_000
    pts = data.random_cloud(n, 21, extent=(120.0, 10.0, 120.0))
    out = ctx.voxel_downsample(pts, 0.4)
    inv = np.float32(1.0) / np.float32(0.4)
    key_in = np.floor(pts[:, :3] * inv).astype(np.int64)
    key_out = np.floor(out[:, :3] * inv).astype(np.int64)
    # one output per occupied voxel
    assert out.shape[0] == np.unique(key_in, axis=0).shape[0]
    # output sorted by (z, y, x) voxel index, strictly increasing
    mn = key_in.min(0); dv = key_in.max(0) - mn + 1
    lin = lambda k: (k[:, 0] - mn[0]) + dv[0] * ((k[:, 1] - mn[1]) + dv[1] * (k[:, 2] - mn[2]))
    lo = lin(key_out)
    # centroids of points inside a voxel can round onto the voxel's upper face; allow equality of neighbours
    assert np.all(np.diff(lo) >= 0) and np.mean(np.diff(lo) > 0) > 0.999
    # mass conservation: count-weighted centroids reproduce the global sum
    cnt = np.bincount(np.unique(lin(key_in), return_inverse=True)[1])
    assert np.allclose((out[:, :3].astype(np.float64) * cnt[:, None]).sum(0), pts[:, :3].astype(np.float64).sum(0),
                       rtol=1e-5, atol=1.0)


# ------------------------------------------------------------------ a2 / a3

def _setup(ctx, case, trig_mode):
    oracle.set_trig_mode(trig_mode)
    mo = oracle.MapOptimization()
    mo.set_map_raw(case["map_corner_raw"], case["map_surf_raw"])
    mo.set_scan(case["corner"], case["surf"], case["outlier"])
    mo.downsampleCurrentScan()
    ctx.map_set_raw(case["map_corner_raw"], case["map_surf_raw"])
    ctx.scan_set(case["corner"], case["surf"], case["outlier"])
    counts = ctx.downsample_current_scan()
    return mo, counts


def test_downsample_current_scan_and_map_bitexact(ctx):
    case = data.mapping_case(1)
    mo, counts = _setup(ctx, case, 0)
    for which in range(4):                                   # cornerDS, surfDS, outlierDS, surfTotalDS (C12)
        ref = mo.scan_ds(which)
        assert counts[which] == ref.shape[0]
        assert_clouds_bitexact(ctx.scan_get_ds(which), ref, f"scan ds {which}")
    for which in range(2):                                   # MO:1057-1064
        assert_clouds_bitexact(ctx.map_get_ds(which), mo.map_ds(which), f"map ds {which}")


# ------------------------------------------------------------------ K2 + K3 + K4, one iteration

@pytest.mark.parametrize("seed", [1, 2, 3])
def test_iteration_knn_rows_normal_equations_bitexact(ctx, seed):
    case = data.mapping_case(seed)
    mo, _ = _setup(ctx, case, 0)
    mo.build_kdtrees()
    T = case["init"].copy()
    for it in range(3):
        mo.transformTobeMapped = T
        mo.clear_correspondences()
        mo.cornerOptimization(it); mo.surfOptimization(it)
        conv_ref = mo.LMOptimization(it)
        T_gpu, conv, n_corr = ctx.s2m_iterate(T, it)

        # kNN index sets: exact wherever the reference uses the result (d2[4] < 1.0)
        for which in range(2):
            ri, rd = mo.knn(which)
            gi, gd = ctx.get_knn(which)
            assert ri.shape == gi.shape
            used = rd[:, 4] < 1.0
            used_gpu = (gi[:, 4] >= 0) & (gd[:, 4] < 1.0)
            assert np.array_equal(used, used_gpu)
            assert np.array_equal(ri[used], gi[used]), f"kNN sets differ (which={which})"
            assert np.array_equal(rd[used].view(np.uint32), gd[used].view(np.uint32))

        ori_r, co_r = mo.correspondences()
        ori_g, co_g = ctx.get_correspondences()
        assert n_corr == ori_r.shape[0] > 50
        assert_clouds_bitexact(ori_g, ori_r, "laserCloudOri")
        assert_clouds_bitexact(co_g, co_r, "coeffSel")

        A_r, B_r, X_r = mo.normal_eq()
        A_g, B_g, X_g = ctx.get_normal_equations()
        # fp64 accumulation in a different order, then one rounding to fp32: identical bits expected;
        # allow 1 ulp on an element to keep the test honest about what is guaranteed
        assert np.allclose(A_g, A_r, rtol=2e-7, atol=0) and np.allclose(B_g, B_r, rtol=2e-7, atol=1e-9)
        exact = np.array_equal(A_g, A_r) and np.array_equal(B_g, B_r)
        if exact:
            assert np.array_equal(X_g, X_r)
            assert np.array_equal(T_gpu, mo.transformTobeMapped)
            assert conv == conv_ref
        else:
            assert np.allclose(T_gpu, mo.transformTobeMapped, atol=1e-6)
        T = mo.transformTobeMapped.copy()


# ------------------------------------------------------------------ a9: the fused loop

def _ref_pose(case):
    """the compiled reference itself (tier B: unmodified mapOptmization.cpp) on the same inputs, or None"""
    from oracle import ref_harness as rh
    if not rh.available():
        return None
    rmo = rh.MapOptimization()
    rmo.set_map_raw(case["map_corner_raw"], case["map_surf_raw"])
    rmo.set_scan(case["corner"], case["surf"], case["outlier"])
    rmo.downsampleCurrentScan()
    rmo.transformTobeMapped = case["init"]
    rmo.scan2MapOptimization()
    return rmo.transformTobeMapped


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_scan2map_pose_parity(ctx, seed):
    case = data.mapping_case(seed)
    # the oracle in the reference's arithmetic (libm sinf / cosf): everything agrees to the bit
    mo, _ = _setup(ctx, case, 0)
    mo.transformTobeMapped = case["init"]
    it_ref = mo.scan2MapOptimization()
    T_gpu, st = ctx.s2m_optimize(case["init"])
    T_ref = mo.transformTobeMapped
    assert not st.skipped and st.iterations == it_ref
    assert np.array_equal(T_gpu.view(np.uint32), np.asarray(T_ref, np.float32).view(np.uint32)), (T_gpu, T_ref)
    deg_r, P_r = mo.degenerate(); deg_g, P_g = ctx.get_degeneracy()
    assert deg_r == deg_g and np.allclose(P_r, P_g, atol=1e-5)
    # north-star tolerance (kept as the documented bar)
    assert np.max(np.abs(T_gpu[:3] - T_ref[:3])) < POSE_TOL_RAD
    assert np.max(np.abs(T_gpu[3:] - T_ref[3:])) < POSE_TOL_M
    # the compiled reference itself
    T_b = _ref_pose(case)
    if T_b is not None:
        assert np.array_equal(T_gpu.view(np.uint32), T_b.view(np.uint32)), (T_gpu, T_b)
    # and it actually registered: closer to the truth than the initial guess
    err0 = np.linalg.norm(case["init"][3:] - case["pose"][3:]); err1 = np.linalg.norm(T_gpu[3:] - case["pose"][3:])
    assert err1 < err0


# ------------------------------------------------------------------ BASELINE configs 1 / 3 / 4 (sensor + map size)

@pytest.mark.parametrize("workload", ["vlp16_50k", "hdl32e_300k", "vls128_2m"])
def test_scan2map_parity_baseline_configs(workload):
    """BASELINE.json configs[0] (VLP-16 vs a 50k map), [2] (HDL-32E vs 300k), [3] (VLS-128 vs 2M) on bench.py's own generators:
    DS maps / DS scans bit-exact, first-iteration kNN sets + rows bit-exact against the restatement, final pose
    bit-identical to the restatement and to the compiled reference (tier B)."""
    import bench
    from lego_loam_b200 import api
    mc, ms, scans = bench.make_inputs(workload, 0, 1)
    sc, init = scans[0]
    case = dict(map_corner_raw=mc, map_surf_raw=ms, corner=sc.corner_last, surf=sc.surf_last, outlier=sc.outlier_last, init=init)
    ctx = api.Context(0)
    try:
        mo, counts = _setup(ctx, case, 0)
        for which in range(4):
            assert_clouds_bitexact(ctx.scan_get_ds(which), mo.scan_ds(which), f"{workload} scan ds {which}")
        for which in range(2):
            assert_clouds_bitexact(ctx.map_get_ds(which), mo.map_ds(which), f"{workload} map ds {which}")
        # one iteration: kNN sets and rows
        mo.build_kdtrees()
        mo.transformTobeMapped = init
        mo.clear_correspondences()
        mo.cornerOptimization(0); mo.surfOptimization(0)
        mo.LMOptimization(0)
        T1, conv, n_corr = ctx.s2m_iterate(init, 0)
        for which in range(2):
            ri, rd = mo.knn(which); gi, gd = ctx.get_knn(which)
            used = rd[:, 4] < 1.0
            assert np.array_equal(used, (gi[:, 4] >= 0) & (gd[:, 4] < 1.0))
            assert np.array_equal(ri[used], gi[used]) and np.array_equal(rd[used].view(np.uint32), gd[used].view(np.uint32))
        ori_r, co_r = mo.correspondences(); ori_g, co_g = ctx.get_correspondences()
        assert n_corr == ori_r.shape[0] > 50
        assert_clouds_bitexact(ori_g, ori_r, "laserCloudOri"); assert_clouds_bitexact(co_g, co_r, "coeffSel")
        # the whole registration
        mo.transformTobeMapped = init
        it_ref = mo.scan2MapOptimization()
        T_gpu, st = ctx.s2m_optimize(init)
        assert st.iterations == it_ref and not st.skipped
        assert np.array_equal(T_gpu.view(np.uint32), np.asarray(mo.transformTobeMapped, np.float32).view(np.uint32)), \
            (T_gpu, mo.transformTobeMapped)
        T_b = _ref_pose(case)
        if T_b is not None:
            assert np.array_equal(T_gpu.view(np.uint32), T_b.view(np.uint32)), (T_gpu, T_b)
        # the batched engine on the same registration (one slot)
        mc_ds, ms_ds = ctx.map_get_ds(0), ctx.map_get_ds(1)
        cap_scan = 1 << int(np.ceil(np.log2(max(sc.corner_last.shape[0], sc.surf_last.shape[0] + sc.outlier_last.shape[0], 1024))))
        if cap_scan <= 16384:                                # scan capacity of a batch slot
            b = api.Batch(0, 1, cap_scan, max(mc_ds.shape[0], ms_ds.shape[0]) + 64)
            b.scan_set(0, sc.corner_last, sc.surf_last, sc.outlier_last); b.map_set_ds(0, mc_ds, ms_ds)
            Tb, stb = b.register(np.stack([init]))
            b.close()
            assert np.array_equal(Tb[0].view(np.uint32), T_gpu.view(np.uint32)) and stb[0].iterations == st.iterations
    finally:
        ctx.close()


def test_scan2map_guard_small_map(ctx):
    """MO:1331: with <= 10 corner or <= 100 surf map points nothing runs and the pose is untouched."""
    case = data.mapping_case(1)
    ctx.map_set_ds(case["map_corner_raw"][:10], case["map_surf_raw"][:5000])
    ctx.scan_set(case["corner"], case["surf"], case["outlier"])
    ctx.downsample_current_scan()
    T, st = ctx.s2m_optimize(case["init"])
    assert st.skipped == 1 and st.iterations == 0 and np.array_equal(T, case["init"])


def test_scan2map_too_few_correspondences(ctx):
    """MO:1238 / C9: < 50 rows -> LMOptimization returns false without touching the pose, all 10 iterations run."""
    case = data.mapping_case(2)
    far = case["init"].copy(); far[3:] += 500.0              # scan nowhere near the map
    mo, _ = _setup(ctx, case, 0)
    mo.transformTobeMapped = far
    it_ref = mo.scan2MapOptimization()
    T, st = ctx.s2m_optimize(far)
    assert it_ref == 10 and st.iterations == 10 and st.converged == 0
    assert np.array_equal(T, far) and np.array_equal(mo.transformTobeMapped, far)
    oracle.set_trig_mode(0)


def test_degeneracy_persists_across_registrations(ctx):
    """C6: isDegenerate / matP are written at iteration 0 only and persist."""
    case = data.mapping_case(3)
    # a map that is a single plane (ground only) is degenerate in x, z and yaw
    ground = case["map_surf_raw"][np.abs(case["map_surf_raw"][:, 1] + 0.8) < 0.1]
    corner = case["map_corner_raw"][:12].copy()
    corner[:, :3] = 1000.0 + 3.0 * np.arange(12)[:, None]        # 12 far-away voxels: pass the guard, never match
    oracle.set_trig_mode(0)
    mo = oracle.MapOptimization()
    mo.set_map_raw(corner, ground)
    mo.set_scan(case["corner"], case["surf"], case["outlier"])
    mo.downsampleCurrentScan()
    mo.transformTobeMapped = case["init"]
    mo.scan2MapOptimization()
    ctx.map_set_raw(corner, ground)
    ctx.scan_set(case["corner"], case["surf"], case["outlier"])
    ctx.downsample_current_scan()
    T, st = ctx.s2m_optimize(case["init"])
    deg_r, P_r = mo.degenerate()
    deg_g, P_g = ctx.get_degeneracy()
    assert deg_r and deg_g and st.is_degenerate == 1
    assert np.allclose(P_r, P_g, atol=1e-4)
    assert np.allclose(T, mo.transformTobeMapped, atol=1e-5)
    oracle.set_trig_mode(0)


# ------------------------------------------------------------------ K5 odometry

def _odom_case(seed):
    w = data.world()
    rng = np.random.default_rng(50 + seed)
    pose = np.array([0.0, rng.uniform(-3, 3), 0.0, rng.uniform(-15, 15), 0.0, rng.uniform(-15, 15)])
    cur = np.array([rng.uniform(-0.004, 0.004), rng.uniform(-0.02, 0.02), rng.uniform(-0.004, 0.004),
                    rng.uniform(-0.02, 0.02), rng.uniform(-0.01, 0.01), -0.15 + rng.uniform(-0.05, 0.05)])
    return synth.make_odometry_pair(w, synth.VLP16, pose, cur, seed=seed)


@pytest.mark.parametrize("seed", [1, 2])
def test_odometry_single_steps_bitexact(ctx, seed):
    od = _odom_case(seed)
    oracle.set_trig_mode(0)
    fa = oracle.FeatureAssociation()
    fa.set_last(od.corner_last, od.surf_last, force=True)
    fa.set_features(od.corner_sharp, od.surf_flat)
    ctx.odom_set_last(od.corner_last, od.surf_last)
    ctx.odom_set_features(od.corner_sharp, od.surf_flat)
    T = np.zeros(6, np.float32)
    for which, find, calc in ((0, fa.findCorrespondingSurfFeatures, fa.calculateTransformationSurf),
                              (1, fa.findCorrespondingCornerFeatures, fa.calculateTransformationCorner)):
        for it in (0, 1, 5, 6):
            fa.transformCur = T
            fa.clear_correspondences()
            find(it)
            ori_r, co_r = fa.correspondences()
            more_ref = calc(it) if ori_r.shape[0] >= 10 else True
            T_gpu, more, n_corr = ctx.odom_iterate(which, T, it)
            i1r, i2r, i3r = fa.search_ind(1 - which)          # oracle: 0 corner, 1 surf
            i1g, i2g, i3g = ctx.odom_get_search_ind(which)        # C ABI: 0 surf, 1 corner
            assert np.array_equal(i1r, i1g) and np.array_equal(i2r, i2g)
            if which == 0:
                assert np.array_equal(i3r, i3g)
            ori_g, co_g = ctx.odom_get_correspondences()
            assert n_corr == ori_r.shape[0]
            assert_clouds_bitexact(ori_g, ori_r, "odom laserCloudOri")
            assert_clouds_bitexact(co_g, co_r, "odom coeffSel")
            assert np.array_equal(T_gpu.view(np.uint32), np.asarray(fa.transformCur, np.float32).view(np.uint32))
            assert more == more_ref
            T = fa.transformCur.copy()
    oracle.set_trig_mode(0)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_odometry_update_transformation_parity(ctx, seed):
    od = _odom_case(seed)
    oracle.set_trig_mode(0)
    fa = oracle.FeatureAssociation()
    fa.set_last(od.corner_last, od.surf_last, force=True)
    fa.set_features(od.corner_sharp, od.surf_flat)
    fa.transformCur = np.zeros(6, np.float32)
    it1, it2 = fa.updateTransformation()
    ctx.odom_set_last(od.corner_last, od.surf_last)
    ctx.odom_set_features(od.corner_sharp, od.surf_flat)
    T, s_surf, s_corner = ctx.odom_optimize(np.zeros(6, np.float32))
    assert (s_surf.iterations, s_corner.iterations) == (it1, it2)
    assert np.array_equal(T.view(np.uint32), np.asarray(fa.transformCur, np.float32).view(np.uint32)), (T, fa.transformCur)
    from oracle import ref_harness as rh
    if rh.available():                                       # the compiled reference itself (unmodified featureAssociation.cpp)
        rfa = rh.FeatureAssociation()
        rfa.set_last(od.corner_last, od.surf_last, force=True)
        rfa.set_features(od.corner_sharp, od.surf_flat)
        rfa.transformCur = np.zeros(6, np.float32)
        rfa.updateTransformation()
        assert np.array_equal(T.view(np.uint32), rfa.transformCur.view(np.uint32)), (T, rfa.transformCur)


def test_odometry_stale_trees_quirk_c20(ctx):
    """C20 (FA:1668 vs FA:1785): with exactly 10 corner points in the new last-sweep cloud updateTransformation still
    runs (>= 10 / >= 100) but the kd-trees were NOT rebuilt (> 10 / > 100): the 1-NN search runs in the PREVIOUS sweep's
    trees and its indices are used in the NEW clouds.  Constructed so that every stale index stays inside the new
    clouds (anything else is undefined behaviour in the reference): previous corner cloud = 10 real points + 1 far-away
    point, new corner cloud = 10 other points; the new surf cloud is at least as long as the previous one."""
    od_a, od_b = _odom_case(4), _odom_case(5)
    corner_a = np.vstack([od_a.corner_last[:10], np.array([[900.0, 900.0, 900.0, 3.0]], np.float32)]).astype(np.float32)
    surf_a, surf_b = od_a.surf_last, od_b.surf_last
    n = min(surf_a.shape[0], surf_b.shape[0])
    surf_a = np.ascontiguousarray(surf_a[:n]); surf_b = np.ascontiguousarray(surf_b)      # |surf_b| >= |surf_a|
    corner_b = np.ascontiguousarray(od_a.corner_last[10:20] + np.float32(0.01))            # exactly 10 points
    oracle.set_trig_mode(0)
    fa = oracle.FeatureAssociation()
    fa.set_last(corner_a, surf_a)                       # 11 / n points: trees built (FA:1785)
    fa.set_last(corner_b, surf_b)                       # 10 corner points: trees stay those of sweep A
    fa.set_features(od_a.corner_sharp, od_a.surf_flat)
    fa.transformCur = np.zeros(6, np.float32)
    it1, it2 = fa.updateTransformation()
    ctx.odom_set_last(corner_a, surf_a)
    ctx.odom_set_last(corner_b, surf_b)
    ctx.odom_set_features(od_a.corner_sharp, od_a.surf_flat)
    T, s0, s1 = ctx.odom_optimize(np.zeros(6, np.float32))
    assert s0.skipped == 0 and (s0.iterations, s1.iterations) == (it1, it2)
    assert np.array_equal(T.view(np.uint32), np.asarray(fa.transformCur, np.float32).view(np.uint32)), (T, fa.transformCur)
    # and it is NOT what fresh trees give (the quirk is observable)
    fa2 = oracle.FeatureAssociation()
    fa2.set_last(corner_b, surf_b, force=True)
    fa2.set_features(od_a.corner_sharp, od_a.surf_flat)
    fa2.transformCur = np.zeros(6, np.float32)
    fa2.updateTransformation()
    assert not np.array_equal(np.asarray(fa2.transformCur, np.float32), T)
    from oracle import ref_harness as rh
    if rh.available():                                       # the compiled reference itself
        rfa = rh.FeatureAssociation()
        rfa.set_last(corner_a, surf_a); rfa.set_last(corner_b, surf_b)
        rfa.set_features(od_a.corner_sharp, od_a.surf_flat)
        rfa.transformCur = np.zeros(6, np.float32)
        rfa.updateTransformation()
        assert np.array_equal(T.view(np.uint32), rfa.transformCur.view(np.uint32)), (T, rfa.transformCur)
    ctx.odom_set_last(od_a.corner_last, od_a.surf_last)      # leave the shared context with fresh indices


def test_odometry_guard(ctx):
    """FA:1668: fewer than 10 corner / 100 surf points in the last sweep -> nothing happens."""
    od = _odom_case(1)
    ctx.odom_set_last(od.corner_last[:9], od.surf_last)
    ctx.odom_set_features(od.corner_sharp, od.surf_flat)
    T0 = np.array([0.001, 0.002, 0.003, 0.01, 0.02, 0.03], np.float32)
    T, s0, s1 = ctx.odom_optimize(T0)
    assert s0.skipped == 1 and np.array_equal(T, T0)


def test_device_resident_family_matches_host_path(ctx):
    """llb_scan_set_dev (borrowed device sweeps) + llb_map_set_ds_dev + llb_s2m_optimize_dev give the bits of the host-cloud
    calls."""
    import torch
    case = data.mapping_case(3)
    ctx.map_set_raw(case["map_corner_raw"], case["map_surf_raw"])
    mc_ds, ms_ds = ctx.map_get_ds(0), ctx.map_get_ds(1)
    ctx.map_set_ds(mc_ds, ms_ds); ctx.scan_set(case["corner"], case["surf"], case["outlier"])
    counts = ctx.downsample_current_scan()
    T_host, st = ctx.s2m_optimize(case["init"])
    dev = torch.device("cuda", 0)
    d = {k: torch.from_numpy(np.ascontiguousarray(v, np.float32)).to(dev)
         for k, v in (("c", case["corner"]), ("s", case["surf"]), ("o", case["outlier"]), ("mc", mc_ds), ("ms", ms_ds))}
    d_T = torch.from_numpy(np.asarray(case["init"], np.float32).copy()).to(dev)
    torch.cuda.synchronize()
    ctx.scan_set_dev(d["c"].data_ptr(), d["c"].shape[0], d["s"].data_ptr(), d["s"].shape[0], d["o"].data_ptr(), d["o"].shape[0])
    assert ctx.downsample_current_scan() == counts
    ctx.map_set_ds_dev(d["mc"].data_ptr(), d["mc"].shape[0], d["ms"].data_ptr(), d["ms"].shape[0])
    ctx.s2m_optimize_dev(d_T.data_ptr())
    ctx.synchronize()
    assert np.array_equal(d_T.cpu().numpy().view(np.uint32), T_host.view(np.uint32))
