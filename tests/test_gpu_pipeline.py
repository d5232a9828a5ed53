"""cfg2 (BASELINE configs[1] as the nodes run it): raw VLP-16 sweeps -> imageProjection -> featureAssociation ->
mapOptimization for >= 200 sweeps, twice over the same synthetic drive:
  * reference arm: the compiled, unmodified reference classes (oracle/_ref) for every step;
  * device arm: the SAME node bookkeeping (the reference's own integrateTransformation, key-frame store, iSAM2
    stand-in, transformAssociateToMap) with every data-parallel step through the C ABI on the GPU — projection +
    segmentation (llb_projection_*), adjustDistortion .. extractFeatures, updateTransformation (llb_odom_optimize),
    publishCloudsLast's TransformToEnd + last-cloud index, the map voxel tail, downsampleCurrentScan, scan2MapOptimization.
Bars of the north star: per-scan pose within 1e-4 m / 1e-4 rad, accumulated drift within 0.1 % of the path."""
import json
import os
import time

import numpy as np
import pytest

from lego_loam_b200 import api, synth
from oracle import ref_harness as rh

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")]

SWEEP_PERIOD = 0.1            # scanPeriod UT:107
MAPPING_INTERVAL = 0.3        # mappingProcessInterval UT:131


def make_drive(n):
    """a car at 1.5 m/s on a gentle curve: pose of every sweep (rx, ry, rz, x, y, z in the camera convention)"""
    w = synth.make_world()
    poses, sweeps = [], []
    yaw, x, z = 0.3, -6.0, -8.0
    for k in range(n):
        yaw += 0.004
        x += 0.15 * np.sin(yaw); z += 0.15 * np.cos(yaw)
        pose = np.array([0.004 * np.sin(0.05 * k), yaw, 0.004 * np.cos(0.04 * k), x, 0.0, z])
        poses.append(pose)
        sweeps.append(synth.make_raw_sweep(w, synth.VLP16, pose, 5000 + k))
    return np.array(poses), sweeps


def adjust_outlier(o):
    """adjustOutlierCloud FA:1746-1757"""
    out = o.copy(); out[:, 0] = o[:, 1]; out[:, 1] = o[:, 2]; out[:, 2] = o[:, 0]
    return out


class Mapping:
    """mapOptimization::run MO:1493-1524 around the reference's key-frame store; ctx: the hot path goes to the device"""

    def __init__(self, ctx):
        self.mo = rh.MapOptimization(); self.ctx = ctx; self.t_last = -1.0; self.traj = []; self.used = 0; self.ds_equal = True
        self.lat_ms = []                                                  # device arm: wall clock of the registration calls

    def feed(self, stamp, transform_sum, corner, surf, outlier):
        if stamp - self.t_last < MAPPING_INTERVAL:                    # MO:1499
            return
        self.t_last = stamp
        mo, ctx = self.mo, self.ctx
        mo.set_odometry(transform_sum, stamp)
        mo.set_scan(corner, surf, outlier)
        mo.transformAssociateToMap()
        mo.extractSurroundingKeyFrames()
        mo.downsampleCurrentScan()
        if ctx is None:
            mo.scan2MapOptimization()
        else:
            nc, ns = mo.map_ds_sizes()
            if nc > 10 and ns > 100:
                ctx.map_set_raw(mo.map_raw(0), mo.map_raw(1))
                self.ds_equal &= np.array_equal(ctx.map_get_ds(0).view(np.uint32), mo.map_ds(0).view(np.uint32))
                self.ds_equal &= np.array_equal(ctx.map_get_ds(1).view(np.uint32), mo.map_ds(1).view(np.uint32))
                t0 = time.perf_counter()
                ctx.scan_set(corner, surf, outlier)
                ctx.downsample_current_scan()
                T, _ = ctx.s2m_optimize(mo.transformTobeMapped)
                self.lat_ms.append((time.perf_counter() - t0) * 1e3)  # host sweep clouds in -> pose out (DS map resident)
                mo.transformTobeMapped = T
                mo.transformUpdate()
                self.used += 1
        mo.saveKeyFramesAndFactor(); mo.correctPoses(); mo.clearCloud()
        self.traj.append(mo.transformAftMapped.copy())


def run_reference(sweeps):
    rip, rfa, mp = rh.ImageProjection(), rh.FeatureAssociation(), Mapping(None)
    odo = []
    frame = 1                                                          # frameCount = skipFrameNum FA:314
    for k, (cloud, ring) in enumerate(sweeps):
        sw = rip.process(cloud, ring)
        rfa.set_segmented(sw); rfa.extract_features()
        if k == 0:                                                     # checkSystemInitialization FA:1604-1637
            rfa.set_last(rfa.feature_cloud(1), rfa.feature_cloud(3), force=True)
            odo.append(rfa.transformSum.copy())
            continue
        rfa.updateInitialGuess(); rfa.updateTransformation(); rfa.integrateTransformation(); rfa.publishCloudsLast()
        odo.append(rfa.transformSum.copy())
        frame += 1
        if frame >= 2:                                                 # skipFrameNum + 1, FA:1790
            frame = 0
            mp.feed(SWEEP_PERIOD * k, rfa.transformSum, rfa.feature_cloud(5), rfa.feature_cloud(6), adjust_outlier(sw.outlier))
    return np.array(odo), np.array(mp.traj), mp.mo.num_keyframes()


def run_device(sweeps):
    fctx = api.Context(0); fctx.projection_init(16, 1800, 0.2, 2.0, 7); fctx.features_init(16, 1800)
    mctx = api.Context(0)
    book = rh.FeatureAssociation()                                     # host bookkeeping only: integrateTransformation
    mp = Mapping(mctx)
    odo = []
    frame = 1
    T = np.zeros(6, np.float32)
    fa_ms = []                                                         # per sweep: raw sweep in -> odometry pose out + last clouds
    try:
        for k, (cloud, ring) in enumerate(sweeps):
            t0 = time.perf_counter()
            fctx.projection_process(cloud, ring)
            fctx.projection_to_features()
            if k == 0:
                fctx.odom_set_last(fctx.features_get(1), fctx.features_get(3))
                odo.append(book.transformSum.copy())
                continue
            fctx.features_to_odometry()
            T, _, _ = fctx.odom_optimize(T)                            # transformCur carries over as the initial guess
            book.transformCur = T; book.integrateTransformation()
            fctx.features_publish_last(T)
            fctx.synchronize()
            fa_ms.append((time.perf_counter() - t0) * 1e3)
            odo.append(book.transformSum.copy())
            frame += 1
            if frame >= 2:
                frame = 0
                mp.feed(SWEEP_PERIOD * k, book.transformSum, fctx.features_get(5), fctx.features_get(6),
                        adjust_outlier(fctx.projection_get_cloud(1)))
    finally:
        fctx.close(); mctx.close()
    return np.array(odo), np.array(mp.traj), mp.mo.num_keyframes(), mp.used, mp.ds_equal, np.array(fa_ms), np.array(mp.lat_ms)


def test_full_pipeline_replay_200_sweeps():
    n = int(os.environ.get("LLB_PIPELINE_SWEEPS", "200"))
    poses, sweeps = make_drive(n)
    odo_ref, map_ref, kf_ref = run_reference(sweeps)
    odo_gpu, map_gpu, kf_gpu, used, ds_equal, fa_ms, reg_ms = run_device(sweeps)
    assert odo_ref.shape == odo_gpu.shape == (n, 6) and map_ref.shape == map_gpu.shape and map_ref.shape[0] >= n // 4 - 2
    assert used >= map_ref.shape[0] - 2 and kf_ref == kf_gpu and ds_equal
    d_odo = np.abs(odo_gpu - odo_ref); d_map = np.abs(map_gpu - map_ref)
    path = float(np.sum(np.linalg.norm(np.diff(map_ref[:, 3:], axis=0), axis=1)))
    drift = float(np.linalg.norm(map_gpu[-1, 3:] - map_ref[-1, 3:]) / path)
    truth = float(np.linalg.norm(poses[-1][3:] - poses[0][3:]))       # the nodes' origin is the first sweep's pose
    rec = {"sweeps": n, "mapping_registrations": int(map_ref.shape[0]), "key_frames": int(kf_ref), "path_m": path,
           "odometry_max_abs_diff": float(d_odo.max()), "mapping_max_abs_rot_diff_rad": float(d_map[:, :3].max()),
           "mapping_max_abs_trans_diff_m": float(d_map[:, 3:].max()), "drift_vs_reference_over_path": drift,
           "odometry_sweeps_bit_identical": int(np.sum(np.all(odo_gpu.view(np.uint32) == odo_ref.view(np.uint32), axis=1))),
           "mapping_poses_bit_identical": int(np.sum(np.all(map_gpu.view(np.uint32) == map_ref.view(np.uint32), axis=1))),
           "latency_ms_per_sweep_front_end_and_odometry": {"p50": float(np.percentile(fa_ms, 50)), "p99": float(np.percentile(fa_ms, 99)),
                                                           "max": float(fa_ms.max()), "what": "raw sweep (host) in -> projection, features, "
                                                           "updateTransformation, publishCloudsLast on the device -> pose, wall clock"},
           "latency_ms_per_mapping_registration": {"p50": float(np.percentile(reg_ms, 50)), "p99": float(np.percentile(reg_ms, 99)),
                                                   "max": float(reg_ms.max()), "what": "host sweep clouds in -> downsampleCurrentScan + "
                                                   "scan2MapOptimization -> pose, wall clock (local map already voxel-filtered + indexed)"},
           "distance_from_start_m": {"truth": truth, "device": float(np.linalg.norm(map_gpu[-1, 3:])),
                                     "reference": float(np.linalg.norm(map_ref[-1, 3:]))}}
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        json.dump(rec, open(os.path.join(out_dir, "pipeline_replay.json"), "w"), indent=1)
    assert path > 0.12 * n
    assert d_odo[:, :3].max() < 1e-4 and d_odo[:, 3:].max() < 1e-4, rec     # accumulated odometry, every sweep
    assert d_map[:, :3].max() < 1e-4 and d_map[:, 3:].max() < 1e-4, rec     # per-scan bars of the north star
    assert drift < 1e-3, rec
    if os.environ.get("LLB_REQUIRE_BITEXACT", "1") != "0":            # observed on the B200 box: every pose, bit for bit
        assert rec["odometry_sweeps_bit_identical"] == n and rec["mapping_poses_bit_identical"] == map_ref.shape[0], rec
    assert abs(np.linalg.norm(map_gpu[-1, 3:]) - truth) < 0.02 * truth + 0.3, rec       # and the drive is tracked
