"""IMU branches of featureAssociation against the compiled reference (oracle/_ref), bit for bit:
adjustDistortion's per-point IMU interpolation + TransformToStartIMU (FA:525-613) and the IMU terms of TransformToEnd
(FA:927-950) run on the device; the ring buffers come from the reference's own imuHandler / AccumulateIMUShiftAndRotation
here, so that the test isolates the device part (the adapter's host bookkeeping is checked in test_gpu_adapter.py)."""
import numpy as np
import pytest

from lego_loam_b200 import api, synth
from oracle import ref_harness as rh

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")]


def imu_messages(n, t0, seed, yaw0=0.4, dt=0.005):
    """n messages {stamp, roll, pitch, yaw, linear acceleration, angular velocity}: a vehicle that accelerates and turns"""
    rng = np.random.default_rng(seed)
    out = []
    yaw = yaw0
    for k in range(n):
        roll = 0.02 * np.sin(0.11 * k) + 0.002 * rng.standard_normal()
        pitch = 0.03 * np.cos(0.07 * k) + 0.002 * rng.standard_normal()
        yaw = yaw + 0.004
        if yaw > np.pi:
            yaw -= 2 * np.pi                     # tf's getRPY returns yaw in (-pi, pi]: the wrap FA:549-553 handles
        la = np.array([0.3 + 0.05 * rng.standard_normal(), 0.1 * rng.standard_normal(), 9.81 + 0.05 * rng.standard_normal()])
        av = np.array([0.01 * rng.standard_normal(), 0.02 * rng.standard_normal(), 0.8 + 0.01 * rng.standard_normal()])
        out.append((t0 + dt * k, roll, pitch, yaw, la, av))
    return out


def queue_from_reference(fa, time_scan_cur, pointer_last_iteration):
    q = api.ImuQueue()
    for i in range(api.IMU_QUEUE):
        _, t, o = fa.imu_entry(i)
        q.time[i] = t
        q.roll[i], q.pitch[i], q.yaw[i] = float(o[0]), float(o[1]), float(o[2])
        for a in range(3):
            q.velo[a][i] = float(o[3 + a]); q.shift[a][i] = float(o[6 + a]); q.angular[a][i] = float(o[9 + a])
    q.time_scan_cur = time_scan_cur
    q.pointer_last = fa.imu_entry(-1)[0]
    q.pointer_last_iteration = pointer_last_iteration
    return q


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("case", ["covered", "imu_ends_mid_sweep", "yaw_wrap", "ring_wrap"])
def test_adjust_distortion_and_transform_to_end_with_imu(ctx, case):
    w = synth.make_world()
    fa = rh.FeatureAssociation()
    ctx.features_init(16, 1800)
    t0 = 1000.0
    n_msgs = {"covered": 60, "imu_ends_mid_sweep": 31, "yaw_wrap": 60, "ring_wrap": 260}[case]
    yaw0 = 3.02 if case == "yaw_wrap" else 0.4
    msgs = imu_messages(n_msgs, t0, 5, yaw0)
    for m in msgs:
        fa.push_imu(*m)
    last_iter = 0                                                      # imuPointerLastIteration FA:269
    t_scan = msgs[-1][0] - (0.06 if case == "imu_ends_mid_sweep" else 0.2)      # sweep inside / running past the messages
    prev_ang = np.zeros(3, np.float32)
    for sweep_no in range(2):
        sw = synth.make_segmented_sweep(w, synth.VLP16, [0, 0.05 + 0.1 * sweep_no, 0, 3, 0, 5 + sweep_no], 3 + sweep_no)
        fa.set_time_scan_cur(t_scan)
        fa.set_segmented(sw)
        fa.extract_features()
        st = fa.imu_state()
        q = queue_from_reference(fa, t_scan, last_iter)
        ctx.features_set_imu(q)
        counts, _ = ctx.features_extract(sw)
        got = ctx.features_get_imu()
        assert got.valid == 1 and got.has_velo == 1
        assert np.array_equal(bits(list(got.start)), bits(st[0:9]))
        assert np.array_equal(bits(list(got.cur)), bits(st[9:12]))
        assert np.array_equal(bits(list(got.velo_from_start_cur)), bits(st[12:15]))
        assert np.array_equal(bits(list(got.angular_cur)), bits(st[18:21]))
        ang_from_start = np.array(list(got.angular_cur), np.float32) - prev_ang            # FA:596-598
        assert np.array_equal(bits(ang_from_start), bits(st[15:18]))
        prev_ang = np.array(list(got.angular_cur), np.float32)
        # the de-skewed cloud and the four feature clouds
        seg_ref = fa.feature_cloud(4)
        seg_gpu = ctx.features_get(4)
        assert seg_ref.shape == seg_gpu.shape and np.array_equal(bits(seg_ref), bits(seg_gpu))
        assert not np.array_equal(seg_ref[1:, :3], np.stack([sw.cloud[1:, 1], sw.cloud[1:, 2], sw.cloud[1:, 0]], 1))   # the branch ran
        for k in range(4):
            a = fa.feature_cloud(k); b = ctx.features_get(k)
            assert a.shape[0] == counts[k] and np.array_equal(bits(a), bits(b)), k
        # updateInitialGuess + publishCloudsLast with the IMU terms of TransformToEnd
        fa.transformCur = np.array([0.002, 0.01, -0.001, 0.05, 0.01, 0.12], np.float32)
        fa.updateInitialGuess()
        st = fa.imu_state()
        T = fa.transformCur
        fa.publishCloudsLast()
        ctx.features_publish_last_imu(T, st[0:3], [0.0, 0.0, 0.0], st[21:24])
        for k in (5, 6):
            a = fa.feature_cloud(k); b = ctx.features_get(k)
            assert a.shape == b.shape and a.shape[0] > 10 and np.array_equal(bits(a), bits(b)), k
        # next sweep: more messages, the pointer walk starts at imuPointerLastIteration (FA:527, FA:616)
        last_iter = q.pointer_last
        more = imu_messages(25, msgs[-1][0] + 0.005, 6 + sweep_no, yaw0 + 0.004 * n_msgs)
        for m in more:
            fa.push_imu(*m)
        msgs = msgs + more
        t_scan = t_scan + 0.1


def test_features_without_imu_unchanged(ctx):
    """llb_features_set_imu(NULL) / a queue with imuPointerLast < 0: the branch does not run (FA:525)"""
    w = synth.make_world()
    sw = synth.make_segmented_sweep(w, synth.VLP16, [0, 0.05, 0, 3, 0, 5], 3)
    fa = rh.FeatureAssociation()
    fa.set_segmented(sw); fa.extract_features()
    ctx.features_init(16, 1800)
    q = api.ImuQueue(); q.pointer_last = -1
    ctx.features_set_imu(q)
    ctx.features_extract(sw)
    assert ctx.features_get_imu().valid == 0
    assert np.array_equal(bits(fa.feature_cloud(4)), bits(ctx.features_get(4)))
    ctx.features_set_imu(None)
    ctx.features_extract(sw)
    assert np.array_equal(bits(fa.feature_cloud(4)), bits(ctx.features_get(4)))
