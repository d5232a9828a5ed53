"""The C++ adapter classes (reference member-function names over the C ABI) against the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle
from tests import data
from tests.test_gpu_parity import _odom_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_adapter_test(tmp_path):
    exe = str(tmp_path / "host_adapter_test")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-o", exe, os.path.join(ROOT, "tests", "host_adapter_test.cpp"),
                           "-L", os.path.join(ROOT, "lego_loam_b200"), "-lllb200",
                           "-Wl,-rpath," + os.path.join(ROOT, "lego_loam_b200")])
    return exe


def test_adapter_compiles_without_gpu(tmp_path):
    build_adapter_test(tmp_path)


@pytest.mark.gpu
def test_adapter_matches_oracle(tmp_path):
    exe = build_adapter_test(tmp_path)
    case = data.mapping_case(2)
    od = _odom_case(2)
    path = str(tmp_path / "case.bin")
    with open(path, "wb") as f:
        def w(c):
            c = np.ascontiguousarray(c, np.float32)
            f.write(struct.pack("i", c.shape[0])); f.write(c.tobytes())
        for c in (case["map_corner_raw"], case["map_surf_raw"], case["corner"], case["surf"], case["outlier"]):
            w(c)
        f.write(np.ascontiguousarray(case["init"], np.float32).tobytes())
        rng = np.random.default_rng(9)
        imu = np.stack([100.0 + 0.005 * np.arange(40), 0.02 * rng.standard_normal(40), 0.03 * rng.standard_normal(40)], 1)
        t_odo = 100.0 + 0.005 * 17.3 - float(np.float32(0.1))     # the blend interpolates at timeLaserOdometry + scanPeriod
        t_sum = np.array([0.01, 0.5, -0.02, 4.0, 0.1, -3.0], np.float32)
        f.write(struct.pack("i", imu.shape[0])); f.write(np.ascontiguousarray(imu, np.float64).tobytes())
        f.write(struct.pack("d", t_odo)); f.write(t_sum.tobytes())
        for c in (od.corner_last, od.surf_last, od.corner_sharp, od.surf_flat):
            w(c)
        from lego_loam_b200 import synth
        sw = synth.make_segmented_sweep(synth.make_world(), synth.VLP16, [0, 0.05, 0, 3, 0, 5], 3)
        f.write(struct.pack("ii", 16, 1800)); w(sw.cloud)
        f.write(sw.start_ring.astype(np.int32).tobytes()); f.write(sw.end_ring.astype(np.int32).tobytes())
        f.write(struct.pack("fff", sw.start_ori, sw.end_ori, sw.ori_diff))
        f.write(sw.ground.astype(np.uint8).tobytes()); f.write(sw.col.astype(np.uint32).tobytes())
        f.write(sw.range.astype(np.float32).tobytes())
        # IMU block: two sweeps, each preceded by IMU messages (the second block wraps the 200-entry ring)
        from tests.test_gpu_imu import imu_messages
        imu_sweeps = []
        t_msg, yaw0 = 2000.0, 3.02
        f.write(struct.pack("i", 2))
        for s_no, n_msg in enumerate((60, 170)):
            msgs = imu_messages(n_msg, t_msg, 20 + s_no, yaw0)
            t_msg = msgs[-1][0] + 0.005; yaw0 = msgs[-1][3] + 0.004
            stamp = msgs[-1][0] - 0.15
            swi = synth.make_segmented_sweep(synth.make_world(), synth.VLP16, [0, 0.05 + 0.1 * s_no, 0, 3, 0, 5 + s_no], 5 + s_no)
            t_cur = np.array([0.002, 0.01, -0.001, 0.05, 0.01, 0.12], np.float32) * (1 + s_no)
            f.write(struct.pack("i", n_msg))
            for m in msgs:
                f.write(struct.pack("10d", m[0], m[1], m[2], m[3], *m[4], *m[5]))
            f.write(struct.pack("d", stamp)); w(swi.cloud)
            f.write(swi.start_ring.astype(np.int32).tobytes()); f.write(swi.end_ring.astype(np.int32).tobytes())
            f.write(struct.pack("fff", swi.start_ori, swi.end_ori, swi.ori_diff))
            f.write(swi.ground.astype(np.uint8).tobytes()); f.write(swi.col.astype(np.uint32).tobytes())
            f.write(swi.range.astype(np.float32).tobytes())
            f.write(t_cur.tobytes())
            imu_sweeps.append((msgs, stamp, swi, t_cur))
        # loop closure: a source cloud = part of the target moved by a small rigid motion
        rng = np.random.default_rng(12)
        lc_tgt = np.zeros((12000, 4), np.float32); lc_tgt[:, :3] = rng.uniform(-25, 25, (12000, 3)) * [1, 0.1, 1]
        a = 0.02; Rm = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        lc_src = lc_tgt[rng.choice(12000, 2500, replace=False)].copy()
        lc_src[:, :3] = (lc_src[:, :3] @ Rm.T + np.array([0.25, 0.02, -0.2])).astype(np.float32)
        w(lc_src); w(lc_tgt)
    out = subprocess.run([exe, path], capture_output=True, text=True, check=True).stdout.splitlines()
    mo_line = out[0].split(); tu_line = out[1].split(); fa_line = out[2].split()

    oracle.set_trig_mode(0)
    mo = oracle.MapOptimization()
    mo.set_map_raw(case["map_corner_raw"], case["map_surf_raw"])
    mo.set_scan(case["corner"], case["surf"], case["outlier"])
    mo.downsampleCurrentScan()
    mo.transformTobeMapped = case["init"]
    iters = mo.scan2MapOptimization()
    assert [int(x) for x in mo_line[1:5]] == [mo.scan_ds(i).shape[0] for i in range(4)]
    assert int(mo_line[5]) == iters
    assert np.allclose(np.array(mo_line[7:13], np.float32), mo.transformTobeMapped, atol=1e-6)
    assert int(mo_line[13]) == mo.scan_ds(3).shape[0]

    # transformUpdate with IMU messages against the compiled reference (MO:463-496), bit for bit
    from oracle import ref_harness as rh
    if rh.available():
        rmo = rh.MapOptimization()
        for st, ro, pi in imu:
            rmo.push_imu(st, ro, pi)
        rmo.set_odometry(t_sum, t_odo)
        rmo.transformTobeMapped = np.array(mo_line[7:13], np.float32)
        rmo.transformUpdate()
        bef, aft = rmo.bef_aft()
        assert tu_line[0] == "TU"
        got = np.array(tu_line[1:13], np.float32)
        assert np.array_equal(got[:6].view(np.uint32), bef.view(np.uint32)) and np.array_equal(got[6:].view(np.uint32), aft.view(np.uint32)), (got, bef, aft)
        assert not np.array_equal(aft, np.array(mo_line[7:13], np.float32))        # the blend changed roll / pitch

    fa = oracle.FeatureAssociation()
    fa.set_last(od.corner_last, od.surf_last, force=True)
    fa.set_features(od.corner_sharp, od.surf_flat)
    fa.transformCur = np.zeros(6, np.float32)
    it1, it2 = fa.updateTransformation()
    assert (int(fa_line[1]), int(fa_line[2])) == (it1, it2)
    assert np.allclose(np.array(fa_line[3:9], np.float32), fa.transformCur, atol=1e-5)
    oracle.set_trig_mode(0)

    # extractFeatures through the adapter: sizes and the x, y, z words of the four clouds (FNV-1a) as the oracle has them
    fe_line = out[3].split()
    assert fe_line[0] == "FE" and int(fe_line[1]) == 0
    want = oracle.FeatureExtraction(16, 1800).extract(sw)

    def fnv(cloud):
        h = 2166136261
        for wd in np.ascontiguousarray(cloud[:, :3], np.float32).view(np.uint32).ravel().tolist():
            h = ((h ^ wd) * 16777619) & 0xFFFFFFFF
        return h
    for k in range(4):
        assert int(fe_line[2 + 2 * k]) == want[k].shape[0]
        assert int(fe_line[3 + 2 * k]) == fnv(want[k])

    # IMU branches through the adapter (host ring buffers + device per-point work) against the compiled reference
    if rh.available():
        def fnv4(cloud):
            h = 2166136261
            for wd in np.ascontiguousarray(cloud, np.float32).view(np.uint32).ravel().tolist():
                h = ((h ^ wd) * 16777619) & 0xFFFFFFFF
            return h
        rfa = rh.FeatureAssociation()
        for s_no, (msgs, stamp, swi, t_cur) in enumerate(imu_sweeps):
            for m in msgs:
                rfa.push_imu(*m)
            rfa.set_time_scan_cur(stamp); rfa.set_segmented(swi); rfa.extract_features()
            rfa.transformCur = t_cur
            rfa.updateInitialGuess()
            line = out[4 + s_no].split()
            assert line[0] == "IM" and int(line[1]) == 0
            got = np.array([int(x) for x in line[2:32]], np.uint32)
            want = np.concatenate([rfa.imu_state(), rfa.transformCur]).view(np.uint32)
            assert np.array_equal(got, want), (s_no, got.view(np.float32), want.view(np.float32))
            seg = rfa.feature_cloud(4)
            assert (int(line[32]), int(line[33])) == (seg.shape[0], fnv4(seg))
            rfa.publishCloudsLast()
            cl, sl = rfa.feature_cloud(5), rfa.feature_cloud(6)
            assert (int(line[34]), int(line[35]), int(line[36]), int(line[37])) == (cl.shape[0], fnv4(cl), sl.shape[0], fnv4(sl))

    # loop-closure ICP through the adapter against the restated PCL algorithm
    lc = [l for l in out if l.startswith("LC ")][0].split()
    want_lc = oracle.icp_align(lc_src, lc_tgt)
    assert int(lc[1]) == 0 and int(lc[2]) == 1 and int(lc[3]) == want_lc["iterations"] and int(lc[4]) == want_lc["state"]
    assert abs(float(lc[5]) - want_lc["fitness"]) < 1e-9
    assert np.abs(np.array(lc[6:22], np.float32).reshape(4, 4) - want_lc["T"]).max() < 1e-6
    # loop clouds from the adapter's device key-frame store: the latest cloud is cornerDS + surfDS of the sweep
    lk = [l for l in out if l.startswith("LK ")][0].split()
    assert int(lk[1]) == 0 and int(lk[2]) == 1 and int(lk[3]) == int(mo_line[1]) + int(mo_line[2]) and 0 < int(lk[4]) <= int(lk[3])
