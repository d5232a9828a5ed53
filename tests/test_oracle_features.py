"""Feature-extraction oracle: the C restatement (oracle/llo_features.c) against the UNMODIFIED reference
featureAssociation.cpp compiled in oracle/_ref (adjustDistortion, calculateSmoothness, markOccludedPoints,
extractFeatures, FA:491-784) and the std::sort restatement against libstdc++ itself."""
import dataclasses

import numpy as np
import pytest

import oracle
from oracle import ref_harness as rh
from lego_loam_b200 import synth

needs_ref = pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")


def sweeps(n, quantize=None, seed=3):
    w = synth.make_world()
    out = []
    for k in range(n):
        pose = [0.002 * k, 0.05 + 0.01 * k, 0.001 * k, 3 + 0.4 * k, 0, 5 + 0.2 * k]
        sw = synth.make_segmented_sweep(w, synth.VLP16, pose, seed + k)
        if quantize:
            sw = dataclasses.replace(sw, range=(np.round(sw.range / quantize) * quantize).astype(np.float32))
        out.append(sw)
    return out


@needs_ref
@pytest.mark.parametrize("n,depth", [(0, -1), (1, -1), (2, -1), (16, -1), (17, -1), (300, -1), (1000, -1), (300, 0), (300, 1),
                                     (300, 3), (1000, 2), (33, 0)])
def test_std_sort_restatement_matches_libstdcxx(n, depth):
    rng = np.random.default_rng(n * 7 + depth + 1)
    for levels in (3, 17, 1000, 0):
        for shape in ("random", "sorted", "reverse", "pipe"):
            if levels:
                v = rng.integers(0, levels, n).astype(np.float32) * np.float32(0.25)
            else:
                v = rng.random(n).astype(np.float32)
            if shape == "sorted": v = np.sort(v)
            if shape == "reverse": v = np.sort(v)[::-1].copy()
            if shape == "pipe": v = np.concatenate([np.sort(v[: n // 2]), np.sort(v[n // 2:])[::-1]])
            ind = np.arange(n)
            rv, ri = rh.std_sort(v, ind, depth)
            ov, oi = oracle.std_sort_by_value(v, ind, depth)
            assert np.array_equal(rv, ov)
            assert np.array_equal(ri.astype(np.int64), oi.astype(np.int64))     # same order among equal values
            assert np.all(np.diff(ov) >= 0)


@needs_ref
@pytest.mark.parametrize("quantize", [None, 0.02, 0.1])
def test_feature_extraction_restatement_matches_reference(quantize):
    """Five consecutive sweeps through ONE reference object and ONE restatement object (state survives between sweeps):
    every cloud bit-identical, in the reference's order; quantised ranges make equal curvatures (sort ties) common."""
    fa = rh.FeatureAssociation(); fe = oracle.FeatureExtraction(16, 1800)
    ties = 0
    for sw in sweeps(5, quantize):
        fa.set_segmented(sw); fa.extract_features()
        got = fe.extract(sw)
        for k in range(5):
            ref = fa.feature_cloud(k)
            assert ref.shape == got[k].shape, (k, ref.shape, got[k].shape)
            assert np.array_equal(ref.view(np.uint32), got[k].view(np.uint32)), k
        for a, b in zip(fa.point_state(), fe.point_state()):
            assert np.array_equal(a, b)
        curv = fe.point_state()[0]
        ties += curv.size - np.unique(curv).size
        assert got[0].shape[0] > 50 and got[2].shape[0] > 50 and got[3].shape[0] > 2000
    if quantize:
        assert ties > 1000


def golden_sweeps():
    """Inputs and reference outputs committed under tests/golden (made by tests/golden/make_ref_golden.py from the compiled
    reference): a 2-sweep sequence through one FeatureAssociation object, ranges quantised to 2 cm."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_features_golden.npz"))
    out = []
    for k in range(int(g["n_sweeps"])):
        o = g[f"in_ori_{k}"]
        sw = synth.SegmentedSweep(g[f"in_cloud_{k}"], g[f"in_start_ring_{k}"], g[f"in_end_ring_{k}"], float(o[0]), float(o[1]),
                                  float(o[2]), g[f"in_ground_{k}"], g[f"in_col_{k}"], g[f"in_range_{k}"], np.zeros((0, 4), np.float32))
        want = {n: g[f"out_{n}_{k}"] for n in ("sharp", "less_sharp", "flat", "less_flat", "adjusted_intensity", "label", "picked",
                                               "corner_last", "surf_last")}
        out.append((sw, want))
    return out, g["T"]


def assert_cloud_matches_golden(got, ref, name, xyz_tol=0.0):
    """x, y, z identical (or within xyz_tol where sin/cos flavours enter); intensity within 2 ulp (atan2 flavour)."""
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    if xyz_tol == 0.0:
        assert np.array_equal(got[:, :3].view(np.uint32), ref[:, :3].view(np.uint32)), name
    else:
        assert np.max(np.abs(got[:, :3] - ref[:, :3])) <= xyz_tol, name
    tol = 2 * np.spacing(np.maximum(np.abs(ref[:, 3]), np.float32(1)))
    assert np.all(np.abs(got[:, 3] - ref[:, 3]) <= tol), name


def test_feature_extraction_restatement_matches_committed_reference_vectors():
    """Pins the restatement without /root/reference and without oracle/_ref."""
    sweeps_, T = golden_sweeps()
    fe = oracle.FeatureExtraction(16, 1800)
    for sw, want in sweeps_:
        got = fe.extract(sw)
        for k, n in enumerate(("sharp", "less_sharp", "flat", "less_flat")):
            assert_cloud_matches_golden(got[k], want[n], n)
        assert np.all(np.abs(got[4][:, 3] - want["adjusted_intensity"]) <= 2 * np.spacing(np.maximum(np.abs(want["adjusted_intensity"]), 1)))
        curv, picked, label = fe.point_state()
        assert np.array_equal(label, want["label"].astype(np.int32)) and np.array_equal(picked, want["picked"].astype(np.int32))
        assert_cloud_matches_golden(oracle.transform_to_end(T, got[1]), want["corner_last"], "corner_last", xyz_tol=2e-5)
        assert_cloud_matches_golden(oracle.transform_to_end(T, got[3]), want["surf_last"], "surf_last", xyz_tol=2e-5)


def reference_front_end_sweeps(n, seed=40):
    """Raw sweeps (firing order, ring channel) through the UNMODIFIED reference imageProjection.cpp (oracle/_ref/libref_ip.so):
    the segmented clouds and cloud_info messages featureAssociation receives in the node."""
    w = synth.make_world()
    ip = rh.ImageProjection()
    out = []
    for k in range(n):
        cloud, ring = synth.make_raw_sweep(w, synth.VLP16, [0.002 * k, 0.05 + 0.01 * k, 0, 3 + 0.4 * k, 0, 5 + 0.1 * k], seed + k)
        out.append(ip.process(cloud, ring))
    return out


@needs_ref
def test_feature_extraction_on_reference_image_projection_output():
    """imageProjection (reference) -> feature extraction: restatement == reference, sweep after sweep through one object;
    the segmented clouds have the node's structure (ground rings thinned to every 5th column, small clusters removed)."""
    fa = rh.FeatureAssociation(); fe = oracle.FeatureExtraction(16, 1800)
    for sw in reference_front_end_sweeps(3):
        assert 5000 < sw.cloud.shape[0] < 16 * 1800 and sw.ground.sum() > 1000 and sw.outlier.shape[0] > 0
        assert abs(sw.ori_diff - 2 * np.pi) < 0.05
        fa.set_segmented(sw); fa.extract_features()
        got = fe.extract(sw)
        for k in range(5):
            ref = fa.feature_cloud(k)
            assert ref.shape == got[k].shape and np.array_equal(ref.view(np.uint32), got[k].view(np.uint32)), k
        assert got[0].shape[0] > 50 and got[2].shape[0] > 50


@needs_ref
@pytest.mark.parametrize("name", ["hdl32e", "vls128"])
def test_other_sensors_front_end_and_features_match_reference(name):
    """The reference built for HDL-32E / VLS-128 (its utility.h sensor block switched by the harness recipe): raw sweep ->
    imageProjection -> feature extraction, restatements == reference for both steps, two sweeps through the same objects."""
    if not rh.sensor_available(name):
        pytest.skip("oracle/_ref has no build for this sensor")
    sensor = synth.SENSORS[name]
    ang_res_x = 360.0 / sensor.horizon if name == "hdl32e" else 0.2            # UT:73, UT:81
    ang_res_y = 41.33 / (sensor.n_scan - 1) if name == "hdl32e" else 0.3       # UT:74, UT:82
    w = synth.make_world()
    rip = rh.ImageProjection(name); oip = oracle.ImageProjection(sensor.n_scan, sensor.horizon, ang_res_x, ang_res_y,
                                                                 sensor.ground_scan_ind)
    rfa = rh.FeatureAssociation(name); ofe = oracle.FeatureExtraction(sensor.n_scan, sensor.horizon)
    assert rip.n_scan == sensor.n_scan and rfa.n_scan() == sensor.n_scan
    for k in range(2):
        cloud, ring = synth.make_raw_sweep(w, sensor, [0, 0.05 + 0.02 * k, 0, 3 + 0.5 * k, 0, 5], 70 + k)
        a = rip.process(cloud, ring); b = oip.process(cloud, ring)
        for x, y in zip(rip.images(), oip.images()):
            assert np.array_equal(x, y)
        assert np.array_equal(a.cloud.view(np.uint32), b.cloud.view(np.uint32)) and np.array_equal(a.col, b.col)
        assert np.array_equal(a.start_ring, b.start_ring) and np.array_equal(a.end_ring, b.end_ring)
        assert np.array_equal(a.ground, b.ground) and np.array_equal(a.range.view(np.uint32), b.range.view(np.uint32))
        rfa.set_segmented(a); rfa.extract_features()
        got = ofe.extract(a)
        for which in range(5):
            ref = rfa.feature_cloud(which, cap=sensor.n_scan * sensor.horizon)
            assert ref.shape == got[which].shape and np.array_equal(ref.view(np.uint32), got[which].view(np.uint32)), which
        assert got[0].shape[0] > sensor.n_scan and got[3].shape[0] > 1000
