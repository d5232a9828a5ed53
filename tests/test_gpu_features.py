"""Feature extraction on the device (adjustDistortion, calculateSmoothness, markOccludedPoints,
extractFeatures, FA:491-784) through the C ABI against the oracle restatement and, where oracle/_ref travelled with the
snapshot, against the UNMODIFIED reference featureAssociation.cpp.

Bar: which points are picked, their order, labels, flags, curvatures and coordinates bit-identical (std::sort's order
of equal curvatures included); the relative-time part of the intensity goes through glibc's atan2f, which the kernel
restates operation by operation: observed bit-identical, asserted within 2 ulp of the intensity value so that a host
with another libm does not fail the suite."""
import dataclasses

import numpy as np
import pytest

import oracle
from oracle import ref_harness as rh
from lego_loam_b200 import api, synth

pytestmark = pytest.mark.gpu

INT_ULPS = 2      # of the intensity value (ring index + time): 2.4e-7 on ring 1 ... 1.5e-5 on ring 127


def sweeps(n, sensor=synth.VLP16, quantize=None, seed=3):
    w = synth.make_world()
    out = []
    for k in range(n):
        pose = [0.002 * k, 0.05 + 0.01 * k, 0.001 * k, 3 + 0.4 * k, 0, 5 + 0.2 * k]
        sw = synth.make_segmented_sweep(w, sensor, pose, seed + k)
        if quantize:
            sw = dataclasses.replace(sw, range=(np.round(sw.range / quantize) * quantize).astype(np.float32))
        out.append(sw)
    return out


def check_cloud(got, ref, name):
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    assert np.array_equal(got[:, :3].view(np.uint32), ref[:, :3].view(np.uint32)), name
    if got.size:
        d = np.abs(got[:, 3] - ref[:, 3]); tol = INT_ULPS * np.spacing(np.maximum(np.abs(ref[:, 3]), np.float32(1)))
        assert np.all(d <= tol), (name, float(d.max()))
    return int(np.count_nonzero(got[:, 3].view(np.uint32) != ref[:, 3].view(np.uint32)))


@pytest.mark.parametrize("quantize", [None, 0.02, 0.1])
def test_features_match_oracle_and_reference(quantize):
    ctx = api.Context(0); ctx.features_init(16, 1800)
    fe = oracle.FeatureExtraction(16, 1800)
    fa = rh.FeatureAssociation() if rh.available() else None
    names = ["cornerPointsSharp", "cornerPointsLessSharp", "surfPointsFlat", "surfPointsLessFlat", "segmentedCloud"]
    ulp_diffs = 0; total = 0
    for sw in sweeps(6, quantize=quantize):
        counts, ms = ctx.features_extract(sw)
        want = fe.extract(sw)
        if fa is not None:
            fa.set_segmented(sw); fa.extract_features()
        for k in range(5):
            got = ctx.features_get(k)
            if k < 4:
                assert counts[k] == got.shape[0]
            ulp_diffs += check_cloud(got, want[k], names[k]); total += got.shape[0]
            if fa is not None:
                check_cloud(got, fa.feature_cloud(k), names[k] + " (reference)")
        n = sw.cloud.shape[0]
        for a, b, nm in zip(ctx.features_get_state(n), fe.point_state(), ("cloudCurvature", "cloudNeighborPicked", "cloudLabel")):
            assert np.array_equal(a, b), nm
        assert counts[0] > 50 and counts[2] > 50 and counts[3] > 2000
        print(f"n={n} counts={counts} device_ms={ms:.3f}")
    print(f"intensity words differing from glibc atan2f path: {ulp_diffs} of {total}")
    # the kernel restates glibc's atan2f operation by operation (csrc/glibc_atan2f.cuh): on a host with that libm the
    # intensities are bit-identical too (the 2-ulp bar above stays for hosts whose libm computes atan2f differently)
    import json, os
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/features_intensity_diff_{quantize}.json", "w") as f:
        json.dump({"differing_words": ulp_diffs, "of": total}, f)
    ctx.close()


def test_features_other_sensors_and_edge_cases():
    """HDL-32E / VLS-128 ring counts against the oracle restatement; an empty sweep; a sweep whose first rings are empty."""
    for sensor, nsw in ((synth.HDL32E, 2), (synth.VLS128, 2)):
        ctx = api.Context(0); ctx.features_init(sensor.n_scan, sensor.horizon)
        fe = oracle.FeatureExtraction(sensor.n_scan, sensor.horizon)
        for sw in sweeps(nsw, sensor=sensor, quantize=0.02):
            counts, _ = ctx.features_extract(sw)
            want = fe.extract(sw)
            for k in range(5):
                check_cloud(ctx.features_get(k), want[k], f"{sensor.name} {k}")
        ctx.close()
    ctx = api.Context(0); ctx.features_init(16, 1800)
    fe = oracle.FeatureExtraction(16, 1800)
    sw = sweeps(1)[0]
    # rings 0..5 empty: the cloud starts with ring 6 (ring bounds as IP:318 / IP:358 would leave them)
    first = int(sw.start_ring[6]) - 4
    sr = sw.start_ring - first; er = sw.end_ring - first
    sr[:6] = 4; er[:6] = -6
    cut = dataclasses.replace(sw, cloud=sw.cloud[first:], ground=sw.ground[first:], col=sw.col[first:], range=sw.range[first:],
                              start_ring=sr.astype(np.int32), end_ring=er.astype(np.int32))
    for s in (cut, sw, cut):
        counts, _ = ctx.features_extract(s); want = fe.extract(s)
        for k in range(5):
            check_cloud(ctx.features_get(k), want[k], f"cut {k}")
    empty = dataclasses.replace(sw, cloud=sw.cloud[:0], ground=sw.ground[:0], col=sw.col[:0], range=sw.range[:0],
                                start_ring=np.full(16, 4, np.int32), end_ring=np.full(16, -6, np.int32))
    counts, _ = ctx.features_extract(empty)
    assert counts == [0, 0, 0, 0]
    ctx.close()


def test_features_guards():
    ctx = api.Context(0)
    sw = sweeps(1)[0]
    with pytest.raises(api.LlbError):
        ctx.features_extract(sw)                      # not initialised
    ctx.features_init(16, 1800)
    bad = dataclasses.replace(sw, end_ring=(sw.end_ring + 100).astype(np.int32))
    with pytest.raises(api.LlbError):
        ctx.features_extract(bad)                     # ring bounds outside the cloud
    with pytest.raises(api.LlbError):
        ctx.features_get(0)
    ctx.close()


def test_features_feed_odometry_on_device():
    """cornerPointsSharp / surfPointsFlat handed to the odometry matcher without leaving the device give the pose the
    host round trip gives."""
    ctx = api.Context(0); ctx.features_init(16, 1800)
    a, b = sweeps(2)
    ctx.features_extract(a)
    last_c = ctx.features_get(1); last_s = ctx.features_get(3)
    ctx.odom_set_last(last_c, last_s)
    ctx.features_extract(b)
    sharp = ctx.features_get(0); flat = ctx.features_get(2)
    ctx.features_to_odometry()
    T1, s0, s1 = ctx.odom_optimize(np.zeros(6, np.float32))
    ctx.odom_set_last(last_c, last_s)
    ctx.odom_set_features(sharp, flat)
    T2, _, _ = ctx.odom_optimize(np.zeros(6, np.float32))
    assert np.array_equal(T1, T2)
    assert s0.iterations > 0
    ctx.close()


def test_device_resident_odometry_loop_matches_reference_functions():
    """The FA side of one node cycle without leaving the device - extractFeatures -> updateTransformation ->
    publishCloudsLast (TransformToEnd + last clouds + index) - for a 4-sweep sequence, against the same statements of
    the oracle restatement (correctly rounded trig; and of the compiled reference where oracle/_ref is present):
    last clouds within 5e-5 m of the restatement (libm mode) and of the reference (observed: bit-identical, see the JSON the
    test writes), poses within 1e-6 of the restatement."""
    ctx = api.Context(0); ctx.features_init(16, 1800)
    fe = oracle.FeatureExtraction(16, 1800)
    ofa = oracle.FeatureAssociation()
    rfa = rh.FeatureAssociation() if rh.available() else None
    oracle.set_trig_mode(1)
    try:
        have_last = False
        exact, exact_ref = [], []
        for i, sw in enumerate(sweeps(4)):
            ctx.features_extract(sw)
            want = fe.extract(sw)
            if rfa is not None:
                rfa.set_segmented(sw); rfa.extract_features()
            T = np.zeros(6, np.float32)
            if have_last:
                ctx.features_to_odometry()
                T, s0, s1 = ctx.odom_optimize(np.zeros(6, np.float32))
                # the restatement gets the device's own feature clouds: their intensities differ from the CPU's in the last
                # place for ~0.2 % of the points (atan2), everything else is identical (tests above)
                ofa.set_features(ctx.features_get(0), ctx.features_get(2)); ofa.transformCur = np.zeros(6, np.float32)
                ofa.updateTransformation()
                assert np.max(np.abs(T - ofa.transformCur)) <= 1e-6, (i, T, ofa.transformCur)
                assert s0.iterations > 0
            ctx.features_publish_last(T)
            # TransformToEnd takes glibc's sinf / cosf restated on the device: compare with the restatement in libm mode
            oracle.set_trig_mode(0)
            oc = oracle.transform_to_end(T, ctx.features_get(1)); os_ = oracle.transform_to_end(T, ctx.features_get(3))
            oracle.set_trig_mode(1)
            gc, gs = ctx.features_get(5), ctx.features_get(6)
            exact.append(bool(np.array_equal(gc.view(np.uint32), oc.view(np.uint32)) and np.array_equal(gs.view(np.uint32), os_.view(np.uint32))))
            assert np.max(np.abs(gc - oc)) <= 5e-5 and np.max(np.abs(gs - os_)) <= 5e-5, i
            if rfa is not None:                  # the reference's own publishCloudsLast with the device's pose
                rfa.transformCur = T; rfa.publishCloudsLast()
                rc, rs = rfa.feature_cloud(5), rfa.feature_cloud(6)
                assert rc.shape == gc.shape and rs.shape == gs.shape
                assert np.max(np.abs(rc - gc)) <= 5e-5 and np.max(np.abs(rs - gs)) <= 5e-5
                exact_ref.append(bool(np.array_equal(rc.view(np.uint32), gc.view(np.uint32)) and np.array_equal(rs.view(np.uint32), gs.view(np.uint32))))
            ofa.set_last(oc, os_, force=True)
            have_last = True
    finally:
        oracle.set_trig_mode(0)
    ctx.close()
    # observed on the B200 box: every sweep bit-identical to the restatement in libm mode AND to the compiled reference;
    # the asserted bar stays 5e-5 m for hosts whose libm computes sinf / cosf differently (no FMA variant)
    import json, os
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/features_to_end_exact.json", "w") as f:
        json.dump({"bit_identical_to_restatement_libm_mode": exact, "bit_identical_to_compiled_reference": exact_ref}, f)


def test_features_match_committed_reference_vectors():
    """The device path against the golden vectors made from the compiled reference (tests/golden/ref_features_golden.npz)."""
    from tests.test_oracle_features import golden_sweeps, assert_cloud_matches_golden
    sweeps_, T = golden_sweeps()
    ctx = api.Context(0); ctx.features_init(16, 1800)
    for sw, want in sweeps_:
        ctx.features_extract(sw)
        for k, n in enumerate(("sharp", "less_sharp", "flat", "less_flat")):
            assert_cloud_matches_golden(ctx.features_get(k), want[n], n)
        curv, picked, label = ctx.features_get_state(sw.cloud.shape[0])
        assert np.array_equal(label, want["label"].astype(np.int32)) and np.array_equal(picked, want["picked"].astype(np.int32))
        ctx.features_publish_last(T)
        assert_cloud_matches_golden(ctx.features_get(5), want["corner_last"], "corner_last", xyz_tol=2e-5)
        assert_cloud_matches_golden(ctx.features_get(6), want["surf_last"], "surf_last", xyz_tol=2e-5)
    ctx.close()


def test_batch_features_equal_single_contexts():
    """One sweep per slot through the batched form (one set of five launches for all slots) == each sequence through its
    own context, over 3 steps (every slot keeps its own state); slots see different sweeps, one slot an empty one."""
    import dataclasses as dc
    S = 5
    seqs = [sweeps(3, quantize=(0.02 if s % 2 else None), seed=3 + 10 * s) for s in range(S - 1)]
    sw0 = seqs[0][0]
    empty = dc.replace(sw0, cloud=sw0.cloud[:0], ground=sw0.ground[:0], col=sw0.col[:0], range=sw0.range[:0],
                       start_ring=np.full(16, 4, np.int32), end_ring=np.full(16, -6, np.int32))
    seqs.append([empty, seqs[1][1], empty])
    b = api.Batch(0, S, 4096, 4096); b.features_init(16, 1800)
    ctxs = [api.Context(0) for _ in range(S)]
    for c in ctxs:
        c.features_init(16, 1800)
    for step in range(3):
        counts, ms = b.features_extract(b.features_pack([seqs[s][step] for s in range(S)]))
        for s in range(S):
            cs, _ = ctxs[s].features_extract(seqs[s][step])
            assert list(counts[s]) == cs, (step, s)
            for k in range(4):
                assert np.array_equal(b.features_get(s, k).view(np.uint32), ctxs[s].features_get(k).view(np.uint32)), (step, s, k)
    for c in ctxs:
        c.close()
    b.close()


@pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")
def test_features_on_reference_image_projection_output():
    """Raw sweeps -> the reference's own imageProjection (oracle/_ref) -> feature extraction on the device vs the
    reference's featureAssociation: the node's real data path in front of the device kernels."""
    from tests.test_oracle_features import reference_front_end_sweeps
    ctx = api.Context(0); ctx.features_init(16, 1800)
    fa = rh.FeatureAssociation()
    for sw in reference_front_end_sweeps(3):
        counts, _ = ctx.features_extract(sw)
        fa.set_segmented(sw); fa.extract_features()
        for k in range(5):
            check_cloud(ctx.features_get(k), fa.feature_cloud(k), f"cloud {k}")
        for a, b in zip(ctx.features_get_state(sw.cloud.shape[0]), fa.point_state()):
            assert np.array_equal(a, b)
    ctx.close()
