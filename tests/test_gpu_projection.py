"""imageProjection on the device: llb_projection_* against the UNMODIFIED reference imageProjection.cpp
compiled in oracle/_ref (tier B) and against the C restatement (tier A): range / ground / label images, segmented cloud,
cloud_info and outlier cloud bit-identical; then the device-resident hand-over to the feature extraction."""
import dataclasses

import numpy as np
import pytest

import oracle
from oracle import ref_harness as rh
from lego_loam_b200 import api, synth

pytestmark = pytest.mark.gpu

ANG_RES_X = {"vlp16": 0.2, "hdl32e": 360.0 / 1800.0, "vls128": 0.2}                     # UT:65, UT:73, UT:81


def same_sweep(a, b):
    for f in dataclasses.fields(a):
        x, y = getattr(a, f.name), getattr(b, f.name)
        if isinstance(x, np.ndarray):
            assert x.shape == y.shape, (f.name, x.shape, y.shape)
            xe = x.view(np.uint32) if x.dtype == np.float32 else x
            ye = y.view(np.uint32) if y.dtype == np.float32 else y
            assert np.array_equal(xe, ye), f.name
        else:
            assert np.float32(x).tobytes() == np.float32(y).tobytes(), f.name


def make_ctx(sensor_key):
    sn = synth.SENSORS[sensor_key]
    c = api.Context(0)
    c.projection_init(sn.n_scan, sn.horizon, ANG_RES_X[sensor_key], sn.ang_res_y, sn.ground_scan_ind)
    return c, sn


def reference_for(sensor_key):
    """the compiled reference for the sensor when its build is present, else the restatement (pinned against it)"""
    if rh.available() and (sensor_key == "vlp16" or rh.sensor_available(sensor_key)):
        return rh.ImageProjection(None if sensor_key == "vlp16" else sensor_key), "reference"
    sn = synth.SENSORS[sensor_key]
    return oracle.ImageProjection(sn.n_scan, sn.horizon, ANG_RES_X[sensor_key], sn.ang_res_y, sn.ground_scan_ind), "restatement"


@pytest.mark.parametrize("noise,dropout", [(0.02, 0.02), (0.0, 0.0), (0.05, 0.3)])
def test_projection_matches_reference_vlp16(noise, dropout):
    w = synth.make_world()
    ctx, sn = make_ctx("vlp16")
    ref, _ = reference_for("vlp16")
    try:
        for k in range(3):
            cloud, ring = synth.make_raw_sweep(w, sn, [0.01 * k, 0.3 * k, -0.01 * k, 3 + 2.0 * k, 0, 5 - 1.5 * k], 50 + k,
                                               noise=noise, dropout=dropout)
            # a few duplicate pixels: a second return on some beams (the later one must win)
            dup = cloud[::37].copy(); dup[:, :3] *= np.float32(1.013)
            cloud = np.concatenate([cloud, dup]); ring = np.concatenate([ring, ring[::37]])
            a = ref.process(cloud, ring)
            ns, no, ms = ctx.projection_process(cloud, ring)
            for x, y, name in zip(ref.images(), ctx.projection_get_images(), ("rangeMat", "groundMat", "labelMat")):
                assert np.array_equal(x, y), (name, int(np.sum(x != y)))
            b = ctx.projection_get_sweep()
            assert (ns, no) == (a.cloud.shape[0], a.outlier.shape[0])
            same_sweep(a, b)
            assert a.cloud.shape[0] > 3000 and a.ground.sum() > 500
    finally:
        ctx.close()


@pytest.mark.parametrize("sensor_key", ["hdl32e", "vls128"])
def test_projection_other_sensors(sensor_key):
    w = synth.make_world()
    ctx, sn = make_ctx(sensor_key)
    ref, kind = reference_for(sensor_key)
    try:
        for k in range(2):
            cloud, ring = synth.make_raw_sweep(w, sn, [0.0, 0.5 * k, 0.01, 4 + k, 0, -3 + 2 * k], 70 + k)
            a = ref.process(cloud, ring)
            ctx.projection_process(cloud, ring)
            for x, y, name in zip(ref.images(), ctx.projection_get_images(), ("rangeMat", "groundMat", "labelMat")):
                assert np.array_equal(x, y), (sensor_key, kind, name, int(np.sum(x != y)))
            same_sweep(a, ctx.projection_get_sweep())
    finally:
        ctx.close()


def test_projection_edge_cases():
    """Points on rings beyond N_SCAN, below the minimum range, tiny sweeps, a cloud of tiny clusters (everything rejected by
    the segment-size test), an empty sweep."""
    ctx, sn = make_ctx("vlp16")
    ref, _ = reference_for("vlp16")
    rng = np.random.default_rng(1)
    n = 4000
    az = rng.uniform(-np.pi, np.pi, n); r = rng.uniform(0.2, 60.0, n); ring = rng.integers(0, 20, n).astype(np.uint16)
    el = np.deg2rad(-15 + 2.0 * np.minimum(ring, 15))
    cloud = np.zeros((n, 4), np.float32)
    cloud[:, 0] = r * np.cos(el) * np.cos(az); cloud[:, 1] = r * np.cos(el) * np.sin(az); cloud[:, 2] = r * np.sin(el)
    try:
        for c, rg in ((cloud, ring), (cloud[:1], ring[:1]), (cloud[:50], ring[:50])):
            a = ref.process(c, rg)
            ctx.projection_process(c, rg)
            for x, y, name in zip(ref.images(), ctx.projection_get_images(), ("rangeMat", "groundMat", "labelMat")):
                assert np.array_equal(x, y), name
            same_sweep(a, ctx.projection_get_sweep())
        ns, no, _ = ctx.projection_process(cloud[:0], ring[:0])
        assert (ns, no) == (0, 0)
    finally:
        ctx.close()


def test_projection_feeds_features_on_device():
    """raw sweep -> imageProjection -> feature extraction without leaving the device: the four feature clouds equal
    those of the host hand-over (llb_features_extract on the same segmented sweep) and of the compiled reference."""
    w = synth.make_world()
    ctx, sn = make_ctx("vlp16"); ctx.features_init(sn.n_scan, sn.horizon)
    c2 = api.Context(0); c2.features_init(sn.n_scan, sn.horizon)
    rip = rh.ImageProjection() if rh.available() else None
    rfa = rh.FeatureAssociation() if rh.available() else None
    try:
        for k in range(3):                                   # state survives between sweeps on both sides
            cloud, ring = synth.make_raw_sweep(w, sn, [0.0, 0.2 * k, 0.0, 2 + 1.5 * k, 0, 4 - k], 90 + k)
            ctx.projection_process(cloud, ring)
            counts, ms = ctx.projection_to_features()
            sw = ctx.projection_get_sweep()
            c2.features_extract(sw)
            for which in range(4):
                a, b = ctx.features_get(which), c2.features_get(which)
                assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (k, which)
            if rfa is not None:
                rsw = rip.process(cloud, ring)
                rfa.set_segmented(rsw); rfa.extract_features()
                for which in range(4):
                    a, b = ctx.features_get(which), rfa.feature_cloud(which)
                    assert a.shape == b.shape and np.array_equal(a[:, :3].view(np.uint32), b[:, :3].view(np.uint32)), (k, which)
                    assert np.max(np.abs(a[:, 3] - b[:, 3]), initial=0.0) <= 2e-6
            assert counts[0] > 20 and counts[3] > 1000
    finally:
        ctx.close(); c2.close()
