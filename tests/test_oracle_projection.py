"""imageProjection oracle (the row after feature extraction): the C restatement (oracle/llo_projection.c)
against the UNMODIFIED reference imageProjection.cpp compiled in oracle/_ref: range / ground / label images, segmented
cloud, cloud_info and outlier cloud bit-identical."""
import dataclasses

import numpy as np
import pytest

import oracle
from oracle import ref_harness as rh
from lego_loam_b200 import synth

needs_ref = pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")


def same_sweep(a, b):
    for f in dataclasses.fields(a):
        x, y = getattr(a, f.name), getattr(b, f.name)
        if isinstance(x, np.ndarray):
            assert x.shape == y.shape, f.name
            assert np.array_equal(x.view(np.uint32) if x.dtype == np.float32 else x, y.view(np.uint32) if y.dtype == np.float32 else y), f.name
        else:
            assert np.float32(x).tobytes() == np.float32(y).tobytes(), f.name


@needs_ref
@pytest.mark.parametrize("noise,dropout", [(0.02, 0.02), (0.0, 0.0), (0.05, 0.3)])
def test_projection_restatement_matches_reference(noise, dropout):
    w = synth.make_world()
    ref = rh.ImageProjection(); mine = oracle.ImageProjection()
    for k in range(3):
        cloud, ring = synth.make_raw_sweep(w, synth.VLP16, [0.01 * k, 0.3 * k, -0.01 * k, 3 + 2.0 * k, 0, 5 - 1.5 * k], 50 + k,
                                           noise=noise, dropout=dropout)
        a = ref.process(cloud, ring); b = mine.process(cloud, ring)
        for x, y, name in zip(ref.images(), mine.images(), ("rangeMat", "groundMat", "labelMat")):
            assert np.array_equal(x, y), name
        same_sweep(a, b)
        assert a.cloud.shape[0] > 3000 and a.ground.sum() > 500


@needs_ref
def test_projection_edge_cases():
    """Points on rings beyond N_SCAN, below the minimum range, NaN-free empty-ish sweeps, and a cloud of tiny clusters
    (everything rejected by the segment-size test)."""
    ref = rh.ImageProjection(); mine = oracle.ImageProjection()
    rng = np.random.default_rng(1)
    n = 4000
    az = rng.uniform(-np.pi, np.pi, n); r = rng.uniform(0.2, 60.0, n); ring = rng.integers(0, 20, n).astype(np.uint16)
    el = np.deg2rad(-15 + 2.0 * np.minimum(ring, 15))
    cloud = np.zeros((n, 4), np.float32)
    cloud[:, 0] = r * np.cos(el) * np.cos(az); cloud[:, 1] = r * np.cos(el) * np.sin(az); cloud[:, 2] = r * np.sin(el)
    for c, rg in ((cloud, ring), (cloud[:1], ring[:1]), (cloud[:50], ring[:50])):
        a = ref.process(c, rg); b = mine.process(c, rg)
        for x, y, name in zip(ref.images(), mine.images(), ("rangeMat", "groundMat", "labelMat")):
            assert np.array_equal(x, y), name
        same_sweep(a, b)
