"""Seeded synthetic cases shared by the CPU and GPU tests (small enough for the oracle)."""
from __future__ import annotations

import functools

import numpy as np

from lego_loam_b200 import synth


@functools.lru_cache(maxsize=None)
def world():
    return synth.make_world(synth.SEED0)


@functools.lru_cache(maxsize=None)
def mapping_case(seed: int = 1, n_corner_raw: int = 40000, n_surf_raw: int = 250000, sensor: str = "vlp16"):
    """-> dict(scan clouds, raw map, true pose, initial guess)"""
    rng = np.random.default_rng(1000 + seed)
    w = world()
    pose = np.array([rng.uniform(-0.02, 0.02), rng.uniform(-3.1, 3.1), rng.uniform(-0.02, 0.02),
                     rng.uniform(-20, 20), rng.uniform(-0.05, 0.05), rng.uniform(-20, 20)])
    sc = synth.make_mapping_scan(w, synth.SENSORS[sensor], pose, seed=seed)
    mc, ms = synth.make_local_map(w, pose[3:6], n_corner_raw, n_surf_raw, seed=seed + 7)
    init = synth.perturb_pose(pose, rng)
    return dict(corner=sc.corner_last, surf=sc.surf_last, outlier=sc.outlier_last, map_corner_raw=mc,
                map_surf_raw=ms, pose=pose, init=init)


def random_cloud(n: int, seed: int, extent=(60.0, 8.0, 60.0), clustered: bool = True) -> np.ndarray:
    """Points with many shared voxels (clustered) and realistic intensities (ring + fraction)."""
    rng = np.random.default_rng(seed)
    if clustered:
        k = max(n // 6, 1)
        centers = rng.uniform(-1, 1, (k, 3)) * np.array(extent)
        pts = centers[rng.integers(0, k, n)] + rng.normal(0, 0.08, (n, 3))
    else:
        pts = rng.uniform(-1, 1, (n, 3)) * np.array(extent)
    out = np.empty((n, 4), np.float32)
    out[:, :3] = pts
    out[:, 3] = rng.integers(0, 16, n) + rng.random(n).astype(np.float32) * 0.1
    return out


_LOOP_CASE = {}


def loop_closure_case(n_out: int = 24, step: float = 0.7):
    """A key-frame store the REFERENCE node logic (oracle/_ref, unmodified mapOptmization.cpp) builds over an out-and-back
    drive: n_out key-frames away from the start, n_out back on a lane 1 m beside it, 2.5 s apart, with an odometry drift that
    grows on the way back - so that at the end detectLoopClosure (MO:814-872) finds a history frame older than 30 s within
    7 m and performLoopClosure (MO:875-945) aligns the latest key-frame against the 51-frame history sub-map.
    -> dict(mo, n_kf, truth poses); cached (the harness object is reused read-only by the tests)."""
    key = (n_out, step)
    if key in _LOOP_CASE:
        return _LOOP_CASE[key]
    from oracle import ref_harness
    w = world()
    mo = ref_harness.MapOptimization()
    truth = []
    n = 2 * n_out
    for k in range(n):
        if k < n_out:
            x, z, yaw = 2.0, -10.0 + step * k, 0.0
        else:
            x, z, yaw = 3.0, -10.0 + step * (n - 1 - k), np.pi
        pose = np.array([0.004 * np.sin(0.3 * k), yaw + 0.01 * np.sin(0.2 * k), 0.004 * np.cos(0.2 * k), x, 0.0, z])
        f = max(0.0, (k - n_out) / max(n_out - 1, 1))
        odo = pose + f * np.array([0.002, 0.012, -0.002, 0.25, 0.03, -0.2])         # drift of the odometry on the way back
        sc = synth.make_mapping_scan(w, synth.VLP16, pose, seed=4000 + k)
        mo.set_odometry(odo.astype(np.float32), 2.5 * k)
        mo.set_scan(sc.corner_last, sc.surf_last, sc.outlier_last)
        mo.transformAssociateToMap()
        mo.downsampleCurrentScan()
        mo.transformUpdate()
        mo.saveKeyFramesAndFactor()
        mo.correctPoses()
        mo.clearCloud()
        truth.append(pose)
    out = dict(mo=mo, n_kf=mo.num_keyframes(), truth=np.array(truth), last_odo=odo, t_last=2.5 * (n - 1))
    _LOOP_CASE[key] = out
    return out
