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
