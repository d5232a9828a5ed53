"""Synthetic range-image-shaped lidar data for the scan-matching hot path.

No bags are available offline (BASELINE.json north_star), so every test and
bench input comes from here: a seeded primitive world (ground plane, boxes,
poles) in the LOAM camera frame (x left, y up, z forward -- the axis swizzle of
FA:500-502), ray-cast with the sensor models of utility.h (UT:63-84), with the
point metadata the hot path reads from ``intensity``:

* mapping inputs (MO:608-627): ``intensity`` = integer ring (after TransformToEnd,
  FA:952);
* odometry inputs (FA:1044-1268): ``intensity`` = ring + scanPeriod * relTime
  (FA:523), array order = ring-major then azimuth (the +-2.5 ring neighbour scan
  of FA:1062-1099 depends on it).

Pose convention (MO:513-527): p_map = Ry(T[1]) Rx(T[0]) Rz(T[2]) p + T[3:6].

This module is host-side tooling (numpy only); it is not on the product path.
"""
from __future__ import annotations

import dataclasses
import numpy as np

SEED0 = 20181001
SCAN_PERIOD = 0.1         # UT:107


@dataclasses.dataclass(frozen=True)
class Sensor:
    name: str
    n_scan: int
    horizon: int
    ang_bottom: float      # deg, elevation of ring 0 is -ang_bottom
    ang_res_y: float       # deg
    ground_scan_ind: int
    max_range: float


VLP16 = Sensor("VLP-16", 16, 1800, 15.0, 2.0, 7, 100.0)            # UT:63-68
HDL32E = Sensor("HDL-32E", 32, 1800, 30.67, 41.33 / 31.0, 20, 100.0)  # UT:70-76
VLS128 = Sensor("VLS-128", 128, 1800, 25.0, 0.3, 10, 200.0)        # UT:78-84 (as the reference models it)
SENSORS = {"vlp16": VLP16, "hdl32e": HDL32E, "vls128": VLS128}


def rot_zxy(rx: float, ry: float, rz: float) -> np.ndarray:
    """R = Ry(ry) Rx(rx) Rz(rz) (float64), the mapping rotation of MO:513-527."""
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]])
    Rx = np.array([[1.0, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1.0, 0], [-sy, 0, cy]])
    return Ry @ Rx @ Rz


def apply_pose(T, pts: np.ndarray) -> np.ndarray:
    """Float64 application of the mapping transform to (n,>=3) points."""
    T = np.asarray(T, np.float64)
    R = rot_zxy(T[0], T[1], T[2])
    out = np.array(pts, np.float64, copy=True)
    out[:, :3] = pts[:, :3].astype(np.float64) @ R.T + T[3:6]
    return out


@dataclasses.dataclass
class World:
    h: float                   # sensor height above ground (ground plane y = -h in world when sensor at y=0)
    box_lo: np.ndarray         # (B,3)
    box_hi: np.ndarray         # (B,3)
    cyl_c: np.ndarray          # (P,2) x,z centre
    cyl_r: np.ndarray          # (P,)
    cyl_top: np.ndarray        # (P,) y of the top (bottom = ground)
    ground_y: float


def make_world(seed: int = SEED0, extent: float = 200.0, n_box: int = 120, n_cyl: int = 400,
               h: float = 0.8, keep_clear: float = 3.0) -> World:
    """Ground plane + axis-aligned boxes + vertical poles."""
    rng = np.random.default_rng(seed)
    gy = -h
    fx = rng.uniform(4, 30, n_box); fz = rng.uniform(4, 30, n_box); hy = rng.uniform(3, 15, n_box)
    cx = rng.uniform(-extent, extent, n_box); cz = rng.uniform(-extent, extent, n_box)
    # keep a corridor around the origin-centred path free
    near = (np.abs(cx) < fx / 2 + keep_clear + 6) & (np.abs(cz) < fz / 2 + keep_clear + 6)
    cx[near] += np.sign(cx[near] + 1e-9) * 25.0
    lo = np.stack([cx - fx / 2, np.full(n_box, gy), cz - fz / 2], 1)
    hi = np.stack([cx + fx / 2, gy + hy, cz + fz / 2], 1)
    pc = rng.uniform(-extent, extent, (n_cyl, 2))
    close = np.hypot(pc[:, 0], pc[:, 1]) < keep_clear
    pc[close] += 2 * keep_clear
    return World(h, lo, hi, pc, np.full(n_cyl, 0.15), gy + rng.uniform(2, 8, n_cyl), gy)


def sensor_dirs(sensor: Sensor) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Unit ray directions in the camera frame, ring index and column index, ring-major."""
    ring = np.repeat(np.arange(sensor.n_scan), sensor.horizon)
    col = np.tile(np.arange(sensor.horizon), sensor.n_scan)
    elev = np.deg2rad(-sensor.ang_bottom + sensor.ang_res_y * ring)
    az = np.deg2rad(360.0 * col / sensor.horizon) - np.pi
    # lidar (x fwd, y left, z up) -> camera (x=y_l, y=z_l, z=x_l)
    d = np.stack([np.cos(elev) * np.sin(az), np.sin(elev), np.cos(elev) * np.cos(az)], 1)
    return d, ring, col


def raycast(world: World, origin: np.ndarray, dirs: np.ndarray, max_range: float,
            min_range: float = 1.0, chunk: int = 8192):
    """Nearest hit per ray.  origin (n,3) or (3,), dirs (n,3) in world frame.

    Returns (range, kind, prim, edge_dist): kind 0 none, 1 ground, 2 box, 3 pole;
    edge_dist = horizontal distance from a box hit to the nearest vertical box edge.
    """
    n = dirs.shape[0]
    origin = np.broadcast_to(np.asarray(origin, np.float64), (n, 3))
    rng_out = np.full(n, np.inf); kind = np.zeros(n, np.int8); prim = np.full(n, -1, np.int32)
    edge = np.full(n, np.inf)
    # cull primitives out of reach of the (roughly common) origin
    oc = origin.mean(0)
    bsel = np.where((world.box_lo[:, 0] < oc[0] + max_range) & (world.box_hi[:, 0] > oc[0] - max_range) &
                    (world.box_lo[:, 2] < oc[2] + max_range) & (world.box_hi[:, 2] > oc[2] - max_range))[0]
    csel = np.where(np.hypot(world.cyl_c[:, 0] - oc[0], world.cyl_c[:, 1] - oc[2]) < max_range + 1)[0]
    for s in range(0, n, chunk):
        o = origin[s:s + chunk]; d = dirs[s:s + chunk]; m = o.shape[0]
        best = np.full(m, np.inf); bk = np.zeros(m, np.int8); bp = np.full(m, -1, np.int32)
        # ground
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (world.ground_y - o[:, 1]) / d[:, 1]
        ok = (d[:, 1] < 0) & (t > 0)
        best = np.where(ok, t, best); bk = np.where(ok, 1, bk).astype(np.int8)
        # boxes (slab test)
        if bsel.size:
            lo = world.box_lo[bsel][None]; hi = world.box_hi[bsel][None]
            with np.errstate(divide="ignore", invalid="ignore"):
                inv = 1.0 / d[:, None, :]
                t0 = (lo - o[:, None, :]) * inv; t1 = (hi - o[:, None, :]) * inv
            tn = np.nanmax(np.minimum(t0, t1), axis=2); tf = np.nanmin(np.maximum(t0, t1), axis=2)
            hit = (tf >= tn) & (tn > 0)
            tn = np.where(hit, tn, np.inf)
            j = np.argmin(tn, 1); tb = tn[np.arange(m), j]
            better = tb < best
            best = np.where(better, tb, best); bk = np.where(better, 2, bk).astype(np.int8)
            bp = np.where(better, bsel[j], bp)
        # poles (vertical cylinders from the ground to cyl_top)
        if csel.size:
            c = world.cyl_c[csel]; r = world.cyl_r[csel]; top = world.cyl_top[csel]
            ox = o[:, None, 0] - c[None, :, 0]; oz = o[:, None, 2] - c[None, :, 1]
            dx = d[:, None, 0]; dz = d[:, None, 2]
            a = dx * dx + dz * dz; b = ox * dx + oz * dz; cc = ox * ox + oz * oz - r[None] ** 2
            disc = b * b - a * cc
            with np.errstate(divide="ignore", invalid="ignore"):
                tc = (-b - np.sqrt(np.where(disc > 0, disc, np.nan))) / a
            yhit = o[:, None, 1] + tc * d[:, None, 1]
            ok = (disc > 0) & (tc > 0) & (yhit <= top[None]) & (yhit >= world.ground_y)
            tc = np.where(ok, tc, np.inf)
            j = np.argmin(tc, 1); tcb = tc[np.arange(m), j]
            better = tcb < best
            best = np.where(better, tcb, best); bk = np.where(better, 3, bk).astype(np.int8)
            bp = np.where(better, csel[j], bp)
        valid = (best < max_range) & (best > min_range)
        rng_out[s:s + chunk] = np.where(valid, best, np.inf)
        kind[s:s + chunk] = np.where(valid, bk, 0)
        prim[s:s + chunk] = np.where(valid, bp, -1)
        # distance to nearest vertical edge for box hits
        isb = valid & (bk == 2)
        if isb.any():
            hp = o[isb] + best[isb, None] * d[isb]
            lo = world.box_lo[bp[isb]]; hi = world.box_hi[bp[isb]]
            ex = np.minimum(np.abs(hp[:, 0] - lo[:, 0]), np.abs(hp[:, 0] - hi[:, 0]))
            ez = np.minimum(np.abs(hp[:, 2] - lo[:, 2]), np.abs(hp[:, 2] - hi[:, 2]))
            e = np.full(m, np.inf); e[isb] = np.hypot(ex, ez)
            edge[s:s + chunk] = e
    return rng_out, kind, prim, edge


def _pack(xyz: np.ndarray, intensity: np.ndarray) -> np.ndarray:
    out = np.empty((xyz.shape[0], 4), np.float32)
    out[:, :3] = xyz
    out[:, 3] = intensity
    return out


@dataclasses.dataclass
class MappingScan:
    """Inputs of mapOptimization for one registration (sensor frame, de-skewed)."""
    corner_last: np.ndarray    # (n,4) float32: laserCloudCornerLast
    surf_last: np.ndarray      # laserCloudSurfLast
    outlier_last: np.ndarray   # laserCloudOutlierLast
    pose_true: np.ndarray      # (6,) float64


def make_mapping_scan(world: World, sensor: Sensor, pose, seed: int, noise: float = 0.02,
                      dropout: float = 0.02) -> MappingScan:
    """One undistorted sweep at `pose`, split into the three clouds MO receives.

    Labelling stands in for imageProjection + extractFeatures (out of scope):
    poles and box vertical edges -> corner (<= 20 per 6 sectors per ring, FA:713);
    remaining ground/wall hits thinned to ~1/4 -> surf; every 5th column of the
    non-ground rings among the rest -> outliers (IP:328-331).
    """
    rng = np.random.default_rng(seed)
    pose = np.asarray(pose, np.float64)
    d, ring, col = sensor_dirs(sensor)
    R = rot_zxy(pose[0], pose[1], pose[2])
    r, kind, prim, edge = raycast(world, pose[3:6], d @ R.T, sensor.max_range)
    ok = np.isfinite(r) & (rng.random(r.shape[0]) > dropout)
    r = r + rng.normal(0, noise, r.shape[0])
    pts = d * np.where(np.isfinite(r), r, 0.0)[:, None]
    is_corner = ok & ((kind == 3) | ((kind == 2) & (edge < 0.12)))
    # cap corners at 20 per sector per ring
    sector = (col * 6) // sensor.horizon
    key = ring * 6 + sector
    keep = np.zeros_like(is_corner)
    idx = np.where(is_corner)[0]
    if idx.size:
        order = np.argsort(key[idx], kind="stable")
        ks = key[idx][order]
        first = np.r_[0, np.where(np.diff(ks) != 0)[0] + 1]
        rank = np.arange(ks.size) - np.repeat(first, np.diff(np.r_[first, ks.size]))
        keep[idx[order][rank < 20]] = True
    is_corner = keep
    rest = ok & ~is_corner
    is_surf = rest & ((col % 4) == (ring % 4))
    is_out = rest & ~is_surf & (ring > sensor.ground_scan_ind) & (col % 5 == 0) & (kind != 1)
    mk = lambda m: _pack(pts[m], ring[m].astype(np.float32))
    return MappingScan(mk(is_corner), mk(is_surf), mk(is_out), pose)


def make_local_map(world: World, center, n_corner_raw: int, n_surf_raw: int, seed: int,
                   radius: float = 60.0, noise: float = 0.02, surf_radius: float | None = None
                   ) -> tuple[np.ndarray, np.ndarray]:
    """Raw (pre-voxel) local map around `center` in the map frame, as the concatenated
    surrounding key-frames would give it (MO:1051-1055): corner points on pole axes and
    box vertical edges; surf points on ground, walls and roofs.  intensity = ring-like int.
    """
    rng = np.random.default_rng(seed)
    c = np.asarray(center, np.float64)
    # ---- corners
    cyl = np.where(np.hypot(world.cyl_c[:, 0] - c[0], world.cyl_c[:, 1] - c[2]) < radius)[0]
    bx = np.where((world.box_lo[:, 0] < c[0] + radius) & (world.box_hi[:, 0] > c[0] - radius) &
                  (world.box_lo[:, 2] < c[2] + radius) & (world.box_hi[:, 2] > c[2] - radius))[0]
    segs = []   # (x, z, y0, y1)
    for i in cyl:
        segs.append((world.cyl_c[i, 0], world.cyl_c[i, 1], world.ground_y, world.cyl_top[i]))
    for i in bx:
        for x in (world.box_lo[i, 0], world.box_hi[i, 0]):
            for z in (world.box_lo[i, 2], world.box_hi[i, 2]):
                segs.append((x, z, world.ground_y, world.box_hi[i, 1]))
    segs = np.array(segs) if segs else np.zeros((1, 4))
    ln = segs[:, 3] - segs[:, 2]
    pick = rng.choice(len(segs), n_corner_raw, p=ln / ln.sum())
    cy = segs[pick, 2] + rng.random(n_corner_raw) * ln[pick]
    corner = np.stack([segs[pick, 0], cy, segs[pick, 1]], 1) + rng.normal(0, noise, (n_corner_raw, 3))
    # ---- surfaces: ground disc + box walls + roofs, area weighted
    if surf_radius is not None:
        radius = surf_radius
        bx = np.where((world.box_lo[:, 0] < c[0] + radius) & (world.box_hi[:, 0] > c[0] - radius) &
                      (world.box_lo[:, 2] < c[2] + radius) & (world.box_hi[:, 2] > c[2] - radius))[0]
    areas = [np.pi * radius ** 2]
    for i in bx:
        ex = world.box_hi[i] - world.box_lo[i]
        areas += [ex[0] * ex[1], ex[0] * ex[1], ex[2] * ex[1], ex[2] * ex[1], ex[0] * ex[2]]
    areas = np.array(areas)
    which = rng.choice(len(areas), n_surf_raw, p=areas / areas.sum())
    surf = np.empty((n_surf_raw, 3))
    g = which == 0
    rr = radius * np.sqrt(rng.random(g.sum())); th = rng.random(g.sum()) * 2 * np.pi
    surf[g] = np.stack([c[0] + rr * np.cos(th), np.full(g.sum(), world.ground_y), c[2] + rr * np.sin(th)], 1)
    w = np.where(~g)[0]
    bi = bx[(which[w] - 1) // 5]; face = (which[w] - 1) % 5
    u = rng.random(w.size); v = rng.random(w.size)
    lo = world.box_lo[bi]; hi = world.box_hi[bi]; ex = hi - lo
    p = np.empty((w.size, 3))
    f0 = face == 0; f1 = face == 1; f2 = face == 2; f3 = face == 3; f4 = face == 4
    p[:, 0] = np.where(f2, lo[:, 0], np.where(f3, hi[:, 0], lo[:, 0] + u * ex[:, 0]))
    p[:, 2] = np.where(f0, lo[:, 2], np.where(f1, hi[:, 2], lo[:, 2] + np.where(f2 | f3, u, v) * ex[:, 2]))
    p[:, 1] = np.where(f4, hi[:, 1], lo[:, 1] + v * ex[:, 1])
    surf[w] = p
    surf += rng.normal(0, noise, surf.shape)
    return (_pack(corner, rng.integers(0, 16, n_corner_raw).astype(np.float32)),
            _pack(surf, rng.integers(0, 16, n_surf_raw).astype(np.float32)))


def perturb_pose(pose, rng: np.random.Generator, rot: float = 0.01, trans: float = 0.08) -> np.ndarray:
    """Initial guess = truth + odometry-drift-sized error (what transformAssociateToMap hands over)."""
    p = np.array(pose, np.float64)
    p[:3] += rng.uniform(-rot, rot, 3)
    p[3:] += rng.uniform(-trans, trans, 3)
    return p.astype(np.float32)


# ------------------------------------------------------------------ odometry inputs

def _start_from_s(cur, s, p):
    """Float64 TransformToStart (FA:860-883) for arrays: p_start = Ry(-s ry) Rx(-s rx) Rz(-s rz) (p - s t)."""
    cur = np.asarray(cur, np.float64)
    out = np.empty_like(p)
    for k in range(p.shape[0]):
        A = rot_zxy(-s[k] * cur[0], -s[k] * cur[1], -s[k] * cur[2])
        # Ry(-ry)Rx(-rx)Rz(-rz) has the same Y-X-Z order as rot_zxy
        out[k] = A @ (p[k] - s[k] * cur[3:6])
    return out


@dataclasses.dataclass
class OdometryPair:
    corner_sharp: np.ndarray   # current sweep, intensity = ring + 0.1*relTime
    surf_flat: np.ndarray
    corner_last: np.ndarray    # previous sweep projected to its end, intensity = ring
    surf_last: np.ndarray
    cur_true: np.ndarray       # transformCur that maps current-sweep points to the sweep start


def make_odometry_pair(world: World, sensor: Sensor, pose_start, cur_true, seed: int,
                       noise: float = 0.02) -> OdometryPair:
    """Previous sweep (undistorted at pose_start, i.e. already 'TransformToEnd'-ed) and the
    current sweep captured while the sensor moves by `cur_true` (constant velocity model of
    FA:860-883).  Feature picking stands in for extractFeatures (FA:680-784)."""
    rng = np.random.default_rng(seed)
    pose_start = np.asarray(pose_start, np.float64); cur = np.asarray(cur_true, np.float64)
    d, ring, col = sensor_dirs(sensor)
    Rw = rot_zxy(*pose_start[:3]); tw = pose_start[3:6]
    sector = (col * 6) // sensor.horizon

    def label(kind, edge, ok):
        corner = ok & ((kind == 3) | ((kind == 2) & (edge < 0.12)))
        surf = ok & ~corner
        return corner, surf

    def cap(mask, per):
        key = ring * 6 + sector
        keep = np.zeros_like(mask)
        idx = np.where(mask)[0]
        if idx.size:
            order = np.argsort(key[idx], kind="stable")
            ks = key[idx][order]
            first = np.r_[0, np.where(np.diff(ks) != 0)[0] + 1]
            cnt = np.diff(np.r_[first, ks.size])
            rank = np.arange(ks.size) - np.repeat(first, cnt)
            # spread the picks over the sector instead of taking the first ones
            stride = np.maximum(np.repeat(cnt, cnt) // per, 1)
            keep[idx[order][(rank % stride == 0) & (rank // stride < per)]] = True
        return keep

    # previous sweep: undistorted at pose_start
    r0, k0, _, e0 = raycast(world, tw, d @ Rw.T, sensor.max_range)
    ok0 = np.isfinite(r0)
    r0 = r0 + rng.normal(0, noise, r0.shape[0])
    p0 = d * np.where(np.isfinite(r0), r0, 0.0)[:, None]
    c0, s0 = label(k0, e0, ok0)
    c0 = cap(c0, 20)
    s0 = s0 & ((col % 3) == 0)
    corner_last = _pack(p0[c0], ring[c0].astype(np.float32))
    surf_last = _pack(p0[s0], ring[s0].astype(np.float32))

    # current sweep: sensor pose at relative time s is pose_start o (A_s, b_s)
    s = col / sensor.horizon
    uniq = np.unique(col)
    org = np.empty((d.shape[0], 3)); dw = np.empty((d.shape[0], 3))
    for c_ in uniq:
        sc = c_ / sensor.horizon
        A = rot_zxy(-sc * cur[0], -sc * cur[1], -sc * cur[2])
        b = -A @ (sc * cur[3:6])
        m = col == c_
        org[m] = Rw @ b + tw
        dw[m] = d[m] @ (Rw @ A).T
    r1, k1, _, e1 = raycast(world, org, dw, sensor.max_range)
    ok1 = np.isfinite(r1)
    r1 = r1 + rng.normal(0, noise, r1.shape[0])
    p1 = d * np.where(np.isfinite(r1), r1, 0.0)[:, None]
    c1, s1 = label(k1, e1, ok1)
    sharp = cap(c1, 2)
    flat = cap(s1 & (k1 == 1) & (ring <= sensor.ground_scan_ind), 4)
    inten = ring + SCAN_PERIOD * s
    return OdometryPair(_pack(p1[sharp], inten[sharp]), _pack(p1[flat], inten[flat]),
                        corner_last, surf_last, cur.astype(np.float32))


@dataclasses.dataclass
class SegmentedSweep:
    """What imageProjection publishes for one sweep (IP:312-368, cloud_msgs/cloud_info): the input of FA's
    adjustDistortion / calculateSmoothness / markOccludedPoints / extractFeatures."""
    cloud: np.ndarray          # (n,4) float32, LIDAR frame (x fwd, y left, z up), intensity = row + col/10000
    start_ring: np.ndarray     # (n_scan,) int32
    end_ring: np.ndarray       # (n_scan,) int32
    start_ori: float
    end_ori: float
    ori_diff: float
    ground: np.ndarray         # (n,) uint8
    col: np.ndarray            # (n,) uint32
    range: np.ndarray          # (n,) float32
    outlier: np.ndarray        # (m,4) float32


def make_segmented_sweep(world: World, sensor: Sensor, pose, seed: int, noise: float = 0.02,
                         dropout: float = 0.02, clutter: float = 0.03) -> SegmentedSweep:
    """One sweep ray-cast at `pose` and packed the way imageProjection packs its segmented cloud (IP:318-356).

    Ground marking and segmentation stand in for IP:260-310 / IP:370-460 (out of scope): ground = ground
    hits on the rings up to ground_scan_ind, a random `clutter` share of the other hits plays the clusters that
    segmentation rejects (label 999999 -> outliers on every 5th column above the ground rings)."""
    rng = np.random.default_rng(seed)
    pose = np.asarray(pose, np.float64)
    d, ring, col = sensor_dirs(sensor)
    R = rot_zxy(pose[0], pose[1], pose[2])
    r, kind, _, _ = raycast(world, pose[3:6], d @ R.T, sensor.max_range)
    ok = np.isfinite(r) & (rng.random(r.shape[0]) > dropout)
    r = np.where(ok, r + rng.normal(0, noise, r.shape[0]), 0.0)
    cam = (d * r[:, None]).astype(np.float32)
    lid = np.stack([cam[:, 2], cam[:, 0], cam[:, 1]], 1)          # x_l = z_c, y_l = x_c, z_l = y_c
    rngf = np.sqrt(lid[:, 0] * lid[:, 0] + lid[:, 1] * lid[:, 1] + lid[:, 2] * lid[:, 2]).astype(np.float32)
    # column as IP:236-242 computes it from the point
    ha = np.arctan2(lid[:, 0], lid[:, 1]).astype(np.float32) * np.float32(180.0) / np.pi
    cidx = (-np.round((ha - 90.0) / (360.0 / sensor.horizon)) + sensor.horizon / 2).astype(np.int64)
    cidx = np.where(cidx >= sensor.horizon, cidx - sensor.horizon, cidx)
    ok &= (cidx >= 0) & (cidx < sensor.horizon) & (rngf >= 1.0)
    H, N = sensor.horizon, sensor.n_scan
    img_ok = np.zeros((N, H), bool); img_pt = np.zeros((N, H, 3), np.float32); img_r = np.zeros((N, H), np.float32)
    img_g = np.zeros((N, H), bool); img_rej = np.zeros((N, H), bool)
    sel = np.where(ok)[0]
    img_ok[ring[sel], cidx[sel]] = True
    img_pt[ring[sel], cidx[sel]] = lid[sel]
    img_r[ring[sel], cidx[sel]] = rngf[sel]
    img_g[ring[sel], cidx[sel]] = (kind[sel] == 1) & (ring[sel] <= sensor.ground_scan_ind)
    img_rej[ring[sel], cidx[sel]] = (~img_g[ring[sel], cidx[sel]]) & (rng.random(sel.size) < clutter)
    pts, grd, cols, rngs, outl = [], [], [], [], []
    start_ring = np.zeros(N, np.int32); end_ring = np.zeros(N, np.int32)
    size = 0
    jj = np.arange(H)
    for i in range(N):
        start_ring[i] = size - 1 + 5
        rej = img_ok[i] & img_rej[i]
        if i > sensor.ground_scan_ind:
            o = rej & (jj % 5 == 0)
            outl.append(np.concatenate([img_pt[i][o], (i + jj[o] / 10000.0).astype(np.float32)[:, None]], 1))
        keep = img_ok[i] & ~rej
        keep &= ~(img_g[i] & (jj % 5 != 0) & (jj > 5) & (jj < H - 5))
        k = np.where(keep)[0]
        pts.append(np.concatenate([img_pt[i][k], (np.float32(i) + (k.astype(np.float32) / np.float32(10000.0)))[:, None]], 1))
        grd.append(img_g[i][k]); cols.append(k); rngs.append(img_r[i][k])
        size += k.size
        end_ring[i] = size - 1 - 5
    cloud = np.concatenate(pts).astype(np.float32)
    first = lid[sel[0]]; last = lid[sel[-1]]
    # IP:199-211 on the first / last point of the raw cloud (float members of the message)
    start_ori = np.float32(-np.arctan2(first[1], first[0]))
    end_ori = np.float32(-np.arctan2(last[1], last[0]) + 2 * np.pi)
    if float(end_ori) - float(start_ori) > 3 * np.pi:
        end_ori = np.float32(float(end_ori) - 2 * np.pi)
    elif float(end_ori) - float(start_ori) < np.pi:
        end_ori = np.float32(float(end_ori) + 2 * np.pi)
    ori_diff = np.float32(end_ori - start_ori)
    return SegmentedSweep(cloud, start_ring, end_ring, float(start_ori), float(end_ori), float(ori_diff),
                          np.concatenate(grd).astype(np.uint8), np.concatenate(cols).astype(np.uint32),
                          np.concatenate(rngs).astype(np.float32),
                          np.concatenate(outl).astype(np.float32) if outl else np.zeros((0, 4), np.float32))


def make_raw_sweep(world: World, sensor: Sensor, pose, seed: int, noise: float = 0.02, dropout: float = 0.02,
                   quantize: float = 0.002):
    """One raw sweep as the driver publishes it: points in the LIDAR frame (x fwd, y left, z up) in firing order (all
    rings of an azimuth step, then the next step, clockwise as a Velodyne spins), with the ring channel.
    -> (cloud (n,4) float32 with intensity 0, ring (n,) uint16).  Ranges are quantised (2 mm, the sensor's resolution)."""
    rng = np.random.default_rng(seed)
    pose = np.asarray(pose, np.float64)
    d, ring, col = sensor_dirs(sensor)
    R = rot_zxy(pose[0], pose[1], pose[2])
    r, kind, _, _ = raycast(world, pose[3:6], d @ R.T, sensor.max_range)
    ok = np.isfinite(r) & (rng.random(r.shape[0]) > dropout)
    r = np.where(ok, r + rng.normal(0, noise, r.shape[0]), 0.0)
    if quantize:
        r = np.round(r / quantize) * quantize
    cam = d * r[:, None]
    lid = np.stack([cam[:, 2], cam[:, 0], cam[:, 1]], 1).astype(np.float32)
    order = np.lexsort((ring, -col))          # azimuth steps in decreasing column (clockwise), rings inside a step
    order = order[ok[order]]
    cloud = np.zeros((order.size, 4), np.float32); cloud[:, :3] = lid[order]
    return cloud, ring[order].astype(np.uint16)
