"""Design check for the next row (imageProjection on the device): the data-parallel formulation the kernels
will use, written with numpy / scipy, gives exactly the images of the UNMODIFIED reference (oracle/_ref/libref_ip.so):
  * projection with "the last point of the cloud that falls into a pixel wins" (the reference overwrites sequentially),
  * ground marking as a per-pixel closed form of the reference's column sweep (a later row pair can overwrite a mark with
    "invalid"),
  * labelComponents' BFS = connected components of the symmetric edge predicate angle > segmentTheta on the row/column
    (column-wrapped) 4-neighbourhood, validity from the component's size and the rows of its points EXCEPT the seed
    (raster-first point), label numbers in raster order of the seeds of the valid components."""
import ctypes
import ctypes.util

import numpy as np
import pytest
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components

from oracle import ref_harness as rh
from lego_loam_b200 import synth

pytestmark = pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")

_m = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _f in ("atan2f", "sinf", "cosf", "sqrtf"):
    getattr(_m, _f).restype = ctypes.c_float
    getattr(_m, _f).argtypes = [ctypes.c_float] * (2 if _f == "atan2f" else 1)
atan2f = np.vectorize(lambda y, x: _m.atan2f(float(y), float(x)), otypes=[np.float32])
F = np.float32


def parallel_form(cloud, ring, N=16, H=1800, ang_res_x=0.2, ang_res_y=2.0, gsi=7):
    x, y, z = cloud[:, 0], cloud[:, 1], cloud[:, 2]
    ha = (atan2f(x, y).astype(np.float64) * np.float64(F(180.0))).astype(F)             # float * 180 in float
    ha = (ha.astype(np.float64) / np.pi).astype(F)
    colf = -np.round((ha.astype(np.float64) - 90.0) / np.float64(F(ang_res_x))) + H // 2
    col = colf.astype(np.int64); col = np.where(col >= H, col - H, col)
    rng = np.sqrt((x * x + y * y + z * z).astype(F)).astype(F)                             # IEEE sqrt: same as sqrtf
    ok = (ring < N) & (col < H) & (rng >= F(1.0))
    pix = ring.astype(np.int64) * H + col
    winner = np.full(N * H, -1, np.int64)
    idx = np.where(ok)[0]
    np.maximum.at(winner, pix[idx], idx)                                                   # last point wins
    has = winner >= 0
    rmat = np.full(N * H, np.finfo(F).max, F); rmat[has] = rng[winner[has]]
    pts = np.zeros((N * H, 3), F); pts[has] = cloud[winner[has], :3]
    has2 = has.reshape(N, H); pts2 = pts.reshape(N, H, 3)
    # ---- ground: pair(r) = rows (r, r+1) both valid; pass(r) = |angle| <= 10
    d = pts2[1:gsi + 1] - pts2[:gsi]
    hyp = np.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(F)).astype(F)
    ang = ((atan2f(d[..., 2], hyp).astype(F) * F(180.0)).astype(np.float64) / np.pi).astype(F)
    pair = has2[:gsi] & has2[1:gsi + 1]
    ok_pass = pair & (np.abs(ang - F(0.0)) <= 10)
    ground = np.zeros((N, H), bool)
    for r in range(gsi + 1):
        from_below = ok_pass[r - 1] if r > 0 else np.zeros(H, bool)
        if r < gsi:
            ground[r] = pair[r] & (ok_pass[r] | from_below)        # an invalid pair (r, r+1) overwrites the mark with -1
        else:
            ground[r] = from_below
    # ---- segmentation
    cand = has2 & ~ground
    r2 = rmat.reshape(N, H)
    sx, cx = F(_m.sinf(F(ang_res_x / 180.0 * np.pi))), F(_m.cosf(F(ang_res_x / 180.0 * np.pi)))
    sy, cy = F(_m.sinf(F(ang_res_y / 180.0 * np.pi))), F(_m.cosf(F(ang_res_y / 180.0 * np.pi)))
    theta = F(60.0 / 180.0 * np.pi)

    def passing(a, b, s, c):
        d1 = np.maximum(a, b); d2 = np.minimum(a, b)
        return atan2f((d2 * s).astype(F), (d1 - (d2 * c).astype(F)).astype(F)) > theta
    lin = np.arange(N * H).reshape(N, H)
    right = np.roll(lin, -1, axis=1)
    e_h = cand & np.roll(cand, -1, axis=1)
    e_h[e_h] = passing(r2[e_h], np.roll(r2, -1, axis=1)[e_h], sx, cx)
    e_v = np.zeros((N, H), bool); e_v[:-1] = cand[:-1] & cand[1:]
    down_r = np.zeros((N, H), F); down_r[:-1] = r2[1:]
    e_v[e_v] = passing(r2[e_v], down_r[e_v], sy, cy)
    src = np.concatenate([lin[e_h], lin[e_v]]); dst = np.concatenate([right[e_h], (lin + H)[e_v]])
    g = coo_matrix((np.ones(src.size, np.int8), (src, dst)), shape=(N * H, N * H))
    _, comp = connected_components(g, directed=False)
    label = np.full(N * H, -1, np.int64)
    cidx = np.where(cand.ravel())[0]
    cc = comp[cidx]
    order = np.argsort(cc, kind="stable")                      # inside a component: raster order
    cs, cstart, ccount = np.unique(cc[order], return_index=True, return_counts=True)
    members = cidx[order]
    seeds = members[cstart]
    valid = np.zeros(cs.size, bool)
    for k in range(cs.size):
        m = members[cstart[k]:cstart[k] + ccount[k]]
        rows = np.unique(m[1:] // H)                            # the seed's own row only counts through another point
        valid[k] = ccount[k] >= 30 or (ccount[k] >= 5 and rows.size >= 3)
    number = np.zeros(cs.size, np.int64)
    by_seed = np.argsort(seeds)
    number[by_seed] = np.cumsum(valid[by_seed])                 # labelCount at the time the seed is reached
    for k in range(cs.size):
        label[members[cstart[k]:cstart[k] + ccount[k]]] = number[k] if valid[k] else 999999
    return rmat.reshape(N, H), ground, label.reshape(N, H)


@pytest.mark.parametrize("seed,noise,dropout", [(60, 0.02, 0.02), (61, 0.05, 0.3), (62, 0.0, 0.0)])
def test_parallel_form_equals_reference_images(seed, noise, dropout):
    w = synth.make_world()
    ip = rh.ImageProjection()
    cloud, ring = synth.make_raw_sweep(w, synth.VLP16, [0.01, 0.4, -0.01, 3 + seed % 7, 0, 5 - seed % 5], seed, noise=noise, dropout=dropout)
    # a few duplicate pixels: a second return on some beams (the later one must win)
    dup = cloud[::37].copy(); dup[:, :3] *= np.float32(1.013)
    cloud = np.concatenate([cloud, dup]); ring = np.concatenate([ring, ring[::37]])
    ip.process(cloud, ring)
    rm, gm, lm = ip.images()
    prm, pg, pl = parallel_form(cloud, ring)
    assert np.array_equal(rm, prm)
    assert np.array_equal(gm == 1, pg)
    assert np.array_equal(lm, pl)
    assert (lm == 999999).sum() > 0 and lm[lm < 999999].max() > 5
