"""ctypes bindings of the CPU ORACLE (test infrastructure -- see oracle/llo.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (lego_loam_b200) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblegoloam_oracle.so")
_lib = None

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liblegoloam_oracle.so"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp = ctypes.c_void_p
        L.llo_kdtree_build.restype = vp
        L.llo_mapopt_create.restype = vp
        L.llo_featassoc_create.restype = vp
        for name in ("llo_kdtree_free", "llo_mapopt_destroy", "llo_featassoc_destroy"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = None
        _lib = L
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _pts(a) -> np.ndarray:
    a = _f32(a)
    if a.size == 0:
        return a.reshape(0, 4)
    assert a.ndim == 2 and a.shape[1] == 4, a.shape
    return a


def _fp(a: np.ndarray):
    return a.ctypes.data_as(c_float_p)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(c_int_p)


# ------------------------------------------------------------------ primitives

def cv_eigen(A):
    A = _f32(A).copy(); n = A.shape[0]
    W = np.zeros(n, np.float32); V = np.zeros((n, n), np.float32)
    lib().llo_cv_eigen_f32(n, _fp(A), _fp(W), _fp(V))
    return W, V


def cv_solve_qr(A, b):
    A = _f32(A); b = _f32(b).ravel(); m, n = A.shape
    x = np.zeros(n, np.float32)
    ok = lib().llo_cv_solve_qr_f32(m, n, _fp(A), _fp(b), _fp(x))
    return ok, x


def cv_inv(A):
    A = _f32(A); n = A.shape[0]
    D = np.zeros((n, n), np.float32)
    ok = lib().llo_cv_inv_f32(n, _fp(A), _fp(D))
    return ok, D


def cv_gemm(A, B):
    A = _f32(A); B = _f32(B)
    if B.ndim == 1:
        B = B.reshape(-1, 1)
    m, k = A.shape; n = B.shape[1]
    D = np.zeros((m, n), np.float32)
    lib().llo_cv_gemm_f32(m, k, n, _fp(A), _fp(B), _fp(D))
    return D


def voxel_grid(pts, leaf: float):
    """-> (out (m,4) float32, overflow flag)"""
    pts = _pts(pts); n = pts.shape[0]
    out = np.zeros((max(n, 1), 4), np.float32)
    ovf = ctypes.c_int(0)
    m = lib().llo_voxel_grid(_fp(pts), n, ctypes.c_float(leaf), _fp(out), ctypes.byref(ovf))
    return out[:m].copy(), int(ovf.value)


class KdTree:
    def __init__(self, pts):
        self.pts = _pts(pts)
        self._h = ctypes.c_void_p(lib().llo_kdtree_build(_fp(self.pts), self.pts.shape[0]))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().llo_kdtree_free(self._h); self._h = None

    def knn(self, queries, k: int):
        q = _f32(queries)[:, :3].copy()
        nq = q.shape[0]
        kk = min(k, self.pts.shape[0])
        idx = np.full((nq, k), -1, np.int32); d2 = np.zeros((nq, k), np.float32)
        L = lib()
        ti = np.zeros(k, np.int32); td = np.zeros(k, np.float32)
        for i in range(nq):
            L.llo_kdtree_knn(self._h, _fp(q[i]), k, _ip(ti), _fp(td))
            idx[i, :kk] = ti[:kk]; d2[i, :kk] = td[:kk]
        return idx, d2


def knn_bruteforce(pts, queries, k: int):
    pts = _pts(pts); q = _f32(queries)[:, :3].copy()
    nq = q.shape[0]; kk = min(k, pts.shape[0])
    idx = np.full((nq, k), -1, np.int32); d2 = np.zeros((nq, k), np.float32)
    ti = np.zeros(k, np.int32); td = np.zeros(k, np.float32)
    L = lib()
    for i in range(nq):
        L.llo_knn_bruteforce(_fp(pts), pts.shape[0], _fp(q[i]), k, _ip(ti), _fp(td))
        idx[i, :kk] = ti[:kk]; d2[i, :kk] = td[:kk]
    return idx, d2


# ------------------------------------------------------------------ mapOptimization

class MapOptimization:
    """CPU oracle with the reference's member-function names (MO:1067-1350)."""

    def __init__(self):
        self._h = ctypes.c_void_p(lib().llo_mapopt_create())
        self._keep = []

    def __del__(self):
        if getattr(self, "_h", None):
            lib().llo_mapopt_destroy(self._h); self._h = None

    def set_map_ds(self, corner_ds, surf_ds):
        c = _pts(corner_ds); s = _pts(surf_ds)
        lib().llo_mapopt_set_map_ds(self._h, _fp(c), c.shape[0], _fp(s), s.shape[0])

    def set_map_raw(self, corner, surf):
        c = _pts(corner); s = _pts(surf)
        lib().llo_mapopt_set_map_raw(self._h, _fp(c), c.shape[0], _fp(s), s.shape[0])

    def set_scan(self, corner_last, surf_last, outlier_last):
        c = _pts(corner_last); s = _pts(surf_last); o = _pts(outlier_last)
        lib().llo_mapopt_set_scan(self._h, _fp(c), c.shape[0], _fp(s), s.shape[0], _fp(o), o.shape[0])

    @property
    def transformTobeMapped(self):
        t = np.zeros(6, np.float32); lib().llo_mapopt_get_pose(self._h, _fp(t)); return t

    @transformTobeMapped.setter
    def transformTobeMapped(self, v):
        t = _f32(v); lib().llo_mapopt_set_pose(self._h, _fp(t))

    def set_transform_sum(self, v):
        t = _f32(v); lib().llo_mapopt_set_transform_sum(self._h, _fp(t))

    def bef_aft(self):
        b = np.zeros(6, np.float32); a = np.zeros(6, np.float32)
        lib().llo_mapopt_get_bef_aft(self._h, _fp(b), _fp(a)); return b, a

    def degenerate(self):
        d = ctypes.c_int(0); P = np.zeros((6, 6), np.float32)
        lib().llo_mapopt_get_degenerate(self._h, ctypes.byref(d), _fp(P)); return bool(d.value), P

    def downsampleCurrentScan(self): lib().llo_mapopt_downsampleCurrentScan(self._h)
    def build_kdtrees(self): lib().llo_mapopt_build_kdtrees(self._h)
    def clear_correspondences(self): lib().llo_mapopt_clear_correspondences(self._h)
    def cornerOptimization(self, it: int): lib().llo_mapopt_cornerOptimization(self._h, it)
    def surfOptimization(self, it: int): lib().llo_mapopt_surfOptimization(self._h, it)
    def LMOptimization(self, it: int) -> bool: return bool(lib().llo_mapopt_LMOptimization(self._h, it))
    def scan2MapOptimization(self) -> int: return lib().llo_mapopt_scan2MapOptimization(self._h)

    def _cloud(self, fn, which):
        n = fn(self._h, which, None, 0)
        out = np.zeros((max(n, 1), 4), np.float32)
        fn(self._h, which, _fp(out), n)
        return out[:n].copy()

    def scan_ds(self, which: int): return self._cloud(lib().llo_mapopt_get_scan_ds, which)
    def map_ds(self, which: int): return self._cloud(lib().llo_mapopt_get_map_ds, which)

    def correspondences(self):
        n = lib().llo_mapopt_get_correspondences(self._h, None, None, 0)
        ori = np.zeros((max(n, 1), 4), np.float32); co = np.zeros((max(n, 1), 4), np.float32)
        lib().llo_mapopt_get_correspondences(self._h, _fp(ori), _fp(co), n)
        return ori[:n].copy(), co[:n].copy()

    def knn(self, which: int):
        n = lib().llo_mapopt_get_knn(self._h, which, None, None, 0)
        idx = np.zeros((max(n, 1), 5), np.int32); d2 = np.zeros((max(n, 1), 5), np.float32)
        lib().llo_mapopt_get_knn(self._h, which, _ip(idx), _fp(d2), n)
        return idx[:n].copy(), d2[:n].copy()

    def normal_eq(self):
        A = np.zeros((6, 6), np.float32); B = np.zeros(6, np.float32); X = np.zeros(6, np.float32)
        lib().llo_mapopt_get_normal_eq(self._h, _fp(A), _fp(B), _fp(X)); return A, B, X


# ------------------------------------------------------------------ featureAssociation

class FeatureAssociation:
    """CPU oracle with the reference's member-function names (FA:1044-1478, FA:1666-1695)."""

    def __init__(self):
        self._h = ctypes.c_void_p(lib().llo_featassoc_create())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().llo_featassoc_destroy(self._h); self._h = None

    def set_last(self, corner_last, surf_last, force: bool = False):
        c = _pts(corner_last); s = _pts(surf_last)
        lib().llo_featassoc_set_last(self._h, _fp(c), c.shape[0], _fp(s), s.shape[0], int(force))

    def set_features(self, corner_sharp, surf_flat):
        c = _pts(corner_sharp); s = _pts(surf_flat)
        lib().llo_featassoc_set_features(self._h, _fp(c), c.shape[0], _fp(s), s.shape[0])

    @property
    def transformCur(self):
        t = np.zeros(6, np.float32); lib().llo_featassoc_get_transform(self._h, _fp(t)); return t

    @transformCur.setter
    def transformCur(self, v):
        t = _f32(v); lib().llo_featassoc_set_transform(self._h, _fp(t))

    def degenerate(self):
        d = ctypes.c_int(0); P = np.zeros((3, 3), np.float32)
        lib().llo_featassoc_get_degenerate(self._h, ctypes.byref(d), _fp(P)); return bool(d.value), P

    def clear_correspondences(self): lib().llo_featassoc_clear_correspondences(self._h)
    def findCorrespondingCornerFeatures(self, it): lib().llo_featassoc_findCorrespondingCornerFeatures(self._h, it)
    def findCorrespondingSurfFeatures(self, it): lib().llo_featassoc_findCorrespondingSurfFeatures(self._h, it)
    def calculateTransformationSurf(self, it) -> bool: return bool(lib().llo_featassoc_calculateTransformationSurf(self._h, it))
    def calculateTransformationCorner(self, it) -> bool: return bool(lib().llo_featassoc_calculateTransformationCorner(self._h, it))

    def updateTransformation(self):
        r = lib().llo_featassoc_updateTransformation(self._h)
        return r & 0xFFFF, r >> 16

    def correspondences(self):
        n = lib().llo_featassoc_get_correspondences(self._h, None, None, 0)
        ori = np.zeros((max(n, 1), 4), np.float32); co = np.zeros((max(n, 1), 4), np.float32)
        lib().llo_featassoc_get_correspondences(self._h, _fp(ori), _fp(co), n)
        return ori[:n].copy(), co[:n].copy()

    def search_ind(self, which: int):
        n = lib().llo_featassoc_get_search_ind(self._h, which, None, None, None, 0)
        a = np.zeros(max(n, 1), np.float32); b = np.zeros(max(n, 1), np.float32); c = np.zeros(max(n, 1), np.float32)
        lib().llo_featassoc_get_search_ind(self._h, which, _fp(a), _fp(b), _fp(c), n)
        return a[:n].copy(), b[:n].copy(), c[:n].copy()


def icp_align(src, tgt, max_iterations=100, max_corr_dist=100.0, transformation_epsilon=1e-6, euclidean_fitness_epsilon=1e-6):
    """pcl::IterativeClosestPoint::align + getFitnessScore as MO:892-904 configures them (restated PCL 1.8, unpinned)."""
    s = _pts(src); t = _pts(tgt)
    T = np.zeros(16, np.float32); c = ctypes.c_int(0); it = ctypes.c_int(0); st = ctypes.c_int(0); f = ctypes.c_double(0)
    lib().llo_icp_align(_fp(s), s.shape[0], _fp(t), t.shape[0], int(max_iterations), ctypes.c_double(max_corr_dist),
                        ctypes.c_double(transformation_epsilon), ctypes.c_double(euclidean_fitness_epsilon), _fp(T),
                        ctypes.byref(c), ctypes.byref(it), ctypes.byref(st), ctypes.byref(f))
    return dict(T=T.reshape(4, 4), converged=bool(c.value), iterations=it.value, state=st.value, fitness=f.value)


def icp_step(cur, tgt, max_corr_dist=100.0):
    """one iteration on the CURRENT source cloud: (n, sums p[3], q[3], qp[9], mse, nn index per point, umeyama 4x4)"""
    s = _pts(cur); t = _pts(tgt)
    L = lib()
    tree = ctypes.c_void_p(L.llo_kdtree_build(_fp(t), t.shape[0]))
    sp = np.zeros(3); sq = np.zeros(3); sqp = np.zeros(9); mse = ctypes.c_double(0); nn = np.zeros(max(s.shape[0], 1), np.int32)
    dp = ctypes.POINTER(ctypes.c_double)
    n = L.llo_icp_correspondence_sums(_fp(s), s.shape[0], tree, _fp(t), ctypes.c_double(max_corr_dist * max_corr_dist),
                                      sp.ctypes.data_as(dp), sq.ctypes.data_as(dp), sqp.ctypes.data_as(dp), ctypes.byref(mse), _ip(nn))
    L.llo_kdtree_free(tree)
    Rt = np.zeros(16, np.float32)
    if n >= 3:
        L.llo_umeyama_from_sums(ctypes.c_double(n), sp.ctypes.data_as(dp), sq.ctypes.data_as(dp), sqp.ctypes.data_as(dp), _fp(Rt))
    return dict(n=n, sp=sp, sq=sq, sqp=sqp.reshape(3, 3), mse=mse.value, nn=nn[:s.shape[0]], Rt=Rt.reshape(4, 4))


def set_trig_mode(mode: int):
    """0 = host libm sinf/cosf (reference-faithful on this machine), 1 = correctly rounded."""
    lib().llo_set_trig_mode(int(mode))


# ------------------------------------------------------------------ feature extraction

def std_sort_by_value(value, ind, depth_limit=-1):
    """Restatement of libstdc++ std::sort on (value, ind) records compared by value only (FA:699)."""
    v = np.ascontiguousarray(value, np.float32).copy(); i = np.ascontiguousarray(ind, np.uint32).copy()
    lib().llo_std_sort_by_value(_fp(v), i.ctypes.data_as(ctypes.c_void_p), v.shape[0], int(depth_limit))
    return v, i


class FeatureExtraction:
    """adjustDistortion (no IMU) + calculateSmoothness + markOccludedPoints + extractFeatures, FA:491-784, with the
    state the reference keeps between sweeps."""

    def __init__(self, n_scan=16, horizon=1800):
        L = lib()
        L.llo_features_create.restype = ctypes.c_void_p
        self.n_scan, self.horizon = n_scan, horizon
        self._h = ctypes.c_void_p(L.llo_features_create(n_scan, horizon))
        self._n = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().llo_features_destroy(self._h); self._h = None

    def extract(self, sw):
        """sw: SegmentedSweep fields. -> (sharp, less_sharp, flat, less_flat, adjusted cloud)"""
        cap = self.n_scan * self.horizon
        cloud = _pts(sw.cloud).copy(); n = cloud.shape[0]
        g = np.zeros(cap, np.uint8); g[:n] = sw.ground
        col = np.zeros(cap, np.uint32); col[:n] = sw.col
        rg = np.zeros(cap, np.float32); rg[:n] = sw.range
        sr = np.ascontiguousarray(sw.start_ring, np.int32); er = np.ascontiguousarray(sw.end_ring, np.int32)
        outs = [np.zeros((max(n, 1), 4), np.float32) for _ in range(4)]
        optr = (ctypes.c_void_p * 4)(*[o.ctypes.data for o in outs])
        cnt = (ctypes.c_int * 4)()
        lib().llo_features_extract(self._h, _fp(cloud), n, sr.ctypes.data_as(ctypes.c_void_p), er.ctypes.data_as(ctypes.c_void_p),
                                   ctypes.c_float(sw.start_ori), ctypes.c_float(sw.end_ori), ctypes.c_float(sw.ori_diff),
                                   g.ctypes.data_as(ctypes.c_void_p), col.ctypes.data_as(ctypes.c_void_p), _fp(rg), optr, cnt)
        self._n = n
        return tuple(outs[k][:cnt[k]].copy() for k in range(4)) + (cloud,)

    def point_state(self):
        n = self._n
        curv = np.zeros(n, np.float32); picked = np.zeros(n, np.int32); label = np.zeros(n, np.int32)
        lib().llo_features_get_state(self._h, n, _fp(curv), picked.ctypes.data_as(ctypes.c_void_p),
                                     label.ctypes.data_as(ctypes.c_void_p))
        return curv, picked, label


def transform_to_end(T, cloud):
    """TransformToEnd FA:885-953 on every point (no IMU messages); sin/cos per set_trig_mode."""
    t = np.ascontiguousarray(T, np.float32); c = _pts(cloud).copy()
    lib().llo_transform_to_end(_fp(t), _fp(c), c.shape[0])
    return c


# ------------------------------------------------------------------ imageProjection

class ImageProjection:
    """Restatement of the reference's imageProjection node (IP:181-368): raw sweep -> segmented cloud + cloud_info."""

    def __init__(self, n_scan=16, horizon=1800, ang_res_x=0.2, ang_res_y=2.0, ground_scan_ind=7):
        L = lib()
        L.llo_projection_create.restype = ctypes.c_void_p
        self.n_scan, self.horizon = n_scan, horizon
        self._h = ctypes.c_void_p(L.llo_projection_create(n_scan, horizon, ctypes.c_float(ang_res_x), ctypes.c_float(ang_res_y),
                                                          ground_scan_ind))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().llo_projection_destroy(self._h); self._h = None

    def process(self, cloud, ring):
        from lego_loam_b200 import synth
        pts = _pts(cloud); rg = np.ascontiguousarray(ring, np.uint16)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        lib().llo_projection_process(self._h, _fp(pts), vp(rg), pts.shape[0])
        cap = self.n_scan * self.horizon
        seg = np.zeros((cap, 4), np.float32); n = lib().llo_projection_get_cloud(self._h, 0, _fp(seg), cap); seg = seg[:n].copy()
        out = np.zeros((cap, 4), np.float32); m = lib().llo_projection_get_cloud(self._h, 1, _fp(out), cap); out = out[:m].copy()
        sr = np.zeros(self.n_scan, np.int32); er = np.zeros(self.n_scan, np.int32); ori = np.zeros(3, np.float32)
        g = np.zeros(max(n, 1), np.uint8); col = np.zeros(max(n, 1), np.uint32); r = np.zeros(max(n, 1), np.float32)
        lib().llo_projection_get_info(self._h, vp(sr), vp(er), _fp(ori), vp(g), vp(col), _fp(r), n)
        return synth.SegmentedSweep(seg, sr, er, float(ori[0]), float(ori[1]), float(ori[2]), g[:n], col[:n], r[:n], out)

    def images(self):
        n = self.n_scan * self.horizon
        rm = np.zeros(n, np.float32); gm = np.zeros(n, np.int8); lm = np.zeros(n, np.int32)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        lib().llo_projection_get_images(self._h, _fp(rm), vp(gm), vp(lm))
        shp = (self.n_scan, self.horizon)
        return rm.reshape(shp), gm.reshape(shp), lm.reshape(shp)
