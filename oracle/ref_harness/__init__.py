"""ORACLE tier B (test infrastructure): ctypes view of oracle/_ref/libref_{mo,fa}.so, i.e. the
UNMODIFIED reference sources (mapOptmization.cpp / featureAssociation.cpp) compiled against shim
headers (oracle/ref_harness/shim).  The classes mirror oracle.MapOptimization /
oracle.FeatureAssociation so the same tests drive the restatement and the reference itself.
The libraries can only be (re)built where /root/reference exists; on the GPU box the prebuilt
files travel with the snapshot.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_OUT = os.path.join(os.path.dirname(_HERE), "_ref")
_MO = os.path.join(_OUT, "libref_mo.so")
_FA = os.path.join(_OUT, "libref_fa.so")
_IP = os.path.join(_OUT, "libref_ip.so")
_libs = {}

c_float_p = ctypes.POINTER(ctypes.c_float)


def build() -> bool:
    if os.path.isdir("/root/reference/LeGO-LOAM/src"):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return available()


def available() -> bool:
    return os.path.exists(_MO) and os.path.exists(_FA) and os.path.exists(_IP)


def _lib(path, create):
    if path not in _libs:
        L = ctypes.CDLL(path)
        getattr(L, create).restype = ctypes.c_void_p
        _libs[path] = L
    return _libs[path]


def _pts(a):
    a = np.ascontiguousarray(a, np.float32)
    return a.reshape(-1, 4)


def _fp(a):
    return a.ctypes.data_as(c_float_p)


class MapOptimization:
    """class mapOptimization of the reference (MO:49), driven through its own member functions."""

    def __init__(self):
        self.L = _lib(_MO, "ref_mo_create")
        self._h = ctypes.c_void_p(self.L.ref_mo_create())

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.ref_mo_destroy(self._h); self._h = None

    def set_map_ds(self, c, s):
        c = _pts(c); s = _pts(s)
        self.L.ref_mo_set_map_ds(self._h, _fp(c), c.shape[0], _fp(s), s.shape[0])

    def set_map_raw(self, c, s):
        c = _pts(c); s = _pts(s)
        self.L.ref_mo_set_map_raw(self._h, _fp(c), c.shape[0], _fp(s), s.shape[0])

    def set_scan(self, c, s, o):
        c = _pts(c); s = _pts(s); o = _pts(o)
        self.L.ref_mo_set_scan(self._h, _fp(c), c.shape[0], _fp(s), s.shape[0], _fp(o), o.shape[0])

    @property
    def transformTobeMapped(self):
        t = np.zeros(6, np.float32); self.L.ref_mo_get_pose(self._h, _fp(t)); return t

    @transformTobeMapped.setter
    def transformTobeMapped(self, v):
        t = np.ascontiguousarray(v, np.float32); self.L.ref_mo_set_pose(self._h, _fp(t))

    def set_transform_sum(self, v):
        t = np.ascontiguousarray(v, np.float32); self.L.ref_mo_set_transform_sum(self._h, _fp(t))

    def bef_aft(self):
        b = np.zeros(6, np.float32); a = np.zeros(6, np.float32)
        self.L.ref_mo_get_bef_aft(self._h, _fp(b), _fp(a)); return b, a

    def degenerate(self):
        d = ctypes.c_int(0); P = np.zeros((6, 6), np.float32)
        self.L.ref_mo_get_degenerate(self._h, ctypes.byref(d), _fp(P)); return bool(d.value), P

    def downsampleCurrentScan(self): self.L.ref_mo_downsampleCurrentScan(self._h)
    def build_kdtrees(self): self.L.ref_mo_build_kdtrees(self._h)
    def clear_correspondences(self): self.L.ref_mo_clear_correspondences(self._h)
    def cornerOptimization(self, it): self.L.ref_mo_cornerOptimization(self._h, it)
    def surfOptimization(self, it): self.L.ref_mo_surfOptimization(self._h, it)
    def LMOptimization(self, it) -> bool: return bool(self.L.ref_mo_LMOptimization(self._h, it))
    def scan2MapOptimization(self): self.L.ref_mo_scan2MapOptimization(self._h)

    def _cloud(self, fn, which):
        n = fn(self._h, which, None, 0)
        out = np.zeros((max(n, 1), 4), np.float32)
        fn(self._h, which, _fp(out), n)
        return out[:n].copy()

    def scan_ds(self, which): return self._cloud(self.L.ref_mo_get_scan_ds, which)
    def map_ds(self, which): return self._cloud(self.L.ref_mo_get_map_ds, which)
    def map_raw(self, which): return self._cloud(self.L.ref_mo_get_map_raw, which)

    def correspondences(self):
        n = self.L.ref_mo_get_correspondences(self._h, None, None, 0)
        ori = np.zeros((max(n, 1), 4), np.float32); co = np.zeros((max(n, 1), 4), np.float32)
        self.L.ref_mo_get_correspondences(self._h, _fp(ori), _fp(co), n)
        return ori[:n].copy(), co[:n].copy()

    # the rest of run() MO:1503-1519 (sequence replays)
    def set_odometry(self, transform_sum, stamp: float):
        t = np.ascontiguousarray(transform_sum, np.float32)
        self.L.ref_mo_set_odometry(self._h, _fp(t), ctypes.c_double(stamp))

    def transformAssociateToMap(self): self.L.ref_mo_transformAssociateToMap(self._h)
    def transformUpdate(self): self.L.ref_mo_transformUpdate(self._h)

    def push_imu(self, stamp: float, roll: float, pitch: float):
        """imuHandler MO:643-652 after the quaternion -> roll / pitch conversion"""
        self.L.ref_mo_push_imu(self._h, ctypes.c_double(stamp), ctypes.c_double(roll), ctypes.c_double(pitch))

    @property
    def transformAftMapped(self):
        t = np.zeros(6, np.float32); self.L.ref_mo_get_aft_mapped(self._h, _fp(t)); return t

    def map_ds_sizes(self):
        a = ctypes.c_int(0); b = ctypes.c_int(0)
        self.L.ref_mo_map_ds_sizes(self._h, ctypes.byref(a), ctypes.byref(b)); return a.value, b.value
    def extractSurroundingKeyFrames(self): self.L.ref_mo_extractSurroundingKeyFrames(self._h)
    def saveKeyFramesAndFactor(self): self.L.ref_mo_saveKeyFramesAndFactor(self._h)
    def correctPoses(self): self.L.ref_mo_correctPoses(self._h)
    def clearCloud(self): self.L.ref_mo_clearCloud(self._h)
    def num_keyframes(self) -> int: return self.L.ref_mo_num_keyframes(self._h)

    def surrounding_ids(self) -> np.ndarray:
        n = self.L.ref_mo_get_surrounding_ids(self._h, None, 0)
        out = np.zeros(max(n, 1), np.int32)
        self.L.ref_mo_get_surrounding_ids(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), n)
        return out[:n].copy()

    def keypose6d(self, idx: int) -> np.ndarray:
        t = np.zeros(6, np.float32); self.L.ref_mo_get_keypose6d(self._h, int(idx), _fp(t)); return t

    def keyframe_cloud(self, idx: int, which: int) -> np.ndarray:
        n = self.L.ref_mo_get_keyframe_cloud(self._h, int(idx), which, None, 0)
        out = np.zeros((max(n, 1), 4), np.float32)
        self.L.ref_mo_get_keyframe_cloud(self._h, int(idx), which, _fp(out), n)
        return out[:n].copy()

    # ---- loop closure + global map
    def set_robot_pos(self, x, y, z): self.L.ref_mo_set_robot_pos(self._h, ctypes.c_float(x), ctypes.c_float(y), ctypes.c_float(z))
    def set_time(self, stamp: float): self.L.ref_mo_set_time(self._h, ctypes.c_double(stamp))
    def detectLoopClosure(self) -> bool: return bool(self.L.ref_mo_detectLoopClosure(self._h))

    def loop_ids(self):
        a = ctypes.c_int(0); b = ctypes.c_int(0)
        self.L.ref_mo_loop_ids(self._h, ctypes.byref(a), ctypes.byref(b)); return a.value, b.value

    def loop_cloud(self, which: int) -> np.ndarray:
        """0 latestSurfKeyFrameCloud, 1 nearHistorySurfKeyFrameCloud, 2 nearHistorySurfKeyFrameCloudDS, 3 globalMapKeyFramesDS"""
        return self._cloud(self.L.ref_mo_get_loop_cloud, which)

    def performLoopClosure(self) -> bool: return bool(self.L.ref_mo_performLoopClosure(self._h))

    def icp_last(self):
        """-> dict of the last pcl::IterativeClosestPoint::align the reference ran (restated PCL, oracle/llo_loop.c)"""
        T = np.zeros(16, np.float32); c = ctypes.c_int(0); it = ctypes.c_int(0); st = ctypes.c_int(0); f = ctypes.c_double(0)
        calls = self.L.ref_mo_icp_last(_fp(T), ctypes.byref(c), ctypes.byref(it), ctypes.byref(st), ctypes.byref(f))
        return dict(T=T.reshape(4, 4), converged=bool(c.value), iterations=it.value, state=st.value, fitness=f.value, calls=calls)

    def publishGlobalMap(self) -> np.ndarray:
        """runs publishGlobalMap (MO:758-800) with one subscriber; returns the key-frame ids it assembled, in order"""
        n = self.L.ref_mo_publishGlobalMap(self._h, None, 0)
        out = np.zeros(max(n, 1), np.int32)
        # the function is idempotent on the same state: second call fills the ids
        self.L.ref_mo_publishGlobalMap(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), n)
        return out[:n].copy()


def _sensor_lib(base, sensor):
    """libref_<node>.so is the reference as shipped (VLP-16, UT:62-68); libref_<node>_<sensor>.so the same sources with
    the sensor block of utility.h switched the way the reference's README prescribes (Makefile: select_sensor.awk)."""
    return base if sensor in (None, "vlp16") else base[:-3] + "_" + sensor + ".so"


def sensor_available(sensor) -> bool:
    return os.path.exists(_sensor_lib(_FA, sensor)) and os.path.exists(_sensor_lib(_IP, sensor))


class FeatureAssociation:
    """class FeatureAssociation of the reference (FA:37)."""

    def __init__(self, sensor=None):
        self.L = _lib(_sensor_lib(_FA, sensor), "ref_fa_create")
        self._h = ctypes.c_void_p(self.L.ref_fa_create())

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.ref_fa_destroy(self._h); self._h = None

    def set_last(self, c, s, force=False):
        c = _pts(c); s = _pts(s)
        self.L.ref_fa_set_last(self._h, _fp(c), c.shape[0], _fp(s), s.shape[0], int(force))

    def set_features(self, sharp, flat):
        a = _pts(sharp); b = _pts(flat)
        self.L.ref_fa_set_features(self._h, _fp(a), a.shape[0], _fp(b), b.shape[0])

    @property
    def transformCur(self):
        t = np.zeros(6, np.float32); self.L.ref_fa_get_transform(self._h, _fp(t)); return t

    @transformCur.setter
    def transformCur(self, v):
        t = np.ascontiguousarray(v, np.float32); self.L.ref_fa_set_transform(self._h, _fp(t))

    def degenerate(self):
        d = ctypes.c_int(0); P = np.zeros((3, 3), np.float32)
        self.L.ref_fa_get_degenerate(self._h, ctypes.byref(d), _fp(P)); return bool(d.value), P

    def clear_correspondences(self): self.L.ref_fa_clear_correspondences(self._h)

    # ---- feature extraction (FA:491-784)
    def n_scan(self): return self.L.ref_fa_n_scan()

    def set_segmented(self, sw):
        """sw: lego_loam_b200.synth.SegmentedSweep (or anything with its fields)."""
        pts = _pts(sw.cloud)
        sr = np.ascontiguousarray(sw.start_ring, np.int32); er = np.ascontiguousarray(sw.end_ring, np.int32)
        assert sr.shape[0] == self.n_scan()
        g = np.ascontiguousarray(sw.ground, np.uint8); col = np.ascontiguousarray(sw.col, np.uint32)
        rg = np.ascontiguousarray(sw.range, np.float32)
        self._n_seg = pts.shape[0]
        self.L.ref_fa_set_segmented(self._h, _fp(pts), pts.shape[0], sr.ctypes.data_as(ctypes.c_void_p),
                                    er.ctypes.data_as(ctypes.c_void_p), ctypes.c_float(sw.start_ori),
                                    ctypes.c_float(sw.end_ori), ctypes.c_float(sw.ori_diff),
                                    g.ctypes.data_as(ctypes.c_void_p), col.ctypes.data_as(ctypes.c_void_p), _fp(rg))

    def extract_features(self):
        """adjustDistortion, calculateSmoothness, markOccludedPoints, extractFeatures (FA:1827-1833)."""
        self.L.ref_fa_extract_features(self._h)

    def feature_cloud(self, which, cap=40000):
        """0 cornerPointsSharp, 1 cornerPointsLessSharp, 2 surfPointsFlat, 3 surfPointsLessFlat, 4 segmentedCloud,
        5 laserCloudCornerLast, 6 laserCloudSurfLast."""
        out = np.zeros((cap, 4), np.float32)
        n = self.L.ref_fa_get_cloud(self._h, int(which), _fp(out), cap)
        assert n <= cap
        return out[:n].copy()

    # ---- IMU (FA:417-448 and the IMU branches of adjustDistortion / TransformToEnd / updateInitialGuess)
    def push_imu(self, stamp, roll, pitch, yaw, lin_acc, ang_vel):
        """imuHandler after the quaternion -> roll / pitch / yaw conversion"""
        la = (ctypes.c_double * 3)(*[float(x) for x in lin_acc]); av = (ctypes.c_double * 3)(*[float(x) for x in ang_vel])
        self.L.ref_fa_push_imu(self._h, ctypes.c_double(stamp), ctypes.c_double(roll), ctypes.c_double(pitch), ctypes.c_double(yaw), la, av)

    def set_time_scan_cur(self, t): self.L.ref_fa_set_time_scan_cur(self._h, ctypes.c_double(t))
    def updateInitialGuess(self): self.L.ref_fa_updateInitialGuess(self._h)

    def imu_state(self):
        o = np.zeros(24, np.float32); self.L.ref_fa_get_imu_state(self._h, _fp(o)); return o

    def imu_entry(self, idx=-1):
        t = ctypes.c_double(0); o = np.zeros(12, np.float32)
        i = self.L.ref_fa_get_imu_entry(self._h, int(idx), ctypes.byref(t), _fp(o))
        return i, t.value, o

    def publishCloudsLast(self):
        """FA:1759-1815; afterwards feature_cloud(5) / feature_cloud(6) = laserCloudCornerLast / laserCloudSurfLast."""
        self.L.ref_fa_publishCloudsLast(self._h)

    def point_state(self):
        n = self._n_seg
        curv = np.zeros(n, np.float32); picked = np.zeros(n, np.int32); label = np.zeros(n, np.int32)
        self.L.ref_fa_get_point_state(self._h, n, _fp(curv), picked.ctypes.data_as(ctypes.c_void_p),
                                      label.ctypes.data_as(ctypes.c_void_p))
        return curv, picked, label
    def findCorrespondingCornerFeatures(self, it): self.L.ref_fa_findCorrespondingCornerFeatures(self._h, it)
    def findCorrespondingSurfFeatures(self, it): self.L.ref_fa_findCorrespondingSurfFeatures(self._h, it)
    def calculateTransformationSurf(self, it) -> bool: return bool(self.L.ref_fa_calculateTransformationSurf(self._h, it))
    def calculateTransformationCorner(self, it) -> bool: return bool(self.L.ref_fa_calculateTransformationCorner(self._h, it))
    def updateTransformation(self): self.L.ref_fa_updateTransformation(self._h)
    def integrateTransformation(self): self.L.ref_fa_integrateTransformation(self._h)

    @property
    def transformSum(self):
        t = np.zeros(6, np.float32); self.L.ref_fa_get_transform_sum(self._h, _fp(t)); return t

    def correspondences(self):
        n = self.L.ref_fa_get_correspondences(self._h, None, None, 0)
        ori = np.zeros((max(n, 1), 4), np.float32); co = np.zeros((max(n, 1), 4), np.float32)
        self.L.ref_fa_get_correspondences(self._h, _fp(ori), _fp(co), n)
        return ori[:n].copy(), co[:n].copy()

    def search_ind(self, which):
        n = self.L.ref_fa_get_search_ind(self._h, which, None, None, None, 0)
        a = np.zeros(max(n, 1), np.float32); b = a.copy(); c = a.copy()
        self.L.ref_fa_get_search_ind(self._h, which, _fp(a), _fp(b), _fp(c), n)
        return a[:n].copy(), b[:n].copy(), c[:n].copy()


class ImageProjection:
    """class ImageProjection of the reference (imageProjection.cpp:37): raw sweep -> segmented cloud + cloud_info +
    outlier cloud, through its own member functions (cloudHandler IP:181-197 without publishing)."""

    def __init__(self, sensor=None):
        self.L = _lib(_sensor_lib(_IP, sensor), "ref_ip_create")
        self._h = ctypes.c_void_p(self.L.ref_ip_create())
        self.n_scan = self.L.ref_ip_n_scan(); self.horizon = self.L.ref_ip_horizon_scan()

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.ref_ip_destroy(self._h); self._h = None

    def process(self, cloud, ring):
        """cloud (n,4) in the LIDAR frame in firing order, ring (n,) -> lego_loam_b200.synth.SegmentedSweep"""
        from lego_loam_b200 import synth
        pts = _pts(cloud); rg = np.ascontiguousarray(ring, np.uint16)
        assert rg.shape[0] == pts.shape[0]
        self.L.ref_ip_process(self._h, _fp(pts), rg.ctypes.data_as(ctypes.c_void_p), pts.shape[0])
        cap = self.n_scan * self.horizon
        seg = np.zeros((cap, 4), np.float32); n = self.L.ref_ip_get_cloud(self._h, 0, _fp(seg), cap); seg = seg[:n].copy()
        out = np.zeros((cap, 4), np.float32); m = self.L.ref_ip_get_cloud(self._h, 1, _fp(out), cap); out = out[:m].copy()
        sr = np.zeros(self.n_scan, np.int32); er = np.zeros(self.n_scan, np.int32); ori = np.zeros(3, np.float32)
        g = np.zeros(max(n, 1), np.uint8); col = np.zeros(max(n, 1), np.uint32); r = np.zeros(max(n, 1), np.float32)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        self.L.ref_ip_get_info(self._h, vp(sr), vp(er), _fp(ori), vp(g), vp(col), _fp(r), n)
        return synth.SegmentedSweep(seg, sr, er, float(ori[0]), float(ori[1]), float(ori[2]), g[:n], col[:n], r[:n], out)

    def images(self):
        n = self.n_scan * self.horizon
        rm = np.zeros(n, np.float32); gm = np.zeros(n, np.int8); lm = np.zeros(n, np.int32)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        self.L.ref_ip_get_images(self._h, _fp(rm), vp(gm), vp(lm))
        shp = (self.n_scan, self.horizon)
        return rm.reshape(shp), gm.reshape(shp), lm.reshape(shp)


def std_sort(value, ind, depth_limit=-1):
    """libstdc++ std::sort of (value, ind) records by value (FA:699); depth_limit >= 0: its introsort loop with that limit."""
    L = _lib(_FA, "ref_fa_create")
    v = np.ascontiguousarray(value, np.float32).copy(); i = np.ascontiguousarray(ind, np.uint64).copy()
    L.ref_fa_std_sort(_fp(v), i.ctypes.data_as(ctypes.c_void_p), v.shape[0], int(depth_limit))
    return v, i
