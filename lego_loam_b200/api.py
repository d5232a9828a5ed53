"""ctypes binding of the C ABI in include/llb200.h (libllb200.so).

This is the Python host-side mirror used by tests/ and bench.py; the C++ adapter
classes with the reference's own member-function names live in host/.  There is
no CPU fallback: if the CUDA library is missing or no sm_100 device is present,
loading / context creation raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libllb200.so")
CSRC = os.path.join(_PKG, "csrc")

LLB_OK, LLB_ERR_INVALID, LLB_ERR_CUDA, LLB_ERR_NO_DEVICE, LLB_ERR_CAPACITY, LLB_ERR_STATE = range(6)
_STATUS = {0: "OK", 1: "INVALID", 2: "CUDA", 3: "NO_DEVICE", 4: "CAPACITY", 5: "STATE"}


class LlbError(RuntimeError):
    def __init__(self, status: int, msg: str = ""):
        super().__init__(f"llb200 status {_STATUS.get(status, status)}: {msg}")
        self.status = status


class Params(ctypes.Structure):
    _fields_ = [
        ("corner_leaf", ctypes.c_float), ("surf_leaf", ctypes.c_float), ("outlier_leaf", ctypes.c_float),
        ("knn_max_sqdist", ctypes.c_float),
        ("s2m_max_iterations", ctypes.c_int), ("s2m_min_correspondences", ctypes.c_int),
        ("s2m_degeneracy_thresh", ctypes.c_float), ("s2m_converge_deg", ctypes.c_float),
        ("s2m_converge_cm", ctypes.c_float),
        ("corner_map_min", ctypes.c_int), ("surf_map_min", ctypes.c_int),
        ("odom_nearest_sqdist", ctypes.c_float), ("odom_max_iterations", ctypes.c_int),
        ("odom_min_correspondences", ctypes.c_int), ("odom_degeneracy_thresh", ctypes.c_float),
        ("odom_converge_deg", ctypes.c_float), ("odom_converge_cm", ctypes.c_float),
        ("max_grid_cells", ctypes.c_int),
        ("pin_host_clouds", ctypes.c_int),
        ("s2m_max_ctas", ctypes.c_int),
    ]


class Stats(ctypes.Structure):
    _fields_ = [
        ("iterations", ctypes.c_int), ("converged", ctypes.c_int), ("n_correspondences", ctypes.c_int),
        ("is_degenerate", ctypes.c_int), ("skipped", ctypes.c_int),
        ("n_corner_ds", ctypes.c_int), ("n_surf_ds", ctypes.c_int), ("device_ms", ctypes.c_float),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = [
    "llb_abi_version", "llb_params_default", "llb_create", "llb_destroy", "llb_last_error", "llb_stream",
    "llb_synchronize", "llb_reserve", "llb_voxel_downsample", "llb_map_set_ds", "llb_map_set_raw", "llb_map_get_ds",
    "llb_scan_set", "llb_downsample_current_scan", "llb_scan_get_ds", "llb_s2m_iterate", "llb_s2m_optimize",
    "llb_get_correspondences", "llb_get_knn", "llb_get_normal_equations", "llb_get_degeneracy",
    "llb_set_degeneracy", "llb_odom_set_last", "llb_odom_set_features", "llb_odom_optimize", "llb_odom_iterate",
    "llb_odom_get_correspondences", "llb_odom_get_search_ind", "llb_odom_get_degeneracy",
    "llb_map_set_ds_dev", "llb_map_set_raw_dev", "llb_scan_set_dev", "llb_s2m_optimize_dev",
    "llb_s2m_accumulate", "llb_s2m_solve", "llb_s2m_pose_set", "llb_s2m_pose_get", "llb_launch_count",
    "llb_s2m_time_iteration", "llb_s2m_get_profile", "llb_s2m_get_cta_profile",
    "llb_s2m_optimize_async", "llb_s2m_result",
    "llb_p2p_export", "llb_p2p_import", "llb_s2m_optimize_sharded",
    "llb_map_set_raw_sharded", "llb_map_set_raw_sharded_dev", "llb_map_shard_info", "llb_map_shard_set_global", "llb_shard_plan",
    "llb_loop_params_default", "llb_loop_set_clouds", "llb_loop_set_clouds_host", "llb_loop_icp", "llb_loop_get_cloud",
    "llb_loop_get_nn", "llb_global_map_assemble",
    "llb_features_init", "llb_features_extract", "llb_features_get", "llb_features_get_state", "llb_features_to_odometry", "llb_features_get_profile", "llb_features_publish_last",
    "llb_features_set_imu", "llb_features_get_imu", "llb_features_publish_last_imu",
    "llb_projection_init", "llb_projection_process", "llb_projection_get_cloud", "llb_projection_get_info",
    "llb_projection_get_images", "llb_projection_to_features",
    "llb_batch_features_init", "llb_batch_features_extract", "llb_batch_features_get",
    "llb_keyframe_add", "llb_keyframe_add_clouds", "llb_keyframe_count", "llb_keyframe_clear", "llb_map_assemble",
    "llb_map_get_raw",
    "llb_batch_create", "llb_batch_destroy", "llb_batch_last_error", "llb_batch_stream", "llb_batch_slots",
    "llb_batch_launch_count", "llb_batch_scan_set", "llb_batch_map_set_ds", "llb_batch_scan_set_dev",
    "llb_batch_map_set_ds_dev", "llb_batch_scan_set_all", "llb_batch_map_set_ds_all", "llb_batch_scan_set_dev_all",
    "llb_batch_map_set_ds_dev_all", "llb_batch_register", "llb_batch_register_async", "llb_batch_result",
    "llb_batch_enable_keyframes", "llb_batch_keyframe_add", "llb_batch_keyframe_count", "llb_batch_map_assemble", "llb_batch_map_assemble_all",
    "llb_batch_map_get", "llb_batch_odom_set", "llb_batch_odom_optimize", "llb_batch_scan_get_ds", "llb_batch_get_degeneracy", "llb_batch_set_profile", "llb_batch_get_profile",
]

_lib = None


def build(force: bool = False) -> str:
    """Compile the CUDA library for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(_PKG, "..", "include", "llb200.h"))
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", CSRC, "-s", "-j8"])
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LlbError(LLB_ERR_NO_DEVICE, f"{LIB_PATH} is not built (run __graft_entry__.build())")
        L = ctypes.CDLL(LIB_PATH)
        L.llb_last_error.restype = ctypes.c_char_p
        L.llb_stream.restype = ctypes.c_void_p
        L.llb_launch_count.restype = ctypes.c_longlong
        L.llb_batch_last_error.restype = ctypes.c_char_p
        L.llb_batch_stream.restype = ctypes.c_void_p
        L.llb_batch_launch_count.restype = ctypes.c_longlong
        for name in EXPORTS:
            getattr(L, name)   # raises AttributeError if a declared symbol is missing
        _lib = L
    return _lib


IMU_QUEUE = 200
_LIBM = None


def _libm():
    """the C library's cosf / sinf (numpy's float32 cos is its own SIMD routine, not the reference's libm)"""
    global _LIBM
    if _LIBM is None:
        import ctypes.util
        _LIBM = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        for f in (_LIBM.cosf, _LIBM.sinf):
            f.restype = ctypes.c_float; f.argtypes = [ctypes.c_float]
    return _LIBM


class ImuQueue(ctypes.Structure):
    """llb_imu_queue: the IMU ring buffers of FeatureAssociation (FA:82-135) as they stand when a sweep arrives."""
    _fields_ = [("time", ctypes.c_double * IMU_QUEUE),
                ("roll", ctypes.c_float * IMU_QUEUE), ("pitch", ctypes.c_float * IMU_QUEUE), ("yaw", ctypes.c_float * IMU_QUEUE),
                ("velo", (ctypes.c_float * IMU_QUEUE) * 3), ("shift", (ctypes.c_float * IMU_QUEUE) * 3),
                ("angular", (ctypes.c_float * IMU_QUEUE) * 3),
                ("time_scan_cur", ctypes.c_double), ("pointer_last", ctypes.c_int), ("pointer_last_iteration", ctypes.c_int)]


class ImuSweep(ctypes.Structure):
    """llb_imu_sweep: the members adjustDistortion's IMU branch leaves behind (FA:556-611)."""
    _fields_ = [("start", ctypes.c_float * 9), ("angular_cur", ctypes.c_float * 3), ("cur", ctypes.c_float * 3),
                ("velo_from_start_cur", ctypes.c_float * 3), ("valid", ctypes.c_int), ("has_velo", ctypes.c_int)]


class ImuEnd(ctypes.Structure):
    """llb_imu_end: IMU terms of TransformToEnd (FA:927-950)."""
    _fields_ = [("cs_start", ctypes.c_float * 6), ("shift_from_start", ctypes.c_float * 3), ("last", ctypes.c_float * 3)]


class SegmentedCloud(ctypes.Structure):
    """llb_segmented_cloud"""
    _fields_ = [("cloud", ctypes.c_void_p), ("n", ctypes.c_int), ("start_ring", ctypes.c_void_p),
                ("end_ring", ctypes.c_void_p), ("start_orientation", ctypes.c_float), ("end_orientation", ctypes.c_float),
                ("orientation_diff", ctypes.c_float), ("ground_flag", ctypes.c_void_p), ("col_ind", ctypes.c_void_p),
                ("range", ctypes.c_void_p)]


def segmented_struct(sw, seg=None):
    """SegmentedSweep-like object -> (llb_segmented_cloud, arrays that must stay alive during the call)"""
    seg = SegmentedCloud() if seg is None else seg
    keep = [sw.cloud32 if hasattr(sw, "cloud32") else to_pcl(sw.cloud), np.ascontiguousarray(sw.start_ring, np.int32),
            np.ascontiguousarray(sw.end_ring, np.int32), np.ascontiguousarray(sw.ground, np.uint8),
            np.ascontiguousarray(sw.col, np.uint32), np.ascontiguousarray(sw.range, np.float32)]
    seg.cloud = _vp(keep[0]); seg.n = keep[0].shape[0]
    seg.start_ring = _vp(keep[1]); seg.end_ring = _vp(keep[2])
    seg.start_orientation = sw.start_ori; seg.end_orientation = sw.end_ori; seg.orientation_diff = sw.ori_diff
    seg.ground_flag = _vp(keep[3]); seg.col_ind = _vp(keep[4]); seg.range = _vp(keep[5])
    return seg, keep


def to_pcl(pts) -> np.ndarray:
    """(n,4) float32 {x,y,z,intensity} -> (n,8) float32 in pcl::PointXYZI's 32-byte layout."""
    a = np.ascontiguousarray(pts, np.float32).reshape(-1, 4)
    out = np.zeros((a.shape[0], 8), np.float32)
    out[:, :3] = a[:, :3]
    out[:, 3] = 1.0
    out[:, 4] = a[:, 3]
    return out


def from_pcl(p32: np.ndarray) -> np.ndarray:
    out = np.empty((p32.shape[0], 4), np.float32)
    out[:, :3] = p32[:, :3]
    out[:, 3] = p32[:, 4]
    return out


def _vp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _fp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


class Context:
    """One llb_ctx: one CUDA stream + workspaces on one device."""

    def __init__(self, device: int = 0, params: Params | None = None):
        L = lib()
        self._h = ctypes.c_void_p()
        if params is None:
            params = default_params()
        self.params = params
        rc = L.llb_create(ctypes.byref(params), int(device), ctypes.byref(self._h))
        if rc != LLB_OK:
            self._h = ctypes.c_void_p()
            raise LlbError(rc, "llb_create failed (no CPU fallback exists)")
        self.device = device
        # converted host clouds of the convenience setters: the library may page-lock them in place (pin_host_clouds) and
        # copy from them after the call has returned, so they live until the next cloud of the same kind replaces them
        self._held = {}

    def _hold(self, kind: str, arrays):
        self._held[kind] = arrays

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().llb_destroy(self._h)
            self._h = ctypes.c_void_p()
        self._held = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != LLB_OK:
            raise LlbError(rc, (lib().llb_last_error(self._h) or b"").decode())

    @property
    def stream(self) -> int:
        return lib().llb_stream(self._h) or 0

    def synchronize(self):
        self._ck(lib().llb_synchronize(self._h))

    def launch_count(self) -> int:
        return int(lib().llb_launch_count(self._h))

    def reserve(self, max_scan_points: int, max_raw_map_points: int, max_keyframes: int = 0):
        self._ck(lib().llb_reserve(self._h, int(max_scan_points), int(max_raw_map_points), int(max_keyframes)))

    # ---- voxel
    def voxel_downsample(self, pts, leaf: float) -> np.ndarray:
        src = to_pcl(pts)
        n = src.shape[0]
        out = np.zeros((max(n, 1), 8), np.float32)
        m = ctypes.c_int(0)
        self._ck(lib().llb_voxel_downsample(self._h, _vp(src), n, ctypes.c_float(leaf), _vp(out), n, ctypes.byref(m)))
        return from_pcl(out[:m.value])

    # ---- map
    def map_set_ds(self, corner_ds, surf_ds):
        c = to_pcl(corner_ds); s = to_pcl(surf_ds)
        self._hold("map_ds", (c, s))
        self._ck(lib().llb_map_set_ds(self._h, _vp(c), c.shape[0], _vp(s), s.shape[0]))

    def map_set_ds_pcl(self, c32: np.ndarray, s32: np.ndarray):
        self._ck(lib().llb_map_set_ds(self._h, _vp(c32), c32.shape[0], _vp(s32), s32.shape[0]))

    def map_set_raw(self, corner, surf):
        c = to_pcl(corner); s = to_pcl(surf)
        self._hold("map_raw", (c, s))
        self._ck(lib().llb_map_set_raw(self._h, _vp(c), c.shape[0], _vp(s), s.shape[0]))

    def map_set_raw_pcl(self, c32: np.ndarray, s32: np.ndarray):
        self._ck(lib().llb_map_set_raw(self._h, _vp(c32), c32.shape[0], _vp(s32), s32.shape[0]))

    def map_get_ds(self, which: int) -> np.ndarray:
        n = ctypes.c_int(0)
        self._ck(lib().llb_map_get_ds(self._h, which, None, 0, ctypes.byref(n)))
        out = np.zeros((max(n.value, 1), 8), np.float32)
        self._ck(lib().llb_map_get_ds(self._h, which, _vp(out), n.value, ctypes.byref(n)))
        return from_pcl(out[:n.value])

    # ---- device-resident key-frame store
    def keyframe_add(self) -> int:
        k = ctypes.c_int(-1)
        self._ck(lib().llb_keyframe_add(self._h, ctypes.byref(k)))
        return k.value

    def keyframe_add_clouds(self, corner_ds, surf_ds, outlier_ds) -> int:
        c = to_pcl(corner_ds); s = to_pcl(surf_ds); o = to_pcl(outlier_ds)
        k = ctypes.c_int(-1)
        self._ck(lib().llb_keyframe_add_clouds(self._h, _vp(c), c.shape[0], _vp(s), s.shape[0], _vp(o), o.shape[0],
                                               ctypes.byref(k)))
        return k.value

    def keyframe_count(self) -> int:
        n = ctypes.c_int(0)
        self._ck(lib().llb_keyframe_count(self._h, ctypes.byref(n)))
        return n.value

    def keyframe_clear(self):
        self._ck(lib().llb_keyframe_clear(self._h))

    def map_assemble(self, ids, poses6d):
        """ids: key-frame indices in order; poses6d: (n, 6) {roll, pitch, yaw, x, y, z} (cloudKeyPoses6D)"""
        i = np.ascontiguousarray(ids, np.int32)
        p = np.ascontiguousarray(poses6d, np.float32).reshape(-1, 6)
        assert p.shape[0] == i.shape[0]
        self._ck(lib().llb_map_assemble(self._h, i.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _fp(p), int(i.shape[0])))

    def map_get_raw(self, which: int) -> np.ndarray:
        n = ctypes.c_int(0)
        self._ck(lib().llb_map_get_raw(self._h, which, None, 0, ctypes.byref(n)))
        out = np.zeros((max(n.value, 1), 8), np.float32)
        self._ck(lib().llb_map_get_raw(self._h, which, _vp(out), n.value, ctypes.byref(n)))
        return from_pcl(out[:n.value])

    # ---- scan
    def scan_set(self, corner_last, surf_last, outlier_last):
        c = to_pcl(corner_last); s = to_pcl(surf_last); o = to_pcl(outlier_last)
        self._hold("scan", (c, s, o))                       # the library may page-lock / DMA from these after the call returns
        self.scan_set_pcl(c, s, o)

    def scan_set_pcl(self, c32, s32, o32):
        self._ck(lib().llb_scan_set(self._h, _vp(c32), c32.shape[0], _vp(s32), s32.shape[0], _vp(o32), o32.shape[0]))

    def downsample_current_scan(self, want_counts: bool = True):
        if not want_counts:
            self._ck(lib().llb_downsample_current_scan(self._h, None))
            return None
        cnt = (ctypes.c_int * 4)()
        self._ck(lib().llb_downsample_current_scan(self._h, cnt))
        return list(cnt)

    def scan_get_ds(self, which: int) -> np.ndarray:
        n = ctypes.c_int(0)
        self._ck(lib().llb_scan_get_ds(self._h, which, None, 0, ctypes.byref(n)))
        out = np.zeros((max(n.value, 1), 8), np.float32)
        self._ck(lib().llb_scan_get_ds(self._h, which, _vp(out), n.value, ctypes.byref(n)))
        return from_pcl(out[:n.value])

    # ---- scan-to-map
    def s2m_iterate(self, T, it: int):
        t = np.ascontiguousarray(T, np.float32).copy()
        conv = ctypes.c_int(0); nc = ctypes.c_int(0)
        self._ck(lib().llb_s2m_iterate(self._h, _fp(t), it, ctypes.byref(conv), ctypes.byref(nc)))
        return t, bool(conv.value), nc.value

    def s2m_optimize(self, T):
        t = np.ascontiguousarray(T, np.float32).copy()
        st = Stats()
        self._ck(lib().llb_s2m_optimize(self._h, _fp(t), ctypes.byref(st)))
        return t, st

    def s2m_optimize_async(self, T):
        t = np.ascontiguousarray(T, np.float32)
        self._async_T = t.copy()
        self._ck(lib().llb_s2m_optimize_async(self._h, _fp(t)))

    def s2m_result(self):
        t = self._async_T.copy()          # left untouched when the map-size guard skips the registration
        st = Stats()
        self._ck(lib().llb_s2m_result(self._h, _fp(t), ctypes.byref(st)))
        return t, st

    def get_correspondences(self):
        n = ctypes.c_int(0)
        self._ck(lib().llb_get_correspondences(self._h, None, None, 0, ctypes.byref(n)))
        ori = np.zeros((max(n.value, 1), 8), np.float32); co = np.zeros((max(n.value, 1), 8), np.float32)
        self._ck(lib().llb_get_correspondences(self._h, _vp(ori), _vp(co), n.value, ctypes.byref(n)))
        return from_pcl(ori[:n.value]), from_pcl(co[:n.value])

    def get_knn(self, which: int):
        n = ctypes.c_int(0)
        self._ck(lib().llb_get_knn(self._h, which, None, None, 0, ctypes.byref(n)))
        idx = np.zeros((max(n.value, 1), 5), np.int32); d2 = np.zeros((max(n.value, 1), 5), np.float32)
        self._ck(lib().llb_get_knn(self._h, which, _vp(idx), _vp(d2), n.value, ctypes.byref(n)))
        return idx[:n.value], d2[:n.value]

    def get_normal_equations(self):
        A = np.zeros((6, 6), np.float32); B = np.zeros(6, np.float32); X = np.zeros(6, np.float32)
        self._ck(lib().llb_get_normal_equations(self._h, _fp(A), _fp(B), _fp(X)))
        return A, B, X

    def get_degeneracy(self):
        d = ctypes.c_int(0); P = np.zeros((6, 6), np.float32)
        self._ck(lib().llb_get_degeneracy(self._h, ctypes.byref(d), _fp(P)))
        return bool(d.value), P

    def set_degeneracy(self, deg: bool, P):
        P = np.ascontiguousarray(P, np.float32)
        self._ck(lib().llb_set_degeneracy(self._h, int(deg), _fp(P)))

    # ---- odometry
    def odom_set_last(self, corner_last, surf_last):
        c = to_pcl(corner_last); s = to_pcl(surf_last)
        self._hold("odom_last", (c, s))
        self._ck(lib().llb_odom_set_last(self._h, _vp(c), c.shape[0], _vp(s), s.shape[0]))

    # ---- imageProjection (IP:181-460)
    def projection_init(self, n_scan: int, horizon_scan: int, ang_res_x: float, ang_res_y: float, ground_scan_ind: int):
        self._ip_shape = (int(n_scan), int(horizon_scan))
        self._ck(lib().llb_projection_init(self._h, int(n_scan), int(horizon_scan), ctypes.c_float(ang_res_x),
                                           ctypes.c_float(ang_res_y), int(ground_scan_ind)))

    def projection_process(self, cloud, ring):
        """cloud (n, 4) in firing order, ring (n,) uint16 -> (n_segmented, n_outlier, device ms)"""
        c32 = to_pcl(cloud); rg = np.ascontiguousarray(ring, np.uint16)
        assert rg.shape[0] == c32.shape[0]
        ns = ctypes.c_int(0); no = ctypes.c_int(0); ms = ctypes.c_float(0)
        self._ck(lib().llb_projection_process(self._h, _vp(c32), rg.ctypes.data_as(ctypes.c_void_p), c32.shape[0],
                                              ctypes.byref(ns), ctypes.byref(no), ctypes.byref(ms)))
        return ns.value, no.value, ms.value

    def projection_get_cloud(self, which: int) -> np.ndarray:
        n = ctypes.c_int(0)
        self._ck(lib().llb_projection_get_cloud(self._h, which, None, 0, ctypes.byref(n)))
        out = np.zeros((max(n.value, 1), 8), np.float32)
        self._ck(lib().llb_projection_get_cloud(self._h, which, _vp(out), n.value, ctypes.byref(n)))
        return from_pcl(out[:n.value])

    def projection_get_sweep(self):
        """-> lego_loam_b200.synth.SegmentedSweep of the last projection_process (segmentedCloud + cloud_info + outliers)"""
        from . import synth
        seg = self.projection_get_cloud(0); out = self.projection_get_cloud(1)
        n = seg.shape[0]; N = self._ip_shape[0]
        sr = np.zeros(N, np.int32); er = np.zeros(N, np.int32); ori = np.zeros(3, np.float32)
        g = np.zeros(max(n, 1), np.uint8); col = np.zeros(max(n, 1), np.uint32); r = np.zeros(max(n, 1), np.float32)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        self._ck(lib().llb_projection_get_info(self._h, vp(sr), vp(er), _fp(ori), vp(g), vp(col), _fp(r), max(n, 1)))
        return synth.SegmentedSweep(seg, sr, er, float(ori[0]), float(ori[1]), float(ori[2]), g[:n], col[:n], r[:n], out)

    def projection_get_images(self):
        N, H = self._ip_shape
        rm = np.zeros((N, H), np.float32); gm = np.zeros((N, H), np.int8); lm = np.zeros((N, H), np.int32)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        self._ck(lib().llb_projection_get_images(self._h, _fp(rm), vp(gm), vp(lm)))
        return rm, gm, lm

    def projection_to_features(self):
        cnt = (ctypes.c_int * 4)(); ms = ctypes.c_float(0)
        self._ck(lib().llb_projection_to_features(self._h, cnt, ctypes.byref(ms)))
        return list(cnt), ms.value

    # ---- feature extraction (FA:491-784)
    def features_init(self, n_scan: int, horizon_scan: int):
        self._ck(lib().llb_features_init(self._h, int(n_scan), int(horizon_scan)))

    def features_extract(self, sw):
        """sw: a SegmentedSweep (lego_loam_b200.synth) or anything with its fields.  -> (counts[4], device_ms)"""
        seg, keep = segmented_struct(sw)
        counts = (ctypes.c_int * 4)(); ms = ctypes.c_float(0)
        self._ck(lib().llb_features_extract(self._h, ctypes.byref(seg), counts, ctypes.byref(ms)))
        return list(counts), float(ms.value)

    def features_get(self, which: int):
        """0 cornerPointsSharp, 1 cornerPointsLessSharp, 2 surfPointsFlat, 3 surfPointsLessFlat, 4 adjusted segmentedCloud"""
        n = ctypes.c_int(0)
        self._ck(lib().llb_features_get(self._h, int(which), None, 0, ctypes.byref(n)))
        out = np.zeros((max(n.value, 1), 8), np.float32)
        self._ck(lib().llb_features_get(self._h, int(which), _vp(out), out.shape[0], ctypes.byref(n)))
        return from_pcl(out[:n.value])

    def features_get_state(self, n: int):
        curv = np.zeros(max(n, 1), np.float32); picked = np.zeros(max(n, 1), np.int32); label = np.zeros(max(n, 1), np.int32)
        self._ck(lib().llb_features_get_state(self._h, _vp(curv), _vp(picked), _vp(label), curv.shape[0]))
        return curv[:n], picked[:n], label[:n]

    def features_publish_last(self, transformCur):
        """TransformToEnd of the less-sharp / less-flat clouds -> laserCloudCornerLast / laserCloudSurfLast (FA:1759-1788)"""
        t = np.ascontiguousarray(transformCur, np.float32)
        self._ck(lib().llb_features_publish_last(self._h, _fp(t)))

    def features_set_imu(self, queue):
        """ring buffers (ImuQueue) for the next features_extract / projection_to_features; None: no IMU data"""
        self._ck(lib().llb_features_set_imu(self._h, ctypes.byref(queue) if queue is not None else None))

    def features_get_imu(self):
        out = ImuSweep()
        self._ck(lib().llb_features_get_imu(self._h, ctypes.byref(out)))
        return out

    def features_publish_last_imu(self, transformCur, start_rpy, shift_from_start, last_rpy):
        """publishCloudsLast with the IMU terms of TransformToEnd: imuRoll/Pitch/YawStart, imuShiftFromStartX/Y/Z,
        imuRoll/Pitch/YawLast (the cos / sin of the start angles are taken here, in float, as FA:317-324 does)"""
        e = ImuEnd()
        for k in range(3):
            a = ctypes.c_float(float(np.float32(start_rpy[k])))
            e.cs_start[2 * k] = _libm().cosf(a); e.cs_start[2 * k + 1] = _libm().sinf(a)
            e.shift_from_start[k] = float(np.float32(shift_from_start[k])); e.last[k] = float(np.float32(last_rpy[k]))
        t = np.ascontiguousarray(transformCur, np.float32)
        self._ck(lib().llb_features_publish_last_imu(self._h, _fp(t), ctypes.byref(e)))

    def features_get_profile(self):
        cyc = (ctypes.c_int * 10)()
        self._ck(lib().llb_features_get_profile(self._h, cyc))
        return {"sort_cycles": int(cyc[0]), "pick_cycles": int(cyc[1]), "partition_cycles": int(cyc[2]),
                "leaf_cycles": int(cyc[3]), "edge_pick_cycles": int(cyc[4]), "flat_pick_cycles": int(cyc[5])}

    def features_to_odometry(self):
        self._ck(lib().llb_features_to_odometry(self._h))

    def odom_set_features(self, sharp, flat):
        c = to_pcl(sharp); s = to_pcl(flat)
        self._hold("odom_features", (c, s))
        self._ck(lib().llb_odom_set_features(self._h, _vp(c), c.shape[0], _vp(s), s.shape[0]))

    def odom_optimize(self, T):
        t = np.ascontiguousarray(T, np.float32).copy()
        s0 = Stats(); s1 = Stats()
        self._ck(lib().llb_odom_optimize(self._h, _fp(t), ctypes.byref(s0), ctypes.byref(s1)))
        return t, s0, s1

    def odom_iterate(self, which: int, T, it: int):
        t = np.ascontiguousarray(T, np.float32).copy()
        more = ctypes.c_int(0); nc = ctypes.c_int(0)
        self._ck(lib().llb_odom_iterate(self._h, which, _fp(t), it, ctypes.byref(more), ctypes.byref(nc)))
        return t, bool(more.value), nc.value

    def odom_get_correspondences(self):
        n = ctypes.c_int(0)
        self._ck(lib().llb_odom_get_correspondences(self._h, None, None, 0, ctypes.byref(n)))
        ori = np.zeros((max(n.value, 1), 8), np.float32); co = np.zeros((max(n.value, 1), 8), np.float32)
        self._ck(lib().llb_odom_get_correspondences(self._h, _vp(ori), _vp(co), n.value, ctypes.byref(n)))
        return from_pcl(ori[:n.value]), from_pcl(co[:n.value])

    def odom_get_search_ind(self, which: int):
        n = ctypes.c_int(0)
        self._ck(lib().llb_odom_get_search_ind(self._h, which, None, None, None, 0, ctypes.byref(n)))
        a = np.zeros(max(n.value, 1), np.float32); b = a.copy(); c = a.copy()
        self._ck(lib().llb_odom_get_search_ind(self._h, which, _fp(a), _fp(b), _fp(c), n.value, ctypes.byref(n)))
        return a[:n.value], b[:n.value], c[:n.value]

    def odom_get_degeneracy(self):
        d = ctypes.c_int(0); P = np.zeros((3, 3), np.float32)
        self._ck(lib().llb_odom_get_degeneracy(self._h, ctypes.byref(d), _fp(P)))
        return bool(d.value), P

    # ---- device-resident family (pointers are raw device addresses, e.g. torch .data_ptr())
    def map_set_ds_dev(self, corner_ptr: int, mc: int, surf_ptr: int, ms: int):
        self._ck(lib().llb_map_set_ds_dev(self._h, ctypes.c_void_p(corner_ptr), mc, ctypes.c_void_p(surf_ptr), ms))

    def map_set_raw_dev(self, corner_ptr: int, rc: int, surf_ptr: int, rs: int):
        self._ck(lib().llb_map_set_raw_dev(self._h, ctypes.c_void_p(corner_ptr), rc, ctypes.c_void_p(surf_ptr), rs))

    def scan_set_dev(self, c_ptr: int, nc: int, s_ptr: int, ns: int, o_ptr: int, no: int):
        self._ck(lib().llb_scan_set_dev(self._h, ctypes.c_void_p(c_ptr), nc, ctypes.c_void_p(s_ptr), ns,
                                        ctypes.c_void_p(o_ptr), no))

    def s2m_optimize_dev(self, T_ptr: int):
        self._ck(lib().llb_s2m_optimize_dev(self._h, ctypes.c_void_p(T_ptr)))

    def s2m_time_iteration(self, T, reps: int = 20):
        """-> (ms per launch of the fused K3+K4 kernel, queries per launch)"""
        t = np.ascontiguousarray(T, np.float32)
        ms = ctypes.c_float(0); nq = ctypes.c_int(0)
        self._ck(lib().llb_s2m_time_iteration(self._h, _fp(t), reps, ctypes.byref(ms), ctypes.byref(nq)))
        return float(ms.value), int(nq.value)

    def s2m_get_profile(self, it: int = 0):
        """clock64 stamps of CTA 0 in iteration `it` of the last run -> cycle deltas per phase"""
        st = (ctypes.c_longlong * 8)()
        self._ck(lib().llb_s2m_get_profile(self._h, it, st))
        v = list(st)
        names = ["A_knn", "B_fit", "C_products", "sync1", "reduce", "lm_solve", "sync2"]
        return {n: v[i + 1] - v[i] for i, n in enumerate(names)}

    def s2m_get_cta_profile(self) -> np.ndarray:
        """(n_ctas, 4) cycles {A, B, C, wait} per CTA for the last iteration of the last run"""
        n = ctypes.c_int(0)
        self._ck(lib().llb_s2m_get_cta_profile(self._h, None, 0, ctypes.byref(n)))
        out = np.zeros((max(n.value, 1), 4), np.float64)
        self._ck(lib().llb_s2m_get_cta_profile(self._h, out.ctypes.data_as(ctypes.c_void_p), n.value, ctypes.byref(n)))
        return out[:n.value]

    def s2m_pose_set(self, T):
        t = np.ascontiguousarray(T, np.float32)
        self._ck(lib().llb_s2m_pose_set(self._h, _fp(t)))

    def s2m_pose_get(self) -> np.ndarray:
        t = np.zeros(6, np.float32)
        self._ck(lib().llb_s2m_pose_get(self._h, _fp(t)))
        return t

    def p2p_export(self) -> bytes:
        h = (ctypes.c_ubyte * 64)()
        self._ck(lib().llb_p2p_export(self._h, h))
        return bytes(h)

    def p2p_import(self, rank: int, world: int, handles):
        buf = b"".join(handles)
        assert len(buf) == 64 * world
        arr = (ctypes.c_ubyte * len(buf)).from_buffer_copy(buf)
        self._ck(lib().llb_p2p_import(self._h, rank, world, arr))

    def s2m_optimize_sharded(self, T):
        t = np.ascontiguousarray(T, np.float32).copy()
        st = Stats()
        self._ck(lib().llb_s2m_optimize_sharded(self._h, _fp(t), ctypes.byref(st)))
        return t, st

    # ---- loop closure + global map
    def loop_set_clouds(self, latest_id: int, latest_pose, hist_ids, hist_poses, history_leaf: float = 0.4):
        """cloud part of detectLoopClosure MO:838-861 on the device key-frame store -> (n latest, n history DS)"""
        lp = np.ascontiguousarray(latest_pose, np.float32)
        ids = np.ascontiguousarray(hist_ids, np.int32); hp = np.ascontiguousarray(hist_poses, np.float32).reshape(-1, 6)
        cnt = (ctypes.c_int * 2)()
        self._ck(lib().llb_loop_set_clouds(self._h, int(latest_id), _fp(lp), ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                           _fp(hp), ids.shape[0], ctypes.c_float(history_leaf), cnt))
        return cnt[0], cnt[1]

    def loop_set_clouds_host(self, latest, history_ds):
        a = to_pcl(latest); b = to_pcl(history_ds)
        self._hold("loop", (a, b))
        self._ck(lib().llb_loop_set_clouds_host(self._h, _vp(a), a.shape[0], _vp(b), b.shape[0]))

    def loop_icp(self, params: "LoopParams | None" = None) -> "IcpResult":
        out = IcpResult()
        self._ck(lib().llb_loop_icp(self._h, ctypes.byref(params) if params is not None else None, ctypes.byref(out)))
        return out

    def loop_get_cloud(self, which: int) -> np.ndarray:
        n = ctypes.c_int(0)
        self._ck(lib().llb_loop_get_cloud(self._h, which, None, 0, ctypes.byref(n)))
        out = np.zeros((max(n.value, 1), 8), np.float32)
        self._ck(lib().llb_loop_get_cloud(self._h, which, _vp(out), n.value, ctypes.byref(n)))
        return from_pcl(out[:n.value])

    def loop_get_nn(self):
        n = ctypes.c_int(0)
        self._ck(lib().llb_loop_get_nn(self._h, None, None, 0, ctypes.byref(n)))
        idx = np.zeros(max(n.value, 1), np.int32); d2 = np.zeros(max(n.value, 1), np.float32)
        self._ck(lib().llb_loop_get_nn(self._h, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _fp(d2), n.value, ctypes.byref(n)))
        return idx[:n.value], d2[:n.value]

    def global_map_assemble(self, ids, poses, leaf: float = 0.4) -> int:
        """cloud part of publishGlobalMap MO:780-788; the cloud: loop_get_cloud(3)"""
        i = np.ascontiguousarray(ids, np.int32); p = np.ascontiguousarray(poses, np.float32).reshape(-1, 6)
        n = ctypes.c_int(0)
        self._ck(lib().llb_global_map_assemble(self._h, i.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _fp(p), i.shape[0],
                                               ctypes.c_float(leaf), ctypes.byref(n)))
        return n.value

    # ---- sharded local map (BASELINE config 4, map sharded)
    def map_set_raw_sharded(self, corner, surf, rank: int, world: int):
        c = to_pcl(corner); s = to_pcl(surf)
        self._hold("map_raw", (c, s))
        self._ck(lib().llb_map_set_raw_sharded(self._h, _vp(c), c.shape[0], _vp(s), s.shape[0], rank, world))

    def map_set_raw_sharded_dev(self, corner_ptr: int, rc: int, surf_ptr: int, rs: int, rank: int, world: int):
        self._ck(lib().llb_map_set_raw_sharded_dev(self._h, ctypes.c_void_p(corner_ptr), rc, ctypes.c_void_p(surf_ptr), rs,
                                                   rank, world))

    def map_shard_info(self) -> "ShardInfo":
        out = ShardInfo()
        self._ck(lib().llb_map_shard_info(self._h, ctypes.byref(out)))
        return out

    def map_shard_set_global(self, n_corner_ds: int, n_surf_ds: int):
        g = (ctypes.c_int * 2)(int(n_corner_ds), int(n_surf_ds))
        self._ck(lib().llb_map_shard_set_global(self._h, g))

    def s2m_accumulate(self, it: int, rank: int, world: int) -> int:
        p = ctypes.c_void_p()
        self._ck(lib().llb_s2m_accumulate(self._h, it, rank, world, ctypes.byref(p)))
        return p.value

    def s2m_solve(self, it: int, want_converged: bool = False):
        conv = ctypes.c_int(0)
        self._ck(lib().llb_s2m_solve(self._h, it, ctypes.byref(conv) if want_converged else None))
        return bool(conv.value)


class LoopParams(ctypes.Structure):
    """llb_loop_params (defaults = MO:893-896)"""
    _fields_ = [("max_iterations", ctypes.c_int), ("max_correspondence_distance", ctypes.c_double),
                ("transformation_epsilon", ctypes.c_double), ("euclidean_fitness_epsilon", ctypes.c_double)]


class IcpResult(ctypes.Structure):
    """llb_icp_result"""
    _fields_ = [("T", ctypes.c_float * 16), ("has_converged", ctypes.c_int), ("iterations", ctypes.c_int),
                ("convergence_state", ctypes.c_int), ("n_correspondences", ctypes.c_int), ("fitness_score", ctypes.c_double),
                ("sums", ctypes.c_double * 17), ("device_ms", ctypes.c_float), ("n_source", ctypes.c_int), ("n_target", ctypes.c_int)]


class ShardInfo(ctypes.Structure):
    """llb_shard_info"""
    _fields_ = [("axis", ctypes.c_int), ("lo", ctypes.c_float), ("hi", ctypes.c_float), ("rank", ctypes.c_int),
                ("world", ctypes.c_int), ("raw_kept", ctypes.c_int * 2), ("ds_local", ctypes.c_int * 2),
                ("ds_owned", ctypes.c_int * 2)]


class Batch:
    """One llb_batch: n_slots independent sequences registered per step with a slot-count-independent number of
    launches (BASELINE config 5).  Mirrors Context's scan/map/optimise calls with a leading slot index."""

    def __init__(self, device: int, n_slots: int, max_scan_points: int, max_map_points: int, params: Params | None = None):
        L = lib()
        self._h = ctypes.c_void_p()
        if params is None:
            params = default_params()
        self.params = params
        rc = L.llb_batch_create(ctypes.byref(params), int(device), int(n_slots), int(max_scan_points),
                                int(max_map_points), ctypes.byref(self._h))
        if rc != LLB_OK:
            self._h = ctypes.c_void_p()
            raise LlbError(rc, "llb_batch_create failed (no CPU fallback exists)")
        self.n_slots = n_slots
        self._keep = {}                   # host clouds stay alive (and unchanged) until the step has finished

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().llb_batch_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != LLB_OK:
            raise LlbError(rc, (lib().llb_batch_last_error(self._h) or b"").decode())

    @property
    def stream(self) -> int:
        return lib().llb_batch_stream(self._h) or 0

    def launch_count(self) -> int:
        return int(lib().llb_batch_launch_count(self._h))

    def scan_set(self, slot: int, corner_last, surf_last, outlier_last):
        self.scan_set_pcl(slot, to_pcl(corner_last), to_pcl(surf_last), to_pcl(outlier_last))

    def scan_set_pcl(self, slot: int, c32, s32, o32):
        self._keep[("scan", slot)] = (c32, s32, o32)
        self._ck(lib().llb_batch_scan_set(self._h, slot, _vp(c32), c32.shape[0], _vp(s32), s32.shape[0],
                                          _vp(o32), o32.shape[0]))

    def map_set_ds(self, slot: int, corner_ds, surf_ds):
        self.map_set_ds_pcl(slot, to_pcl(corner_ds), to_pcl(surf_ds))

    def map_set_ds_pcl(self, slot: int, c32, s32):
        self._keep[("map", slot)] = (c32, s32)
        self._ck(lib().llb_batch_map_set_ds(self._h, slot, _vp(c32), c32.shape[0], _vp(s32), s32.shape[0]))

    def scan_set_dev(self, slot: int, c_ptr: int, nc: int, s_ptr: int, ns: int, o_ptr: int, no: int):
        self._ck(lib().llb_batch_scan_set_dev(self._h, slot, ctypes.c_void_p(c_ptr), nc, ctypes.c_void_p(s_ptr), ns,
                                              ctypes.c_void_p(o_ptr), no))

    def map_set_ds_dev(self, slot: int, c_ptr: int, mc: int, s_ptr: int, ms: int):
        self._ck(lib().llb_batch_map_set_ds_dev(self._h, slot, ctypes.c_void_p(c_ptr), mc, ctypes.c_void_p(s_ptr), ms))

    @staticmethod
    def pack(ptrs, counts):
        """-> (uint64 pointer array, int32 count array) for the *_all calls (build once, reuse every step)"""
        return np.ascontiguousarray(ptrs, np.uint64), np.ascontiguousarray(counts, np.int32)

    @staticmethod
    def _pa(a):
        return a.ctypes.data_as(ctypes.c_void_p)

    def scan_set_all(self, c, s, o, dev: bool):
        """c, s, o: (pointer array, count array) pairs from pack(); dev selects device float4 / host PCL clouds"""
        fn = lib().llb_batch_scan_set_dev_all if dev else lib().llb_batch_scan_set_all
        self._ck(fn(self._h, self._pa(c[0]), self._pa(c[1]), self._pa(s[0]), self._pa(s[1]), self._pa(o[0]), self._pa(o[1])))

    def map_set_ds_all(self, c, s, dev: bool):
        fn = lib().llb_batch_map_set_ds_dev_all if dev else lib().llb_batch_map_set_ds_all
        self._ck(fn(self._h, self._pa(c[0]), self._pa(c[1]), self._pa(s[0]), self._pa(s[1])))

    def register(self, T):
        """T: (n_slots, 6) initial transformTobeMapped -> (poses (n_slots, 6), [Stats] * n_slots)"""
        t = np.ascontiguousarray(T, np.float32).reshape(self.n_slots, 6).copy()
        st = (Stats * self.n_slots)()
        self._ck(lib().llb_batch_register(self._h, _fp(t), st))
        return t, list(st)

    def register_async(self, T):
        t = np.ascontiguousarray(T, np.float32).reshape(self.n_slots, 6)
        self._async_T = t.copy()
        self._ck(lib().llb_batch_register_async(self._h, _fp(t)))

    def result(self):
        t = self._async_T.copy()
        st = (Stats * self.n_slots)()
        self._ck(lib().llb_batch_result(self._h, _fp(t), st))
        return t, list(st)

    def scan_get_ds(self, slot: int, which: int) -> np.ndarray:
        n = ctypes.c_int(0)
        self._ck(lib().llb_batch_scan_get_ds(self._h, slot, which, None, 0, ctypes.byref(n)))
        out = np.zeros((max(n.value, 1), 8), np.float32)
        self._ck(lib().llb_batch_scan_get_ds(self._h, slot, which, _vp(out), n.value, ctypes.byref(n)))
        return from_pcl(out[:n.value])

    # ---- per-slot device-resident key-frame stores
    def enable_keyframes(self, max_raw_map_points: int, max_keyframes: int):
        self._ck(lib().llb_batch_enable_keyframes(self._h, int(max_raw_map_points), int(max_keyframes)))

    def keyframe_add(self, slot: int) -> int:
        k = ctypes.c_int(-1)
        self._ck(lib().llb_batch_keyframe_add(self._h, slot, ctypes.byref(k)))
        return k.value

    def keyframe_count(self, slot: int) -> int:
        n = ctypes.c_int(0)
        self._ck(lib().llb_batch_keyframe_count(self._h, slot, ctypes.byref(n)))
        return n.value

    def map_assemble(self, slot: int, ids, poses6d):
        i = np.ascontiguousarray(ids, np.int32)
        p = np.ascontiguousarray(poses6d, np.float32).reshape(-1, 6)
        assert p.shape[0] == i.shape[0]
        self._ck(lib().llb_batch_map_assemble(self._h, slot, i.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _fp(p),
                                              int(i.shape[0])))

    @staticmethod
    def pack_assemble(ids_per_slot, poses_per_slot):
        """-> (ids, poses, offset) arrays for map_assemble_all (build once while the surrounding key-frames do not change)"""
        off = np.zeros(len(ids_per_slot) + 1, np.int32)
        off[1:] = np.cumsum([len(i) for i in ids_per_slot])
        ids = np.ascontiguousarray(np.concatenate([np.asarray(i, np.int32).reshape(-1) for i in ids_per_slot]), np.int32)
        poses = np.ascontiguousarray(np.concatenate([np.asarray(p, np.float32).reshape(-1, 6) for p in poses_per_slot]), np.float32)
        assert poses.shape[0] == ids.shape[0]
        return ids, poses, off

    def map_assemble_all(self, packed):
        ids, poses, off = packed
        assert off.shape[0] == self.n_slots + 1
        self._ck(lib().llb_batch_map_assemble_all(self._h, ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _fp(poses),
                                                  off.ctypes.data_as(ctypes.POINTER(ctypes.c_int))))

    def map_get(self, slot: int, which: int) -> np.ndarray:
        n = ctypes.c_int(0)
        self._ck(lib().llb_batch_map_get(self._h, slot, which, None, 0, ctypes.byref(n)))
        out = np.zeros((max(n.value, 1), 8), np.float32)
        self._ck(lib().llb_batch_map_get(self._h, slot, which, _vp(out), n.value, ctypes.byref(n)))
        return from_pcl(out[:n.value])

    # ---- featureAssociation of the slots
    def odom_set(self, slot: int, corner_last, surf_last, corner_sharp, surf_flat):
        arrs = tuple(to_pcl(x) for x in (corner_last, surf_last, corner_sharp, surf_flat))
        self._keep[("odom", slot)] = arrs
        args = []
        for a in arrs:
            args += [_vp(a), a.shape[0]]
        self._ck(lib().llb_batch_odom_set(self._h, slot, *args))

    # ---- feature extraction of the slots
    def features_init(self, n_scan: int, horizon_scan: int):
        self._ck(lib().llb_batch_features_init(self._h, int(n_scan), int(horizon_scan)))

    def features_pack(self, sweeps):
        """Builds the llb_segmented_cloud array for one sweep per slot (reusable across steps)."""
        arr = (SegmentedCloud * len(sweeps))()
        keep = [segmented_struct(sw, arr[i])[1] for i, sw in enumerate(sweeps)]
        return arr, keep

    def features_extract(self, packed):
        arr, keep = packed
        counts = np.zeros((len(arr), 4), np.int32); ms = ctypes.c_float(0)
        self._ck(lib().llb_batch_features_extract(self._h, arr, _vp(counts), ctypes.byref(ms)))
        return counts, float(ms.value)

    def features_get(self, slot: int, which: int):
        n = ctypes.c_int(0)
        self._ck(lib().llb_batch_features_get(self._h, int(slot), int(which), None, 0, ctypes.byref(n)))
        out = np.zeros((max(n.value, 1), 8), np.float32)
        self._ck(lib().llb_batch_features_get(self._h, int(slot), int(which), _vp(out), out.shape[0], ctypes.byref(n)))
        return from_pcl(out[:n.value])

    def odom_optimize(self, T):
        """T: (n_slots, 6) transformCur -> (poses, [surf Stats], [corner Stats])"""
        t = np.ascontiguousarray(T, np.float32).reshape(self.n_slots, 6).copy()
        s0 = (Stats * self.n_slots)(); s1 = (Stats * self.n_slots)()
        self._ck(lib().llb_batch_odom_optimize(self._h, _fp(t), s0, s1))
        return t, list(s0), list(s1)

    def get_degeneracy(self, slot: int) -> bool:
        d = ctypes.c_int(0)
        self._ck(lib().llb_batch_get_degeneracy(self._h, slot, ctypes.byref(d)))
        return bool(d.value)

    def set_profile(self, on: bool):
        self._ck(lib().llb_batch_set_profile(self._h, int(on)))

    def get_profile(self):
        ms = (ctypes.c_float * 6)(); geo = (ctypes.c_int * 4)()
        self._ck(lib().llb_batch_get_profile(self._h, ms, geo))
        names = ["unpack", "downsample", "index_build", "knn", "fit", "lm_step"]
        return dict(zip(names, [float(x) for x in ms])), {"knn_ctas_per_slot": geo[0], "fit_ctas_per_slot": geo[1],
                                                          "index_ctas_per_map": geo[2], "query_capacity": geo[3]}


def shard_plan(sample_xyz, rank: int, world: int):
    """llb_shard_plan: (axis, lo, hi) of the slab of `rank` from a sample of map points (pure host function)"""
    a = np.ascontiguousarray(sample_xyz, np.float32).reshape(-1, 3)
    ax = ctypes.c_int(0); lo = ctypes.c_float(0); hi = ctypes.c_float(0)
    rc = lib().llb_shard_plan(_fp(a), a.shape[0], rank, world, ctypes.byref(ax), ctypes.byref(lo), ctypes.byref(hi))
    if rc != 0:
        raise LlbError(rc, "llb_shard_plan")
    return ax.value, lo.value, hi.value


def default_loop_params() -> "LoopParams":
    p = LoopParams()
    lib().llb_loop_params_default(ctypes.byref(p))
    return p


def default_params() -> Params:
    p = Params()
    lib().llb_params_default(ctypes.byref(p))
    return p
