"""CPU tests of the oracle's pcl::VoxelGrid / pcl::KdTreeFLANN restatements.
Neither library exists offline, so these check the oracle against independent implementations of the
published algorithm (numpy) and against exact-kNN implementations that do exist here (scipy, cv2.flann)."""
import numpy as np
import pytest

import oracle
from tests import data


def voxel_numpy(pts: np.ndarray, leaf: float) -> np.ndarray:
    """Independent numpy statement of A.1 (stable order, sequential float32 sums)."""
    inv = np.float32(1.0) / np.float32(leaf)
    xyz = pts[:, :3]
    mn = xyz.min(0); mx = xyz.max(0)
    min_b = np.floor(mn * inv).astype(np.int32); max_b = np.floor(mx * inv).astype(np.int32)
    div = max_b - min_b + 1
    ijk = (np.floor(xyz * inv) - min_b.astype(np.float32)).astype(np.int32)
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(idx, kind="stable")
    out = []
    i = 0
    while i < len(order):
        j = i
        s = np.zeros(4, np.float32)
        while j < len(order) and idx[order[j]] == idx[order[i]]:
            s = (s + pts[order[j]]).astype(np.float32)
            j += 1
        out.append(s / np.float32(j - i))
        i = j
    return np.array(out, np.float32)


@pytest.mark.parametrize("n,leaf,seed", [(1, 0.2, 0), (7, 0.4, 1), (500, 0.2, 2), (4000, 0.4, 3), (4000, 1.0, 4)])
def test_voxel_matches_independent_statement(n, leaf, seed):
    pts = data.random_cloud(n, seed)
    out, ovf = oracle.voxel_grid(pts, leaf)
    ref = voxel_numpy(pts, leaf)
    assert ovf == 0 and out.shape == ref.shape
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


def test_voxel_leaf_inverse_is_exact_for_reference_leaves():
    # A.1 step 1: 1/0.2f == 5.0f, 1/0.4f == 2.5f, 1/1.0f == 1.0f exactly in float32
    assert np.float32(1) / np.float32(0.2) == np.float32(5.0)
    assert np.float32(1) / np.float32(0.4) == np.float32(2.5)


def test_voxel_empty_single_and_intensity_average():
    out, _ = oracle.voxel_grid(np.zeros((0, 4), np.float32), 0.2)
    assert out.shape == (0, 4)
    p = np.array([[0.01, 0.02, 0.03, 5.0], [0.05, 0.06, 0.07, 6.0]], np.float32)
    out, _ = oracle.voxel_grid(p, 0.2)
    assert out.shape == (1, 4)
    assert out[0, 3] == np.float32(5.5)            # intensity is averaged with xyz (Survey section 0)


def test_voxel_overflow_passthrough():
    pts = data.random_cloud(300, 5, extent=(900.0, 300.0, 900.0), clustered=False)
    out, ovf = oracle.voxel_grid(pts, 0.2)
    assert ovf == 1 and np.array_equal(out, pts)   # C18: PCL copies the input


def test_voxel_properties():
    pts = data.random_cloud(20000, 6)
    out, _ = oracle.voxel_grid(pts, 0.4)
    inv = np.float32(2.5)
    kin = np.floor(pts[:, :3] * inv).astype(np.int64)
    assert out.shape[0] == np.unique(kin, axis=0).shape[0]
    # every centroid lies in the bounding box of the cloud
    assert np.all(out[:, :3] >= pts[:, :3].min(0) - 1e-6) and np.all(out[:, :3] <= pts[:, :3].max(0) + 1e-6)


@pytest.mark.parametrize("k", [1, 5])
def test_kdtree_equals_bruteforce_and_scipy(k):
    scipy_spatial = pytest.importorskip("scipy.spatial")
    pts = data.random_cloud(6000, 7, extent=(20, 3, 20))
    q = data.random_cloud(400, 8, extent=(21, 3, 21))
    tree = oracle.KdTree(pts)
    ti, td = tree.knn(q, k)
    bi, bd = oracle.knn_bruteforce(pts, q, k)
    assert np.array_equal(ti, bi) and np.array_equal(td.view(np.uint32), bd.view(np.uint32))
    d, si = scipy_spatial.cKDTree(pts[:, :3].astype(np.float64)).query(q[:, :3].astype(np.float64), k=k)
    si = si.reshape(len(q), k)
    # equal-distance ties are the documented non-determinism (A.2): compare as sets where distances are distinct
    distinct = np.all(np.diff(td.astype(np.float64), axis=1) > 0, axis=1) if k > 1 else np.ones(len(q), bool)
    assert np.array_equal(np.sort(ti[distinct], 1), np.sort(si[distinct], 1))


def test_kdtree_equals_flann_exact():
    cv2 = pytest.importorskip("cv2")
    pts = data.random_cloud(5000, 9, extent=(15, 3, 15))
    q = data.random_cloud(300, 10, extent=(15, 3, 15))
    # FLANN KDTreeSingleIndex (algorithm 4), leaf 15, reorder, exact search: what pcl::KdTreeFLANN builds (A.2)
    index = cv2.flann_Index(pts[:, :3].copy(), {"algorithm": 4, "leaf_max_size": 15, "reorder": True})
    fi, fd = index.knnSearch(q[:, :3].copy(), 5, params={"checks": -1, "eps": 0.0, "sorted": True})
    ti, td = oracle.KdTree(pts).knn(q, 5)
    distinct = np.all(np.diff(td.astype(np.float64), axis=1) > 0, axis=1)
    assert np.array_equal(ti[distinct], fi[distinct])
    assert np.allclose(td, fd, rtol=1e-6, atol=1e-9)     # squared L2 in float, same summation order


def test_kdtree_tie_rule_smaller_index_first():
    pts = np.zeros((40, 4), np.float32)
    pts[:, 0] = np.repeat(np.arange(8), 5)               # 5 coincident points at each of 8 x positions
    ti, td = oracle.KdTree(pts).knn(np.array([[0.1, 0, 0, 0]], np.float32), 5)
    assert ti.tolist() == [[0, 1, 2, 3, 4]]
