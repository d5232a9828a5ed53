"""Pin the oracle (own restatement, oracle/llo_*.c) against the REFERENCE ITSELF: the unmodified
mapOptmization.cpp / featureAssociation.cpp compiled against shim headers (oracle/_ref, built by
oracle/ref_harness/Makefile where /root/reference exists).  Control flow, thresholds, Jacobians,
mixed-precision sub-expressions and pose bookkeeping come from the reference verbatim there, so
bit-equality here is what makes the oracle a faithful stand-in.  Also checked against the golden
vectors generated from that harness (tests/golden/ref_*.npz), which travel to machines without the
reference mount."""
import os

import numpy as np
import pytest

import oracle
from oracle import ref_harness
from tests import data

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


needs_ref = pytest.mark.skipif(not ref_harness.available(), reason="oracle/_ref not built (needs /root/reference)")


def _mapping_pair(seed):
    case = data.mapping_case(seed, 15000, 90000)
    oracle.set_trig_mode(0)
    a = oracle.MapOptimization(); b = ref_harness.MapOptimization()
    for m in (a, b):
        m.set_map_raw(case["map_corner_raw"], case["map_surf_raw"])
        m.set_scan(case["corner"], case["surf"], case["outlier"])
        m.downsampleCurrentScan()
        m.transformTobeMapped = case["init"]
    return case, a, b


@needs_ref
@pytest.mark.parametrize("seed", [1, 2])
def test_mapping_functions_bitexact_vs_reference(seed):
    case, a, b = _mapping_pair(seed)
    for which in range(4):
        assert np.array_equal(_bits(a.scan_ds(which)), _bits(b.scan_ds(which)))
    for which in range(2):
        assert np.array_equal(_bits(a.map_ds(which)), _bits(b.map_ds(which)))
    a.build_kdtrees(); b.build_kdtrees()
    for it in range(3):
        for m in (a, b):
            m.clear_correspondences(); m.cornerOptimization(it); m.surfOptimization(it)
        (oa, ca), (ob, cb) = a.correspondences(), b.correspondences()
        assert oa.shape[0] == ob.shape[0] > 50
        assert np.array_equal(_bits(oa), _bits(ob)) and np.array_equal(_bits(ca), _bits(cb))
        ra, rb = a.LMOptimization(it), b.LMOptimization(it)
        assert ra == rb
        assert np.array_equal(_bits(a.transformTobeMapped), _bits(b.transformTobeMapped))
        assert a.degenerate()[0] == b.degenerate()[0]
        assert np.array_equal(_bits(a.degenerate()[1]), _bits(b.degenerate()[1]))


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_scan2map_bitexact_vs_reference(seed):
    case, a, b = _mapping_pair(seed)
    a.set_transform_sum(case["init"]); b.set_transform_sum(case["init"])
    a.scan2MapOptimization(); b.scan2MapOptimization()
    assert np.array_equal(_bits(a.transformTobeMapped), _bits(b.transformTobeMapped))
    (ba, aa), (bb, ab) = a.bef_aft(), b.bef_aft()                 # transformUpdate MO:463-496
    assert np.array_equal(_bits(ba), _bits(bb)) and np.array_equal(_bits(aa), _bits(ab))


def _odom_pair(seed):
    from tests.test_gpu_parity import _odom_case
    od = _odom_case(seed)
    oracle.set_trig_mode(0)
    a = oracle.FeatureAssociation(); b = ref_harness.FeatureAssociation()
    for f in (a, b):
        f.set_last(od.corner_last, od.surf_last, force=True)
        f.set_features(od.corner_sharp, od.surf_flat)
        f.transformCur = np.zeros(6, np.float32)
    return od, a, b


@needs_ref
@pytest.mark.parametrize("seed", [1, 2])
def test_odometry_functions_bitexact_vs_reference(seed):
    od, a, b = _odom_pair(seed)
    for it in (0, 1, 2, 5, 6):
        for f in (a, b):
            f.clear_correspondences(); f.findCorrespondingSurfFeatures(it)
        (oa, ca), (ob, cb) = a.correspondences(), b.correspondences()
        assert np.array_equal(_bits(oa), _bits(ob)) and np.array_equal(_bits(ca), _bits(cb))
        for x, y in zip(a.search_ind(1), b.search_ind(1)):
            assert np.array_equal(x, y)
        assert a.calculateTransformationSurf(it) == b.calculateTransformationSurf(it)
        assert np.array_equal(_bits(a.transformCur), _bits(b.transformCur))
    for it in (0, 1, 5):
        for f in (a, b):
            f.clear_correspondences(); f.findCorrespondingCornerFeatures(it)
        (oa, ca), (ob, cb) = a.correspondences(), b.correspondences()
        assert np.array_equal(_bits(oa), _bits(ob)) and np.array_equal(_bits(ca), _bits(cb))
        for x, y in zip(a.search_ind(0)[:2], b.search_ind(0)[:2]):
            assert np.array_equal(x, y)
        assert a.calculateTransformationCorner(it) == b.calculateTransformationCorner(it)
        assert np.array_equal(_bits(a.transformCur), _bits(b.transformCur))
        assert a.degenerate()[0] == b.degenerate()[0]


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_update_transformation_bitexact_vs_reference(seed):
    od, a, b = _odom_pair(seed)
    a.updateTransformation(); b.updateTransformation()
    assert np.array_equal(_bits(a.transformCur), _bits(b.transformCur))


def test_golden_vectors_from_reference_harness():
    """Outputs of the reference harness committed as fixtures (tests/golden/make_ref_golden.py)."""
    g = np.load(os.path.join(GOLD, "ref_scan2map_golden.npz"))
    oracle.set_trig_mode(0)
    for k in range(int(g["n_cases"])):
        mo = oracle.MapOptimization()
        mo.set_map_ds(g[f"map_corner_ds_{k}"], g[f"map_surf_ds_{k}"])
        mo.set_scan(g[f"corner_{k}"], g[f"surf_{k}"], g[f"outlier_{k}"])
        mo.downsampleCurrentScan()
        assert np.array_equal(_bits(mo.scan_ds(0)), _bits(g[f"corner_ds_{k}"]))
        assert np.array_equal(_bits(mo.scan_ds(3)), _bits(g[f"surf_total_ds_{k}"]))
        mo.transformTobeMapped = g[f"init_{k}"]
        mo.scan2MapOptimization()
        # libm flavour (sinf/cosf) can differ between the machine that made the fixture and this one
        assert np.allclose(mo.transformTobeMapped, g[f"pose_{k}"], atol=2e-6)
    g = np.load(os.path.join(GOLD, "ref_odometry_golden.npz"))
    for k in range(int(g["n_cases"])):
        fa = oracle.FeatureAssociation()
        fa.set_last(g[f"corner_last_{k}"], g[f"surf_last_{k}"], force=True)
        fa.set_features(g[f"sharp_{k}"], g[f"flat_{k}"])
        fa.transformCur = np.zeros(6, np.float32)
        fa.updateTransformation()
        assert np.allclose(fa.transformCur, g[f"cur_{k}"], atol=2e-6)
