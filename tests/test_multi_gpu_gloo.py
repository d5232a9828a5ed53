"""World-size-2 gloo test (CPU) of the N>1 host logic: round-robin query sharding + all-reduce(sum) of the
28-value accumulator reproduces the single-process normal equations of the oracle after one rounding."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def jacobian_rows(T, ori, co):
    """MO:1252-1271 in float32 numpy with the reference's association order."""
    f = np.float32
    srx, crx = f(np.sin(T[0], dtype=np.float32)), f(np.cos(T[0], dtype=np.float32))
    sry, cry = f(np.sin(T[1], dtype=np.float32)), f(np.cos(T[1], dtype=np.float32))
    srz, crz = f(np.sin(T[2], dtype=np.float32)), f(np.cos(T[2], dtype=np.float32))
    x, y, z = ori[:, 0], ori[:, 1], ori[:, 2]
    cx, cy, cz = co[:, 0], co[:, 1], co[:, 2]
    arx = (crx * sry * srz * x + crx * crz * sry * y - srx * sry * z) * cx \
        + (-srx * srz * x - crz * srx * y - crx * z) * cy \
        + (crx * cry * srz * x + crx * cry * crz * y - cry * srx * z) * cz
    ary = ((cry * srx * srz - crz * sry) * x + (sry * srz + cry * crz * srx) * y + crx * cry * z) * cx \
        + ((-cry * crz - srx * sry * srz) * x + (cry * srz - crz * srx * sry) * y - crx * sry * z) * cz
    arz = ((crz * srx * sry - cry * srz) * x + (-cry * crz - srx * sry * srz) * y) * cx \
        + (crx * crz * x - crx * srz * y) * cy \
        + ((sry * srz + cry * crz * srx) * x + (crz * sry - cry * srx * srz) * y) * cz
    return np.stack([arx, ary, arz, cx, cy, cz, -co[:, 3]], 1).astype(np.float32)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from lego_loam_b200 import multi_gpu
    from tests import data
    case = data.mapping_case(1, 8000, 50000)
    oracle.set_trig_mode(0)
    mo = oracle.MapOptimization()
    mo.set_map_raw(case["map_corner_raw"], case["map_surf_raw"])
    mo.set_scan(case["corner"], case["surf"], case["outlier"])
    mo.downsampleCurrentScan(); mo.build_kdtrees()
    mo.transformTobeMapped = case["init"]
    mo.clear_correspondences(); mo.cornerOptimization(0); mo.surfOptimization(0)
    ori, co = mo.correspondences()
    rows = jacobian_rows(case["init"], ori, co)
    mine = multi_gpu.shard_queries(rows.shape[0], rank, world)
    acc = torch.from_numpy(multi_gpu.partial_sums(rows[mine]))
    multi_gpu.reduce_normal_equations(acc)
    A, B, n = multi_gpu.normal_equations_from_sums(acc.numpy())
    mo.LMOptimization(0)
    A_ref, B_ref, _ = mo.normal_eq()
    q.put((rank, n == rows.shape[0], bool(np.array_equal(A, A_ref)), bool(np.array_equal(B, B_ref)),
           acc.numpy().tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_normal_equations_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] and r[3] for r in res), res
    assert res[0][4] == res[1][4]            # every rank holds identical sums -> identical redundant LM step


def test_shard_partition_is_exact_cover():
    from lego_loam_b200 import multi_gpu
    for n in (0, 1, 7, 1000):
        for world in (1, 2, 3, 8):
            allq = np.concatenate([multi_gpu.shard_queries(n, r, world) for r in range(world)])
            assert np.array_equal(np.sort(allq), np.arange(n))


def _worker_slab(rank, world, port, q):
    """host logic of the MAP-sharded form: every rank plans its slab from the same sample (llb_shard_plan: pure host
    code of the C ABI), the slabs tile the axis, and the sizes of the unsharded maps come out of an all-reduce of the
    owned counts (multi_gpu.set_sharded_map with a stand-in context: the device work needs a GPU)."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lego_loam_b200 import api, multi_gpu
    rng = np.random.default_rng(42)                                  # the same "raw map" on every rank
    pts = (rng.normal(0, 1, (50000, 3)) * [40.0, 2.0, 15.0]).astype(np.float32)
    sample = pts[::7]
    axis, lo, hi = api.shard_plan(sample, rank, world)

    class StandIn:                                                   # what a context reports after filtering its slab
        device = 0
        def map_set_raw_sharded(self, c, s, r, w): self.args = (r, w)
        def map_shard_info(self):
            info = api.ShardInfo(); info.axis = axis; info.lo = lo; info.hi = hi; info.rank = rank; info.world = world
            own = int(np.sum((pts[:, axis] >= lo) & (pts[:, axis] < hi)))
            info.ds_owned[0] = own; info.ds_owned[1] = 2 * own
            return info
        def map_shard_set_global(self, a, b): self.glob = (a, b)
    ctx = StandIn()
    info = multi_gpu.set_sharded_map(ctx, None, None, rank, world)
    q.put((rank, axis, lo, hi, ctx.glob, ctx.args, int(info.ds_owned[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_map_sharded_host_logic_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_slab, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, ax0, lo0, hi0, g0, a0, o0), (r1, ax1, lo1, hi1, g1, a1, o1) = res
    assert ax0 == ax1 == 0                                           # the longest extent of the sample
    assert lo0 < -1e30 and hi1 > 1e30 and hi0 == lo1                 # the slabs tile the axis, open at both ends
    assert g0 == g1 == (50000, 100000) and o0 + o1 == 50000          # all-reduced owned counts = the unsharded sizes
    assert abs(o0 - o1) < 2500                                       # quantile borders: balanced slabs
    assert a0 == (0, 2) and a1 == (1, 2)


def test_shard_plan_tiles_the_axis():
    from lego_loam_b200 import api
    rng = np.random.default_rng(1)
    sample = (rng.uniform(-1, 1, (4096, 3)) * [10.0, 3.0, 80.0]).astype(np.float32)
    for world in (1, 2, 3, 4, 8):
        plans = [api.shard_plan(sample, r, world) for r in range(world)]
        assert all(p[0] == 2 for p in plans)
        assert plans[0][1] < -1e30 and plans[-1][2] > 1e30
        for a, b in zip(plans[:-1], plans[1:]):
            assert a[2] == b[1] and a[1] < a[2]
    with __import__("pytest").raises(api.LlbError):
        api.shard_plan(sample, 2, 2)
