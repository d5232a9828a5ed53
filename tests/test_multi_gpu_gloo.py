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
