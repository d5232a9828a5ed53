"""BASELINE config 4 on real devices: one registration with its queries sharded over 2 GPUs, the 28-value exchange fused
into the persistent kernel (P2P stores into cudaIpc-mapped mailboxes) and, for comparison, as an NCCL all-reduce.
Skipped unless the box has at least 2 GPUs (the driver's single-GPU tier skips it; `gpurun --gpus 2` runs it)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_sharded_registration_two_gpus():
    env = dict(os.environ, LLB_SENSOR="hdl32e", LLB_RAW_CORNER="300000", LLB_RAW_SURF="200000")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "sharded_check.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["world"] == 2 and r["pose_match"] and r["bit_identical_across_ranks"] and r["fused_matches_nccl_bitwise"]
    assert r["iterations_sharded"] == r["iterations_single"]
    assert r["map_sharded_pose_equals_single_bitwise"] and max(r["map_sharded_raw_points_per_rank"]) < 0.8 * r["raw_points"]


def _rows_index(G):
    import numpy as np
    return {G[i].tobytes(): i for i in range(G.shape[0])}


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_map_virtual_ranks(world):
    """Config 4 with the MAP sharded, checked on ONE device: `world` contexts play the ranks.  Each keeps a slab of the raw
    map; its DS map must be a sub-sequence of the unsharded DS map (same bits, same order), the owned centroids must
    partition it, and the registration - every rank accumulating the queries inside its slab, the 28 sums added in rank
    order - must give the unsharded pose bit for bit."""
    import numpy as np
    import torch
    from lego_loam_b200 import api, multi_gpu
    from tests import data
    c = data.mapping_case(seed=4, n_corner_raw=60000, n_surf_raw=300000, sensor="hdl32e")
    ref = api.Context(0)
    ref.map_set_raw(c["map_corner_raw"], c["map_surf_raw"])
    ref.scan_set(c["corner"], c["surf"], c["outlier"]); ref.downsample_current_scan()
    T_single, st = ref.s2m_optimize(c["init"])
    G = [ref.map_get_ds(0), ref.map_get_ds(1)]
    gi = [_rows_index(G[0]), _rows_index(G[1])]
    ranks, infos, owned_idx = [], [], [set(), set()]
    for r in range(world):
        ctx = api.Context(0)
        ctx.map_set_raw_sharded(c["map_corner_raw"], c["map_surf_raw"], r, world)
        info = ctx.map_shard_info()
        assert info.rank == r and info.world == world and 0 <= info.axis <= 2 and info.lo < info.hi
        for k in range(2):
            L = ctx.map_get_ds(k)
            assert L.shape[0] == info.ds_local[k] and 0 < info.raw_kept[k]
            idx = np.array([gi[k].get(L[i].tobytes(), -1) for i in range(L.shape[0])])
            assert (idx >= 0).all(), f"rank {r}: {int((idx < 0).sum())} centroids of map {k} are not centroids of the unsharded map"
            assert (np.diff(idx) > 0).all()                   # the order of the unsharded output
            own = idx[(L[:, info.axis] >= info.lo) & (L[:, info.axis] < info.hi)]
            assert own.shape[0] == info.ds_owned[k]
            assert not (owned_idx[k] & set(own.tolist()))
            owned_idx[k] |= set(own.tolist())
        ranks.append(ctx); infos.append(info)
    for k in range(2):
        assert len(owned_idx[k]) == G[k].shape[0]             # the slabs partition the unsharded map
    if world > 1:
        assert max(i.raw_kept[1] for i in infos) < 0.8 * c["map_surf_raw"].shape[0]    # the map-side work really divides
    for ctx in ranks:
        ctx.map_shard_set_global(G[0].shape[0], G[1].shape[0])
        ctx.scan_set(c["corner"], c["surf"], c["outlier"]); ctx.downsample_current_scan()
        ctx.s2m_pose_set(c["init"])
    iters = 0
    dev = torch.device("cuda", 0)
    for it in range(10):
        accs = []
        for ctx in ranks:
            ptr = ctx.s2m_accumulate(it, 0, 1); ctx.synchronize()
            accs.append(torch.as_tensor(multi_gpu._DevView(ptr, multi_gpu.N_ACC), device=dev))
        total = accs[0].clone()
        for a in accs[1:]:
            total += a                                        # rank order, as the fused exchange adds them
        assert int(total[27].item()) > 0
        for a in accs:
            a.copy_(total)
        torch.cuda.synchronize()
        iters += 1
        conv = [ctx.s2m_solve(it, want_converged=True) for ctx in ranks]
        assert len(set(conv)) == 1
        if conv[0]:
            break
    T = [ctx.s2m_pose_get() for ctx in ranks]
    for t in T[1:]:
        assert np.array_equal(t.view(np.uint32), T[0].view(np.uint32))
    assert iters == st.iterations
    assert np.array_equal(T[0].view(np.uint32), np.asarray(T_single, np.float32).view(np.uint32)), (T[0], T_single)
    for ctx in ranks:
        ctx.close()
    ref.close()


def test_sharded_map_edge_cases():
    """world = 1: the slab is everything, the DS maps are the unsharded ones bit for bit; argument and call-order errors."""
    import numpy as np
    from lego_loam_b200 import api
    from tests import data
    c = data.mapping_case(seed=2, n_corner_raw=20000, n_surf_raw=90000)
    ref = api.Context(0)
    ref.map_set_raw(c["map_corner_raw"], c["map_surf_raw"])
    one = api.Context(0)
    with pytest.raises(api.LlbError) as e:
        one.map_shard_info()                                  # no sharded map yet
    assert e.value.status == api.LLB_ERR_STATE
    one.map_set_raw_sharded(c["map_corner_raw"], c["map_surf_raw"], 0, 1)
    info = one.map_shard_info()
    assert info.world == 1 and info.lo < -1e30 and info.hi > 1e30
    for k in range(2):
        G, L = ref.map_get_ds(k), one.map_get_ds(k)
        assert np.array_equal(G.view(np.uint32), L.view(np.uint32)) and info.ds_owned[k] == G.shape[0] == info.ds_local[k]
    for rank, world in ((2, 2), (-1, 2), (0, 0), (0, 9)):
        with pytest.raises(api.LlbError) as e:
            one.map_set_raw_sharded(c["map_corner_raw"], c["map_surf_raw"], rank, world)
        assert e.value.status == api.LLB_ERR_INVALID
    # an unsharded setter ends the sharded mode
    one.map_set_raw(c["map_corner_raw"], c["map_surf_raw"])
    with pytest.raises(api.LlbError):
        one.map_shard_info()
    # empty raw maps: nothing to own, nothing crashes
    z = np.zeros((0, 4), np.float32)
    one.map_set_raw_sharded(z, z, 1, 2)
    info = one.map_shard_info()
    assert info.ds_owned[0] == 0 and info.ds_owned[1] == 0 and info.raw_kept[0] == 0
    one.close(); ref.close()
