"""Multi-GPU plumbing for the hot path (torch.distributed is plumbing, not the product).

Two cases:
* independent sequences / scans (BASELINE config 5): replicas, one process per GPU, no collective;
* one registration against a large map (BASELINE config 4), two forms: (a) the MAP is sharded
  (set_sharded_map: every rank voxel-filters and indexes a slab of the raw map on the lattice of
  the whole map and takes the queries whose mapped position lies in its slab); (b) the queries
  of the scan are sharded round-robin over the ranks (query i belongs to rank i % world) and
  every rank holds the whole voxel-DS map (2M points = 32 MB).  In both the only exchange per LM
  iteration is an all-reduce(sum)
  of 28 fp64 values (21 upper-triangular J^T J terms, 6 J^T r terms, the row count).  Every rank
  then performs the identical 6x6 LM step redundantly, so no broadcast is needed.
"""
from __future__ import annotations

import numpy as np

N_ACC = 28


def shard_queries(n_queries: int, rank: int, world: int) -> np.ndarray:
    """Indices of the queries rank `rank` accumulates (the kernel's qi = rank + world * j)."""
    return np.arange(rank, n_queries, world)


def pair_table():
    """(ia, ib) of the 28 accumulated products of v = {J0..J5, b, 1} (same order as the kernel)."""
    pairs = [(i, j) for i in range(6) for j in range(i, 6)]
    pairs += [(i, 6) for i in range(6)]
    pairs.append((7, 7))
    return pairs


def normal_equations_from_sums(acc: np.ndarray):
    """28 fp64 sums -> (AtA float32 6x6, AtB float32 6, n_rows): one rounding, like cv::gemm."""
    acc = np.asarray(acc, np.float64)
    A = np.zeros((6, 6), np.float32)
    k = 0
    for i in range(6):
        for j in range(i, 6):
            A[i, j] = A[j, i] = np.float32(acc[k]); k += 1
    B = acc[21:27].astype(np.float32)
    return A, B, int(acc[27])


def partial_sums(rows: np.ndarray) -> np.ndarray:
    """fp64 partial sums of the 28 products over `rows` (n,7) = {J0..J5, b} float32."""
    v = np.concatenate([rows.astype(np.float64), np.ones((rows.shape[0], 1))], axis=1)
    return np.array([np.sum(v[:, a] * v[:, b]) for a, b in pair_table()], np.float64)


def reduce_normal_equations(acc, group=None):
    """all-reduce(sum) of the 28-value accumulator (torch tensor on any backend: nccl on GPUs, gloo in tests)."""
    import torch.distributed as dist
    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


class _DevView:
    """CUDA array interface view of a raw device pointer (the context's 28-double accumulator)."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def sharded_scan2map(ctx, T_init, rank: int, world: int, group=None, max_iterations: int = 10):
    """scan2MapOptimization (MO:1329-1350) with the queries sharded over `world` ranks.

    ctx must already hold the (replicated) map and the down-sampled scan.  Returns (pose, iterations).
    """
    import torch
    stream = torch.cuda.ExternalStream(ctx.stream, device=ctx.device)
    ctx.s2m_pose_set(T_init)
    iters = 0
    with torch.cuda.stream(stream):
        for it in range(max_iterations):
            ptr = ctx.s2m_accumulate(it, rank, world)
            acc = torch.as_tensor(_DevView(ptr, N_ACC), device=torch.device("cuda", ctx.device))
            if world > 1:
                reduce_normal_equations(acc, group)
            iters += 1
            if ctx.s2m_solve(it, want_converged=True):
                break
    return ctx.s2m_pose_get(), iters


def setup_fused_exchange(ctx, rank: int, world: int, group=None):
    """One-off set-up of the fused NVLink exchange: all-gather the cudaIpc handles of the ranks' mailboxes
    (torch.distributed is the plumbing) and map the peers' mailboxes into this context."""
    import torch.distributed as dist
    mine = ctx.p2p_export()
    handles = [None] * world
    if world > 1:
        dist.all_gather_object(handles, mine, group=group)
    else:
        handles = [mine]
    ctx.p2p_import(rank, world, handles)


def set_sharded_map(ctx, corner_raw, surf_raw, rank: int, world: int, group=None, device_ptrs=None):
    """Config 4 with the MAP sharded (preferred form): every rank is handed the same raw local map and keeps,
    voxel-filters and indexes only its slab (+ 1 m halo) of it; the sizes of the unsharded DS maps (guard MO:1331) are the
    all-reduced counts of the centroids each rank owns.  device_ptrs = (corner_ptr, rc, surf_ptr, rs) hands over float4
    clouds already in HBM (replicated key-frame stores) instead of host arrays.  Returns the rank's ShardInfo."""
    if device_ptrs is not None:
        ctx.map_set_raw_sharded_dev(*device_ptrs, rank, world)
    else:
        ctx.map_set_raw_sharded(corner_raw, surf_raw, rank, world)
    info = ctx.map_shard_info()
    owned = np.array([info.ds_owned[0], info.ds_owned[1]], np.int64)
    if world > 1:
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", ctx.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.as_tensor(owned, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        owned = t.cpu().numpy()
    ctx.map_shard_set_global(int(owned[0]), int(owned[1]))
    return info


def sharded_scan2map_fused(ctx, T_init):
    """scan2MapOptimization with the queries sharded over the ranks and the 28-value exchange fused into the persistent
    kernel (P2P stores over NVLink, no NCCL call inside the loop).  Needs setup_fused_exchange() once.
    Returns (pose, Stats)."""
    return ctx.s2m_optimize_sharded(T_init)
