// shard.cuh — BASELINE config 4 with the MAP sharded (preferred form): the raw local map is cut into slabs
// along its longest horizontal axis, rank r keeps the raw points of the voxels that can hold a centroid within the
// kNN gate (1 m, MO:1101 / MO:1183) of a query inside ITS slab, voxel-filters them on the lattice of the WHOLE map
// (bounds from all raw points, so voxel membership, order and centroids are those of the unsharded filter) and indexes
// only that part.  A query belongs to the rank whose slab contains its mapped position; its accepted neighbourhood
// (5 neighbours within the gate) is complete on that rank, so no candidate merge is needed and the only exchange per LM
// iteration stays the 28 fp64 sums.
#pragma once
#include "common.cuh"

namespace llb {

struct ShardPlan {
    int axis = -1;                 // 0 / 1 / 2; -1: not sharded
    float lo = -FLT_MAX, hi = FLT_MAX;   // this rank owns the queries with lo <= coordinate < hi
    int rank = 0, world = 1;
};

// out[3 * k + a] = coordinate a of point k * stride (k < nsamp): the deterministic sample the slab borders are the
// quantiles of (every rank draws the same sample from the same raw map)
int launch_shard_sample(const float4 *pts, int n, int stride, int nsamp, float *out, cudaStream_t s);
// stable compaction of the points whose lattice coordinate floorf(c * inv) on `axis` lies in [ilo, ihi] (whole voxels:
// every point of a kept voxel is kept, in input order).  blk: scratch of div_up(n, 1024) + 1 ints.  Returns launches.
int launch_shard_compact(const float4 *in, int n, int axis, float inv, int ilo, int ihi, float4 *out, int *n_out_dev,
                         int *blk, cudaStream_t s);
// *out = number of the n_dev[0] points of ds whose coordinate on `axis` lies in [lo, hi): the centroids this rank OWNS
// (halo excluded); the sum over the ranks is the size of the unsharded DS map (guard MO:1331)
int launch_shard_count_owned(const float4 *ds, const int *n_dev, int n_upper, int axis, float lo, float hi, int *out,
                             cudaStream_t s);

}  // namespace llb
