// voxel_dev.cuh — device helpers shared by the two voxel paths (voxel.cu, voxel_small.cu).
#pragma once
#include "voxel.cuh"

namespace llb {

struct SegIn {
    const float4 *a; const int *na_dev; int na;
    const float4 *b; const int *nb_dev; int nb;
};

__device__ __forceinline__ int seg_len_a(const SegIn &s) { return s.na_dev ? *s.na_dev : s.na; }
__device__ __forceinline__ int seg_len_b(const SegIn &s) { return s.b ? (s.nb_dev ? *s.nb_dev : s.nb) : 0; }
__device__ __forceinline__ float4 seg_load(const SegIn &s, int na, int i)
{
    return i < na ? __ldg(&s.a[i]) : __ldg(&s.b[i - na]);
}

struct LargeVoxelJob {       // one filter of the multi-kernel (radix) path: input, parameters, scratch, output
    SegIn in; float leaf; VoxelDesc *desc;
    SegIn bounds;            // the cloud whose min / max define the lattice: `in` itself, or - sharded maps - the WHOLE raw
                             // map of which `in` is this rank's part, so that every rank numbers the voxels alike
    unsigned *kA, *kB; int *vA, *vB; int *hist; int *blk;
    float4 *out; int *n_out;
};

inline __host__ SegIn to_seg(const VoxelInput &in)
{
    SegIn s;
    s.a = in.a; s.na_dev = in.na_dev; s.na = in.na;
    s.b = in.b; s.nb_dev = in.nb_dev; s.nb = in.nb;
    return s;
}

// PCL's grid set-up from the cloud bounds (VoxelGrid::applyFilter, A.1 steps 1-3)
__device__ inline void voxel_setup(float inv, const float mn[3], const float mx[3], int n,
                            int min_b[3], int div_b[3], int mul[3], int &overflow, int &nbits)
{
    long long d[3];
    for (int a = 0; a < 3; a++) d[a] = (long long)((mx[a] - mn[a]) * inv) + 1;
    overflow = (d[0] * d[1] * d[2] > (long long)INT_MAX) ? 1 : 0;
    for (int a = 0; a < 3; a++) {
        min_b[a] = (int)floorf(mn[a] * inv);
        int max_b = (int)floorf(mx[a] * inv);
        div_b[a] = max_b - min_b[a] + 1;
    }
    mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];
    unsigned long long maxkey;
    if (overflow) maxkey = n > 0 ? (unsigned long long)(n - 1) : 0;       // pass-through: key = index
    else maxkey = (unsigned long long)div_b[0] * (unsigned long long)div_b[1] * (unsigned long long)div_b[2] - 1;
    nbits = 1;
    while (nbits < 32 && (maxkey >> nbits) != 0) nbits++;
}

__device__ __forceinline__ unsigned voxel_key(const float4 &p, float inv, const int *min_b, const int *mul)
{
    int i0 = (int)(floorf(p.x * inv) - (float)min_b[0]);
    int i1 = (int)(floorf(p.y * inv) - (float)min_b[1]);
    int i2 = (int)(floorf(p.z * inv) - (float)min_b[2]);
    return (unsigned)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]);
}


struct SmallJobs {
    SegIn in[VoxelFilter::MAX_BATCH];
    float leaf[VoxelFilter::MAX_BATCH];
    float4 *out[VoxelFilter::MAX_BATCH];
    int *n_out[VoxelFilter::MAX_BATCH];
};

struct SmallJob { SegIn in; float leaf; float4 *out; int *n_out; };
// the same with a device-resident job table of any length (batched multi-registration path)
void launch_voxel_small_jobs(const SmallJob *jobs_dev, int count, cudaStream_t stream);

// throughput form: ONE CTA per job (shared-memory radix sort); every job must have at most `cap` points
// (cap <= VoxelFilter::SMALL_MAX, dynamic shared memory = 12 B x cap)
void launch_voxel_cta_jobs(const SmallJob *jobs_dev, int count, int cap, cudaStream_t stream);

// one 8-CTA cluster per job; every job must have at most VoxelFilter::SMALL_MAX points
void launch_voxel_small(const SmallJobs &jobs, int count, cudaStream_t stream);

}  // namespace llb
