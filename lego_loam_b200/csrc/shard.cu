// shard.cu — kernels of the sharded local map (shard.cuh): sample, stable slab compaction, owned-centroid count.
#include "shard.cuh"

namespace llb {

namespace {

constexpr int SH_THREADS = 256;
constexpr int SH_ITEMS = 4;
constexpr int SH_TILE = SH_THREADS * SH_ITEMS;      // points per CTA of the compaction

__device__ __forceinline__ float coord_of(const float4 &p, int axis) { return axis == 0 ? p.x : (axis == 1 ? p.y : p.z); }

__global__ void shard_sample_kernel(const float4 *__restrict__ pts, int n, int stride, int nsamp, float *__restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nsamp) return;
    const long long i = (long long)k * stride;
    const float4 p = i < n ? __ldg(&pts[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
    out[3 * k] = p.x; out[3 * k + 1] = p.y; out[3 * k + 2] = p.z;
}

__device__ __forceinline__ bool in_slab(const float4 &p, int axis, float inv, int ilo, int ihi)
{
    const float f = floorf(coord_of(p, axis) * inv);           // the lattice coordinate PCL derives the voxel from
    return f >= (float)ilo && f <= (float)ihi;
}

__global__ void __launch_bounds__(SH_THREADS)
shard_flag_count_kernel(const float4 *__restrict__ in, int n, int axis, float inv, int ilo, int ihi, int *__restrict__ blk)
{
    int c = 0;
#pragma unroll
    for (int r = 0; r < SH_ITEMS; r++) {
        const int i = blockIdx.x * SH_TILE + r * SH_THREADS + threadIdx.x;
        if (i < n) c += in_slab(__ldg(&in[i]), axis, inv, ilo, ihi) ? 1 : 0;
    }
    __shared__ int s_w[SH_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < SH_THREADS / 32; k++) t += s_w[k];
        blk[blockIdx.x] = t;
    }
}

// exclusive scan of blk[0 .. count) in place by one CTA; blk[count] and *total receive the sum
__global__ void __launch_bounds__(1024)
shard_scan_kernel(int *__restrict__ blk, int count, int *__restrict__ total_out)
{
    __shared__ int s_scan[33];
    const int per = (count + 1023) / 1024;
    const int lo = min((int)threadIdx.x * per, count), hi = min(lo + per, count);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += blk[i];
    int total;
    int base = block_excl_scan(sum, s_scan, total);
    for (int i = lo; i < hi; i++) { const int v = blk[i]; blk[i] = base; base += v; }
    if (threadIdx.x == 0) { blk[count] = total; *total_out = total; }
}

__global__ void __launch_bounds__(SH_THREADS)
shard_scatter_kernel(const float4 *__restrict__ in, int n, int axis, float inv, int ilo, int ihi,
                     const int *__restrict__ blk, float4 *__restrict__ out)
{
    // item order inside a CTA = (round r, thread): the rank of a kept point is the number of kept points before it in
    // that order, which is the input order -> stable
    __shared__ int s_scan[33];
    int base = blk[blockIdx.x];
#pragma unroll
    for (int r = 0; r < SH_ITEMS; r++) {
        const int i = blockIdx.x * SH_TILE + r * SH_THREADS + threadIdx.x;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        int keep = 0;
        if (i < n) { p = __ldg(&in[i]); keep = in_slab(p, axis, inv, ilo, ihi) ? 1 : 0; }
        int total;
        const int pos = block_excl_scan(keep, s_scan, total);
        if (keep) out[base + pos] = p;
        base += total;
    }
}

__global__ void __launch_bounds__(SH_THREADS)
shard_count_owned_kernel(const float4 *__restrict__ ds, const int *__restrict__ n_dev, int axis, float lo, float hi,
                         int *__restrict__ out)
{
    const int n = *n_dev;
    int c = 0;
    for (int i = blockIdx.x * SH_THREADS + threadIdx.x; i < n; i += gridDim.x * SH_THREADS) {
        const float v = coord_of(__ldg(&ds[i]), axis);
        c += (v >= lo && v < hi) ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

__global__ void shard_zero_kernel(int *p) { *p = 0; }

}  // namespace

int launch_shard_sample(const float4 *pts, int n, int stride, int nsamp, float *out, cudaStream_t s)
{
    if (nsamp <= 0) return 0;
    shard_sample_kernel<<<div_up(nsamp, 256), 256, 0, s>>>(pts, n, stride, nsamp, out);
    LLB_CUDA(cudaGetLastError());
    return 1;
}

int launch_shard_compact(const float4 *in, int n, int axis, float inv, int ilo, int ihi, float4 *out, int *n_out_dev,
                         int *blk, cudaStream_t s)
{
    const int nblk = std::max(1, div_up(n, SH_TILE));
    shard_flag_count_kernel<<<nblk, SH_THREADS, 0, s>>>(in, n, axis, inv, ilo, ihi, blk);
    shard_scan_kernel<<<1, 1024, 0, s>>>(blk, nblk, n_out_dev);
    shard_scatter_kernel<<<nblk, SH_THREADS, 0, s>>>(in, n, axis, inv, ilo, ihi, blk, out);
    LLB_CUDA(cudaGetLastError());
    return 3;
}

int launch_shard_count_owned(const float4 *ds, const int *n_dev, int n_upper, int axis, float lo, float hi, int *out,
                             cudaStream_t s)
{
    shard_zero_kernel<<<1, 1, 0, s>>>(out);
    shard_count_owned_kernel<<<std::max(1, std::min(div_up(n_upper, SH_THREADS), 148 * 8)), SH_THREADS, 0, s>>>(ds, n_dev, axis, lo, hi, out);
    LLB_CUDA(cudaGetLastError());
    return 2;
}

}  // namespace llb
