// voxel_small.cu — K1, small path: one thread-block CLUSTER (8 CTAs on 8 SMs, distributed
// shared memory) per pcl::VoxelGrid filter of up to 16384 points -> one launch for the three
// independent filters of downsampleCurrentScan (MO:1069-1082) and one for the fourth (MO:1084-1090).
//
// Why a cluster: the filter is a sort of (voxel index << 32 | point index) followed by ordered
// per-voxel sums.  On ONE SM the bitonic network is issue-bound (round-1 profile: ~30 us for 8192
// keys, 66 us per launch); spread over 8 SMs every CTA owns P/8 keys in its own shared memory, the
// stages with partner distance < P/8 stay CTA-local, and the 6 stages that cross CTAs read the
// partner's element through DSMEM between two hardware cluster barriers.  Head detection, the scan
// of the head counts and the segment walks read neighbouring CTAs' keys the same way.
//
// Exactness: PCL's key expression, stable order (the low key word is the input index), strictly
// sequential float sums per voxel, true division by (float)count, int32-overflow pass-through.
#include "voxel_dev.cuh"
#include "cta_radix.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace llb {

namespace {

constexpr int CL = 8;                        // CTAs per cluster
constexpr int TPB = 1024;
constexpr int LMAX = VoxelFilter::SMALL_MAX / CL;   // keys per CTA

// one filter by one cluster; called by the two kernel wrappers below
__device__ __forceinline__ void voxel_small_body(const SegIn in, const float leaf, float4 *__restrict__ out,
                                                 int *__restrict__ n_out_dev)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int cr = (int)cluster.block_rank();

    __shared__ unsigned long long skey[LMAX];
    __shared__ float s_red[6][32];
    __shared__ float s_mm[6];               // this CTA's min xyz, max xyz
    __shared__ int s_scan[33];
    __shared__ int s_heads;                 // head count of this CTA
    __shared__ float s_inv;
    __shared__ int s_min_b[3], s_div_b[3], s_mul[3], s_overflow, s_nbits;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int na = seg_len_a(in), n = na + seg_len_b(in);
    if (n <= 0) {                            // uniform over the cluster
        if (cr == 0 && tid == 0) *n_out_dev = 0;
        return;
    }
    int P = CL * 32;
    while (P < n) P <<= 1;
    const int L = P / CL;                    // keys held by each CTA
    const int g0 = cr * L;                   // first global position of this CTA

    // ---- bounds: CTA-partial over a strided slice, then all-to-all through DSMEM
    float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int i = cr * TPB + tid; i < n; i += CL * TPB) {
        float4 p = seg_load(in, na, i);
        mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
        mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
        mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(FULL, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL, mx[a], o));
        }
        if (lane == 0) { s_red[a][w] = mn[a]; s_red[3 + a][w] = mx[a]; }
    }
    __syncthreads();
    if (tid < 6) {
        float m = s_red[tid][0];
        for (int k = 1; k < TPB / 32; k++) m = tid < 3 ? fminf(m, s_red[tid][k]) : fmaxf(m, s_red[tid][k]);
        s_mm[tid] = m;
    }
    cluster.sync();
    if (tid == 0) {
        float fmn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, fmx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
        for (int r = 0; r < CL; r++) {
            const float *rm = cluster.map_shared_rank(s_mm, r);
            for (int a = 0; a < 3; a++) { fmn[a] = fminf(fmn[a], rm[a]); fmx[a] = fmaxf(fmx[a], rm[3 + a]); }
        }
        float inv = 1.0f / leaf;
        int ovf, nb;
        voxel_setup(inv, fmn, fmx, n, s_min_b, s_div_b, s_mul, ovf, nb);
        s_inv = inv; s_overflow = ovf; s_nbits = nb;
    }
    __syncthreads();
    if (s_overflow) {                        // PCL: output = input (uniform over the cluster: same inputs)
        for (int i = cr * TPB + tid; i < n; i += CL * TPB) out[i] = seg_load(in, na, i);
        if (cr == 0 && tid == 0) *n_out_dev = n;
        cluster.sync();                      // nobody leaves while s_mm may still be read remotely
        return;
    }
    // ---- keys: global position g = g0 + t
    {
        const float inv = s_inv;
        int min_b[3] = { s_min_b[0], s_min_b[1], s_min_b[2] }, mul[3] = { s_mul[0], s_mul[1], s_mul[2] };
        for (int t = tid; t < L; t += TPB) {
            const int g = g0 + t;
            unsigned long long k = ~0ull;
            if (g < n) {
                float4 p = seg_load(in, na, g);
                k = ((unsigned long long)voxel_key(p, inv, min_b, mul) << 32) | (unsigned)g;
            }
            skey[t] = k;
        }
    }
    // ---- bitonic sort over the cluster (keys are unique => the result is the stable order)
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j < L) {                                         // partner inside this CTA
                __syncthreads();
                for (int t = tid; t < (L >> 1); t += TPB) {
                    const int li = 2 * t - (t & (j - 1));
                    const int lx = li + j;
                    const unsigned long long a = skey[li], b = skey[lx];
                    const bool up = ((g0 + li) & k) == 0;
                    if ((a > b) == up) { skey[li] = b; skey[lx] = a; }
                }
            } else {                                             // partner in CTA cr ^ (j / L), same local slot
                cluster.sync();
                const unsigned long long *rk = cluster.map_shared_rank(skey, cr ^ (j / L));
                unsigned long long nv[LMAX / TPB];
#pragma unroll
                for (int u = 0; u < LMAX / TPB; u++) {
                    const int t = tid + u * TPB;
                    if (t < L) {
                        const unsigned long long mine = skey[t], other = rk[t];
                        const int g = g0 + t;
                        const bool up = (g & k) == 0, lower = (g & j) == 0;
                        const bool keep_min = (lower == up);
                        nv[u] = keep_min ? (mine < other ? mine : other) : (mine > other ? mine : other);
                    }
                }
                cluster.sync();
#pragma unroll
                for (int u = 0; u < LMAX / TPB; u++) {
                    const int t = tid + u * TPB;
                    if (t < L) skey[t] = nv[u];
                }
            }
        }
    }
    cluster.sync();
    // ---- heads of this CTA's slice; key at global position g lives in CTA g / L, slot g % L
    auto key_hi = [&](int g) -> unsigned {
        const int r = g / L;
        const unsigned long long *p = (r == cr) ? skey : cluster.map_shared_rank(skey, r);
        return (unsigned)(p[g - r * L] >> 32);
    };
    auto key_lo = [&](int g) -> unsigned {
        const int r = g / L;
        const unsigned long long *p = (r == cr) ? skey : cluster.map_shared_rank(skey, r);
        return (unsigned)p[g - r * L];
    };
    const int lcount = max(0, min(L, n - g0));                   // valid keys in this CTA
    const int chunk = (lcount + TPB - 1) / TPB;
    const int lo = min(tid * chunk, lcount), hi = min(lo + chunk, lcount);
    int heads = 0;
    for (int t = lo; t < hi; t++) {
        const int g = g0 + t;
        heads += (g == 0) || (key_hi(g - 1) != (unsigned)(skey[t] >> 32));
    }
    int total;
    int rank = block_excl_scan(heads, s_scan, total);
    if (tid == 0) s_heads = total;
    cluster.sync();
    int base = 0, grand = 0;
    for (int r = 0; r < CL; r++) {
        const int c = *cluster.map_shared_rank(&s_heads, r);
        if (r < cr) base += c;
        grand += c;
    }
    rank += base;
    for (int t = lo; t < hi; t++) {
        const int g = g0 + t;
        const unsigned cur = (unsigned)(skey[t] >> 32);
        if ((g == 0) || (key_hi(g - 1) != cur)) {
            float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
            int j = g;
            while (j < n && key_hi(j) == cur) {
                float4 p = seg_load(in, na, (int)key_lo(j));
                sx += p.x; sy += p.y; sz += p.z; si += p.w;
                j++;
            }
            const float cnt = (float)(j - g);
            out[rank++] = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
        }
    }
    if (cr == 0 && tid == 0) *n_out_dev = grand;
    cluster.sync();                          // keep shared memory alive until every remote read is done
}

// ------------------------------------------------------------------------------------------------
// Throughput form (batched multi-registration path): ONE CTA per filter, stable LSD radix sort of the 32-bit voxel
// keys in shared memory (8-bit digits, only the significant key bits; per-warp digit counters + __match_any ranking
// keep equal keys in input order, which is what the 64-bit bitonic keys of the cluster kernel encode explicitly).
// ~5x fewer instructions per filter than the bitonic network and one SM instead of eight, so the 4B filters of a
// B-slot batch fill the GPU in about one wave.  Dynamic shared memory: cap x (2 x 4 B keys + 2 x 2 B indices).
constexpr int VC_THREADS = 512;
constexpr int VC_NW = VC_THREADS / 32;
constexpr int VC_ITEMS = 4;

__global__ void __launch_bounds__(VC_THREADS)
voxel_cta_kernel(const SmallJob *__restrict__ jobs, int cap)
{
    extern __shared__ __align__(16) unsigned char vc_smem[];
    unsigned *kin = reinterpret_cast<unsigned *>(vc_smem), *kout = kin + cap;
    unsigned short *vin = reinterpret_cast<unsigned short *>(kout + cap), *vout = vin + cap;
    __shared__ int s_wcnt[VC_NW][256];
    __shared__ int s_base[256];
    __shared__ float s_red[6][VC_NW];
    __shared__ int s_scan[33];
    __shared__ float s_inv;
    __shared__ int s_min_b[3], s_div_b[3], s_mul[3], s_overflow, s_nbits;

    const SmallJob jb = jobs[blockIdx.x];
    const SegIn in = jb.in;
    float4 *__restrict__ out = jb.out;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int na = seg_len_a(in), n = na + seg_len_b(in);
    if (n <= 0 || n > cap) {                 // n > cap cannot happen (the host sizes cap from the upper bounds)
        if (tid == 0) *jb.n_out = 0;
        return;
    }
    // ---- bounds
    float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int i = tid; i < n; i += VC_THREADS) {
        const float4 p = seg_load(in, na, i);
        mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
        mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
        mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(FULL, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL, mx[a], o));
        }
        if (lane == 0) { s_red[a][w] = mn[a]; s_red[3 + a][w] = mx[a]; }
    }
    __syncthreads();
    if (tid == 0) {
        float fmn[3], fmx[3];
        for (int a = 0; a < 3; a++) {
            fmn[a] = s_red[a][0]; fmx[a] = s_red[3 + a][0];
            for (int k = 1; k < VC_NW; k++) { fmn[a] = fminf(fmn[a], s_red[a][k]); fmx[a] = fmaxf(fmx[a], s_red[3 + a][k]); }
        }
        const float inv = 1.0f / jb.leaf;
        int ovf, nb;
        voxel_setup(inv, fmn, fmx, n, s_min_b, s_div_b, s_mul, ovf, nb);
        s_inv = inv; s_overflow = ovf; s_nbits = nb;
    }
    __syncthreads();
    if (s_overflow) {                        // PCL: output = input (C18)
        for (int i = tid; i < n; i += VC_THREADS) out[i] = seg_load(in, na, i);
        if (tid == 0) *jb.n_out = n;
        return;
    }
    // ---- keys
    {
        const float inv = s_inv;
        const int min_b[3] = { s_min_b[0], s_min_b[1], s_min_b[2] }, mul[3] = { s_mul[0], s_mul[1], s_mul[2] };
        for (int i = tid; i < n; i += VC_THREADS) {
            kin[i] = voxel_key(seg_load(in, na, i), inv, min_b, mul);
            vin[i] = (unsigned short)i;
        }
    }
    // ---- stable LSD radix sort, 8 bits per pass (cta_radix.cuh)
    cta_radix_sort<VC_THREADS, VC_ITEMS>(kin, kout, vin, vout, n, s_nbits, s_wcnt, s_base, s_scan);
    // ---- heads, scan, ordered per-voxel sums
    const int chunk = (n + VC_THREADS - 1) / VC_THREADS;
    const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    int heads = 0;
    for (int t = lo; t < hi; t++) heads += (t == 0) || (kin[t - 1] != kin[t]);
    int total;
    int rank = block_excl_scan(heads, s_scan, total);
    for (int t = lo; t < hi; t++) {
        const unsigned cur = kin[t];
        if ((t == 0) || (kin[t - 1] != cur)) {
            float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
            int j = t;
            while (j < n && kin[j] == cur) {
                const float4 p = seg_load(in, na, (int)vin[j]);
                sx += p.x; sy += p.y; sz += p.z; si += p.w;
                j++;
            }
            const float cnt = (float)(j - t);
            out[rank++] = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
        }
    }
    if (tid == 0) *jb.n_out = total;
}

// jobs by value (<= MAX_BATCH filters: the single-context path)
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(TPB, 1)
voxel_small_kernel(SmallJobs jobs)
{
    const int job = blockIdx.x / CL;
    voxel_small_body(jobs.in[job], jobs.leaf[job], jobs.out[job], jobs.n_out[job]);
}

// device-resident job table (any number of filters: the batched multi-registration path)
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(TPB, 1)
voxel_small_jobs_kernel(const SmallJob *__restrict__ jobs)
{
    const SmallJob jb = jobs[blockIdx.x / CL];
    voxel_small_body(jb.in, jb.leaf, jb.out, jb.n_out);
}

}  // namespace

void launch_voxel_cta_jobs(const SmallJob *jobs_dev, int count, int cap, cudaStream_t stream)
{
    const int bytes = cap * 12;
    if (bytes > 48 * 1024)                                   // opt-in above the default limit (idempotent, per device)
        LLB_CUDA(cudaFuncSetAttribute(voxel_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      VoxelFilter::SMALL_MAX * 12));
    voxel_cta_kernel<<<count, VC_THREADS, bytes, stream>>>(jobs_dev, cap);
    LLB_CUDA(cudaGetLastError());
}

void launch_voxel_small_jobs(const SmallJob *jobs_dev, int count, cudaStream_t stream)
{
    voxel_small_jobs_kernel<<<count * CL, TPB, 0, stream>>>(jobs_dev);
    LLB_CUDA(cudaGetLastError());
}

void launch_voxel_small(const SmallJobs &jobs, int count, cudaStream_t stream)
{
    voxel_small_kernel<<<count * CL, TPB, 0, stream>>>(jobs);
    LLB_CUDA(cudaGetLastError());
}

}  // namespace llb
