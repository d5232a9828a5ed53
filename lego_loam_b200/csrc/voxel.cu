// voxel.cu — K1 kernels.  Two paths:
//  * small (<= 16384 points, every current-scan cloud of a VLP-16/HDL-32E): ONE CTA does
//    min/max, key generation, a bitonic sort of (voxel index << 32 | point index) in shared
//    memory, head flags + scan, and the ordered per-voxel sums -> 1 launch per filter;
//  * large (local maps, VLS-128 scans): min/max -> setup -> keys -> stable LSD radix sort
//    (8-bit digits, per-block histograms, warp-match ranking) -> head count -> scan ->
//    ordered per-voxel sums.
// The float arithmetic that decides membership is written exactly as PCL does it
// (float multiply, floorf, float subtract, int cast) and this file is compiled with
// -fmad=false, so voxel membership, order and centroids are bit-identical to the oracle.
#include "voxel.cuh"

namespace llb {

namespace {

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

__host__ SegIn to_seg(const VoxelInput &in)
{
    SegIn s;
    s.a = in.a; s.na_dev = in.na_dev; s.na = in.na;
    s.b = in.b; s.nb_dev = in.nb_dev; s.nb = in.nb;
    return s;
}

// PCL's grid set-up from the cloud bounds (VoxelGrid::applyFilter, A.1 steps 1-3)
__device__ void voxel_setup(float inv, const float mn[3], const float mx[3], int n,
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

// ------------------------------------------------------------------ small path

constexpr int SMALL_THREADS = 1024;

struct SmallJobs {
    SegIn in[VoxelFilter::MAX_BATCH];
    float leaf[VoxelFilter::MAX_BATCH];
    float4 *out[VoxelFilter::MAX_BATCH];
    int *n_out[VoxelFilter::MAX_BATCH];
};

// one CTA per job (blockIdx.x): independent filters share a launch
__global__ void __launch_bounds__(SMALL_THREADS, 1)
voxel_small_kernel(SmallJobs jobs)
{
    const SegIn in = jobs.in[blockIdx.x];
    const float leaf = jobs.leaf[blockIdx.x];
    float4 *__restrict__ out = jobs.out[blockIdx.x];
    int *__restrict__ n_out_dev = jobs.n_out[blockIdx.x];
    extern __shared__ unsigned long long skey[];          // cap_pow2 entries
    __shared__ float s_red[6][32];
    __shared__ int s_scan[33];
    __shared__ float s_inv;
    __shared__ int s_min_b[3], s_div_b[3], s_mul[3], s_overflow, s_nbits;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int na = seg_len_a(in), n = na + seg_len_b(in);
    if (n <= 0) {
        if (tid == 0) *n_out_dev = 0;
        return;
    }
    // power of two >= n (the host sized the shared memory for its upper bound)
    int P = 32;
    while (P < n) P <<= 1;

    // ---- bounds
    float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int i = tid; i < n; i += SMALL_THREADS) {
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
    if (tid == 0) {
        float fmn[3], fmx[3];
        for (int a = 0; a < 3; a++) {
            fmn[a] = s_red[a][0]; fmx[a] = s_red[3 + a][0];
            for (int k = 1; k < SMALL_THREADS / 32; k++) {
                fmn[a] = fminf(fmn[a], s_red[a][k]); fmx[a] = fmaxf(fmx[a], s_red[3 + a][k]);
            }
        }
        float inv = 1.0f / leaf;
        int ovf, nb;
        voxel_setup(inv, fmn, fmx, n, s_min_b, s_div_b, s_mul, ovf, nb);
        s_inv = inv; s_overflow = ovf; s_nbits = nb;
    }
    __syncthreads();
    if (s_overflow) {                                      // PCL: output = input
        for (int i = tid; i < n; i += SMALL_THREADS) out[i] = seg_load(in, na, i);
        if (tid == 0) *n_out_dev = n;
        return;
    }
    // ---- keys
    {
        const float inv = s_inv;
        int min_b[3] = { s_min_b[0], s_min_b[1], s_min_b[2] }, mul[3] = { s_mul[0], s_mul[1], s_mul[2] };
        for (int i = tid; i < P; i += SMALL_THREADS) {
            unsigned long long k = ~0ull;
            if (i < n) {
                float4 p = seg_load(in, na, i);
                k = ((unsigned long long)voxel_key(p, inv, min_b, mul) << 32) | (unsigned)i;
            }
            skey[i] = k;
        }
    }
    __syncthreads();
    // ---- bitonic sort (keys are unique: the low word is the point index => stable)
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += SMALL_THREADS) {
                int i = 2 * t - (t & (j - 1));
                int x = i + j;
                unsigned long long a = skey[i], b = skey[x];
                bool up = (i & k) == 0;
                if ((a > b) == up) { skey[i] = b; skey[x] = a; }
            }
            __syncthreads();
        }
    }
    // ---- heads: contiguous chunk per thread
    const int chunk = (n + SMALL_THREADS - 1) / SMALL_THREADS;
    const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    int heads = 0;
    for (int i = lo; i < hi; i++) {
        unsigned cur = (unsigned)(skey[i] >> 32);
        heads += (i == 0) || ((unsigned)(skey[i - 1] >> 32) != cur);
    }
    int total;
    int rank = block_excl_scan(heads, s_scan, total);
    for (int i = lo; i < hi; i++) {
        unsigned cur = (unsigned)(skey[i] >> 32);
        if ((i == 0) || ((unsigned)(skey[i - 1] >> 32) != cur)) {
            float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
            int j = i;
            while (j < n && (unsigned)(skey[j] >> 32) == cur) {
                float4 p = seg_load(in, na, (int)(unsigned)skey[j]);
                sx += p.x; sy += p.y; sz += p.z; si += p.w;
                j++;
            }
            float cnt = (float)(j - i);
            out[rank++] = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
        }
    }
    if (tid == 0) *n_out_dev = total;
}

// ------------------------------------------------------------------ large path

constexpr int LG_THREADS = 256;
constexpr int RADIX_MAX_BLOCKS = 256;

__global__ void voxel_desc_init_kernel(VoxelDesc *d)
{
    for (int a = 0; a < 3; a++) { d->mn[a] = INT_MAX; d->mx[a] = INT_MIN; }
}

__global__ void __launch_bounds__(LG_THREADS)
voxel_minmax_kernel(SegIn in, VoxelDesc *__restrict__ d)
{
    __shared__ float s_red[6][LG_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int na = seg_len_a(in), n = na + seg_len_b(in);
    float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int i = blockIdx.x * LG_THREADS + tid; i < n; i += gridDim.x * LG_THREADS) {
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
    if (tid < 3) {
        float m0 = s_red[tid][0], m1 = s_red[3 + tid][0];
        for (int k = 1; k < LG_THREADS / 32; k++) { m0 = fminf(m0, s_red[tid][k]); m1 = fmaxf(m1, s_red[3 + tid][k]); }
        if (n > 0) {
            atomicMin(&d->mn[tid], float_to_ordered(m0));
            atomicMax(&d->mx[tid], float_to_ordered(m1));
        }
    }
}

__global__ void voxel_setup_kernel(SegIn in, float leaf, VoxelDesc *d)
{
    const int n = seg_len_a(in) + seg_len_b(in);
    float mn[3], mx[3];
    for (int a = 0; a < 3; a++) {
        mn[a] = ordered_to_float(d->mn[a]); mx[a] = ordered_to_float(d->mx[a]);
        d->mn[a] = INT_MAX; d->mx[a] = INT_MIN;            // ready for the next use
    }
    d->n = n;
    d->inv = 1.0f / leaf;
    d->n_out = 0;
    if (n <= 0) { d->overflow = 0; d->nbits = 0; return; }
    int ovf, nb;
    voxel_setup(d->inv, mn, mx, n, d->min_b, d->div_b, d->mul, ovf, nb);
    d->overflow = ovf; d->nbits = nb;
}

__global__ void __launch_bounds__(LG_THREADS)
voxel_keys_kernel(SegIn in, const VoxelDesc *__restrict__ d, unsigned *__restrict__ keys, int *__restrict__ vals)
{
    const int n = d->n;
    const int na = seg_len_a(in);
    const float inv = d->inv;
    const int min_b[3] = { d->min_b[0], d->min_b[1], d->min_b[2] }, mul[3] = { d->mul[0], d->mul[1], d->mul[2] };
    const int ovf = d->overflow;
    for (int i = blockIdx.x * LG_THREADS + threadIdx.x; i < n; i += gridDim.x * LG_THREADS) {
        float4 p = seg_load(in, na, i);
        keys[i] = ovf ? (unsigned)i : voxel_key(p, inv, min_b, mul);
        vals[i] = i;
    }
}

// --- stable LSD radix sort, one 8-bit digit per pass.  Tile layout is a pure function of
// (n, gridDim) so the histogram and scatter kernels agree.
__device__ __forceinline__ void radix_tile(int n, int nblocks, int b, int &lo, int &hi)
{
    int tile = (n + nblocks - 1) / nblocks;
    tile = (tile + LG_THREADS - 1) / LG_THREADS * LG_THREADS;
    long long l = (long long)b * tile;
    lo = (int)(l < n ? l : n);
    hi = (int)(l + tile < n ? l + tile : n);
}

__global__ void __launch_bounds__(LG_THREADS)
radix_hist_kernel(const VoxelDesc *__restrict__ d, int shift, const unsigned *__restrict__ kA,
                  const unsigned *__restrict__ kB, int *__restrict__ hist)
{
    if (shift >= d->nbits) return;
    const unsigned *keys = ((shift >> 3) & 1) ? kB : kA;
    __shared__ int s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    int lo, hi;
    radix_tile(d->n, gridDim.x, blockIdx.x, lo, hi);
    for (int i = lo + threadIdx.x; i < hi; i += LG_THREADS) atomicAdd(&s_h[(keys[i] >> shift) & 255], 1);
    __syncthreads();
    hist[threadIdx.x * gridDim.x + blockIdx.x] = s_h[threadIdx.x];
}

// exclusive scan of `count` ints in place by ONE block of 1024 threads
__global__ void __launch_bounds__(1024)
scan_single_block_kernel(int *__restrict__ data, int count, const VoxelDesc *d, int shift, int *total_out)
{
    if (d && shift >= 0 && shift >= d->nbits) return;
    __shared__ int s_scan[33];
    const int per = (count + 1023) / 1024;
    const int lo = min((int)threadIdx.x * per, count), hi = min(lo + per, count);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += data[i];
    int total;
    int base = block_excl_scan(sum, s_scan, total);
    for (int i = lo; i < hi; i++) { int v = data[i]; data[i] = base; base += v; }
    if (total_out && threadIdx.x == 0) *total_out = total;
}

__global__ void __launch_bounds__(LG_THREADS)
radix_scatter_kernel(const VoxelDesc *__restrict__ d, int shift, unsigned *__restrict__ kA, unsigned *__restrict__ kB,
                     int *__restrict__ vA, int *__restrict__ vB, const int *__restrict__ hist)
{
    if (shift >= d->nbits) return;
    const bool odd = ((shift >> 3) & 1) != 0;
    const unsigned *kin = odd ? kB : kA; unsigned *kout = odd ? kA : kB;
    const int *vin = odd ? vB : vA; int *vout = odd ? vA : vB;

    constexpr int NW = LG_THREADS / 32;
    __shared__ int s_base[256];
    __shared__ int s_wcnt[NW][256];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    s_base[tid] = hist[tid * gridDim.x + blockIdx.x];
    int lo, hi;
    radix_tile(d->n, gridDim.x, blockIdx.x, lo, hi);
    for (int r = lo; r < hi; r += LG_THREADS) {
#pragma unroll
        for (int k = 0; k < NW; k++) s_wcnt[k][tid] = 0;
        __syncthreads();
        const int i = r + tid;
        const bool valid = i < hi;
        unsigned key = valid ? kin[i] : 0;
        int val = valid ? vin[i] : 0;
        unsigned dig = valid ? ((key >> shift) & 255) : (256 + lane);   // invalid lanes never match
        unsigned m = __match_any_sync(FULL, dig);
        int rank = __popc(m & ((1u << lane) - 1));
        if (valid && rank == 0) s_wcnt[w][dig] = __popc(m);
        __syncthreads();
        {   // thread tid owns digit tid: turn per-warp counts into offsets
            int run = s_base[tid];
#pragma unroll
            for (int k = 0; k < NW; k++) { int c = s_wcnt[k][tid]; s_wcnt[k][tid] = run; run += c; }
            s_base[tid] = run;
        }
        __syncthreads();
        if (valid) {
            int pos = s_wcnt[w][dig] + rank;
            kout[pos] = key; vout[pos] = val;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void sorted_bufs(const VoxelDesc *d, const unsigned *kA, const unsigned *kB,
                                            const int *vA, const int *vB, const unsigned *&k, const int *&v)
{
    int passes = (d->nbits + 7) >> 3;
    if (passes > 4) passes = 4;
    k = (passes & 1) ? kB : kA;
    v = (passes & 1) ? vB : vA;
}

constexpr int HEAD_TILE = 1024;

__global__ void __launch_bounds__(HEAD_TILE)
voxel_heads_kernel(const VoxelDesc *__restrict__ d, const unsigned *kA, const unsigned *kB, int *__restrict__ blk)
{
    const unsigned *k; const int *v;
    sorted_bufs(d, kA, kB, nullptr, nullptr, k, v);
    const int n = d->n;
    const int i = blockIdx.x * HEAD_TILE + threadIdx.x;
    int head = (i < n) && (i == 0 || k[i - 1] != k[i]);
    int cnt = __syncthreads_count(head);
    if (threadIdx.x == 0) blk[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(HEAD_TILE)
voxel_centroid_kernel(SegIn in, const VoxelDesc *__restrict__ d, const unsigned *kA, const unsigned *kB,
                      const int *vA, const int *vB, const int *__restrict__ blk, float4 *__restrict__ out)
{
    __shared__ int s_scan[33];
    const unsigned *k; const int *v;
    sorted_bufs(d, kA, kB, vA, vB, k, v);
    const int n = d->n;
    const int na = seg_len_a(in);
    const int i = blockIdx.x * HEAD_TILE + threadIdx.x;
    int head = (i < n) && (i == 0 || k[i - 1] != k[i]);
    int total;
    int rank = blk[blockIdx.x] + block_excl_scan(head, s_scan, total);
    if (head) {
        const unsigned cur = k[i];
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
        int j = i;
        while (j < n && k[j] == cur) {
            float4 p = seg_load(in, na, v[j]);
            sx += p.x; sy += p.y; sz += p.z; si += p.w;
            j++;
        }
        float cnt = (float)(j - i);
        out[rank] = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
    }
}

}  // namespace

void VoxelFilter::init()
{
    desc_.ensure(1);
    voxel_desc_init_kernel<<<1, 1>>>(desc_.p);
    LLB_CUDA(cudaGetLastError());
}

void VoxelFilter::release()
{
    desc_.release(); keys_[0].release(); keys_[1].release(); vals_[0].release(); vals_[1].release();
    hist_.release(); blk_.release();
}

int VoxelFilter::run(const VoxelInput &in, float leaf, float4 *out, int *n_out_dev, cudaStream_t stream)
{
    return run_batch(&in, &leaf, &out, &n_out_dev, 1, stream);
}

int VoxelFilter::run_batch(const VoxelInput *in, const float *leaf, float4 *const *out, int *const *n_out_dev,
                           int count, cudaStream_t s)
{
    if (count < 1 || count > MAX_BATCH) throw std::invalid_argument("voxel batch size");
    bool all_small = true;
    int upper = 0;
    for (int j = 0; j < count; j++) {
        all_small &= in[j].upper() <= SMALL_MAX;
        upper = std::max(upper, in[j].upper());
    }
    if (!all_small) {
        int launches = 0;
        for (int j = 0; j < count; j++) {
            if (in[j].upper() <= SMALL_MAX) launches += run_batch(&in[j], &leaf[j], &out[j], &n_out_dev[j], 1, s);
            else launches += run_large(in[j], leaf[j], out[j], n_out_dev[j], s);
        }
        return launches;
    }
    int P = 32;
    while (P < upper) P <<= 1;
    size_t smem = (size_t)P * sizeof(unsigned long long);
    if (!small_attr_set_) {
        LLB_CUDA(cudaFuncSetAttribute(voxel_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      SMALL_MAX * (int)sizeof(unsigned long long)));
        small_attr_set_ = true;
    }
    SmallJobs jobs{};
    for (int j = 0; j < count; j++) {
        jobs.in[j] = to_seg(in[j]); jobs.leaf[j] = leaf[j]; jobs.out[j] = out[j]; jobs.n_out[j] = n_out_dev[j];
    }
    voxel_small_kernel<<<count, SMALL_THREADS, smem, s>>>(jobs);
    LLB_CUDA(cudaGetLastError());
    return 1;
}

int VoxelFilter::run_large(const VoxelInput &in, float leaf, float4 *out, int *n_out_dev, cudaStream_t s)
{
    const int n = in.upper();
    keys_[0].ensure(n); keys_[1].ensure(n); vals_[0].ensure(n); vals_[1].ensure(n);
    const int nblk_radix = std::min(RADIX_MAX_BLOCKS, div_up(n, 4096));
    hist_.ensure((size_t)256 * nblk_radix);
    const int nblk_head = div_up(n, HEAD_TILE);
    blk_.ensure(nblk_head + 1);
    SegIn seg = to_seg(in);
    int launches = 0;
    const int grid_stream = std::min(div_up(n, LG_THREADS), 148 * 8);

    voxel_minmax_kernel<<<grid_stream, LG_THREADS, 0, s>>>(seg, desc_.p); launches++;
    voxel_setup_kernel<<<1, 1, 0, s>>>(seg, leaf, desc_.p); launches++;
    voxel_keys_kernel<<<grid_stream, LG_THREADS, 0, s>>>(seg, desc_.p, keys_[0].p, vals_[0].p); launches++;
    for (int pass = 0; pass < 4; pass++) {
        const int shift = pass * 8;
        radix_hist_kernel<<<nblk_radix, LG_THREADS, 0, s>>>(desc_.p, shift, keys_[0].p, keys_[1].p, hist_.p);
        scan_single_block_kernel<<<1, 1024, 0, s>>>(hist_.p, 256 * nblk_radix, desc_.p, shift, nullptr);
        radix_scatter_kernel<<<nblk_radix, LG_THREADS, 0, s>>>(desc_.p, shift, keys_[0].p, keys_[1].p,
                                                               vals_[0].p, vals_[1].p, hist_.p);
        launches += 3;
    }
    voxel_heads_kernel<<<nblk_head, HEAD_TILE, 0, s>>>(desc_.p, keys_[0].p, keys_[1].p, blk_.p); launches++;
    scan_single_block_kernel<<<1, 1024, 0, s>>>(blk_.p, nblk_head, nullptr, -1, n_out_dev); launches++;
    voxel_centroid_kernel<<<nblk_head, HEAD_TILE, 0, s>>>(seg, desc_.p, keys_[0].p, keys_[1].p, vals_[0].p,
                                                          vals_[1].p, blk_.p, out); launches++;
    LLB_CUDA(cudaGetLastError());
    return launches;
}

}  // namespace llb
