// voxel.cu — K1, the multi-kernel path for clouds that do not fit one CTA / cluster (local maps, VLS-128 scans; the
// small paths live in voxel_small.cu): min/max -> setup -> keys (+ first histogram) -> stable LSD radix sort of (key,
// input index) pairs (8-bit digits over the significant bits, per-block histograms, ranking by ballots, sub-tiles put
// into digit order in shared memory before they leave) -> per-tile head counts -> scan -> centroids (tile staged in
// shared memory, one thread per voxel, sums sequential in ascending input index).  Every kernel serves job blockIdx.y
// of a device-resident table.  Measured history: profiles/r02_voxel.md.
// The float arithmetic that decides membership is written exactly as PCL does it
// (float multiply, floorf, float subtract, int cast) and this file is compiled with
// -fmad=false, so voxel membership, order and centroids are bit-identical to the oracle.
#include "voxel_dev.cuh"
#include <cstdlib>
#include <cstring>

namespace llb {

namespace {

// ------------------------------------------------------------------ large path

constexpr int LG_THREADS = 256;
constexpr int RADIX_MAX_BLOCKS = 256;
// keys per CTA of the radix kernels (histogram layout depends on it: the same value sizes the scratch and the launch)
static int radix_keys_per_block()
{
    static const int v = getenv("LLB_RADIX_KEYS") ? std::max(2048, atoi(getenv("LLB_RADIX_KEYS"))) : 8192;
    return v;
}

// every kernel of the large path serves job blockIdx.y of a device-resident table: one entry for a single filter
// (llb_ctx), 2B entries when the map filters of a batched step share each launch
#define LG_JOB(table) const LargeVoxelJob jb = (table)[blockIdx.y]

__global__ void voxel_desc_init_kernel(VoxelDesc *d)
{
    for (int a = 0; a < 3; a++) { d->mn[a] = INT_MAX; d->mx[a] = INT_MIN; }
}

__global__ void __launch_bounds__(LG_THREADS)
voxel_minmax_kernel(const LargeVoxelJob *__restrict__ table)
{
    LG_JOB(table);
    const SegIn in = jb.bounds; VoxelDesc *__restrict__ d = jb.desc;
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

__global__ void voxel_setup_kernel(const LargeVoxelJob *__restrict__ table)
{
    LG_JOB(table);
    const SegIn in = jb.in; const float leaf = jb.leaf; VoxelDesc *d = jb.desc;
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

// voxel keys of the points + the per-block histogram of the FIRST radix digit (same tile layout as the radix kernels: the
// first pass needs no histogram launch).  The sort carries (key, input index) pairs; the index list of the first pass is
// the identity and is never written: radix_scatter_kernel takes val = position when shift == 0.
__global__ void __launch_bounds__(LG_THREADS)
voxel_keys_kernel(const LargeVoxelJob *__restrict__ table)
{
    LG_JOB(table);
    const SegIn in = jb.in; const VoxelDesc *__restrict__ d = jb.desc;
    unsigned *__restrict__ keys = jb.kA; int *__restrict__ hist = jb.hist;
    const int n = d->n;
    const int na = seg_len_a(in);
    const float inv = d->inv;
    const int min_b[3] = { d->min_b[0], d->min_b[1], d->min_b[2] }, mul[3] = { d->mul[0], d->mul[1], d->mul[2] };
    const int ovf = d->overflow;
    __shared__ int s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    int lo, hi;
    radix_tile(n, gridDim.x, blockIdx.x, lo, hi);
    for (int i = lo + threadIdx.x; i < hi; i += LG_THREADS) {
        float4 p = seg_load(in, na, i);
        const unsigned key = ovf ? (unsigned)i : voxel_key(p, inv, min_b, mul);
        keys[i] = key;
        atomicAdd(&s_h[key & 255u], 1);
    }
    __syncthreads();
    hist[threadIdx.x * gridDim.x + blockIdx.x] = s_h[threadIdx.x];
}

__global__ void __launch_bounds__(LG_THREADS)
radix_hist_kernel(const LargeVoxelJob *__restrict__ table, int shift)
{
    LG_JOB(table);
    const VoxelDesc *__restrict__ d = jb.desc; const unsigned *__restrict__ kA = jb.kA, *__restrict__ kB = jb.kB;
    int *__restrict__ hist = jb.hist;
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
scan_single_block_kernel(const LargeVoxelJob *__restrict__ table, int what, int count, int shift)
{
    // what 0: the radix histograms of pass `shift`; what 1: the per-block head counts (total -> output length)
    LG_JOB(table);
    int *__restrict__ data = what == 0 ? jb.hist : jb.blk;
    const VoxelDesc *d = what == 0 ? jb.desc : nullptr;
    int *total_out = what == 0 ? nullptr : jb.n_out;
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
radix_scatter_kernel(const LargeVoxelJob *__restrict__ table, int shift)
{
    LG_JOB(table);
    const VoxelDesc *__restrict__ d = jb.desc; unsigned *__restrict__ kA = jb.kA, *__restrict__ kB = jb.kB;
    int *__restrict__ vA = jb.vA, *__restrict__ vB = jb.vB; const int *__restrict__ hist = jb.hist;
    if (shift >= d->nbits) return;
    const bool odd = ((shift >> 3) & 1) != 0;
    const unsigned *kin = odd ? kB : kA; unsigned *kout = odd ? kA : kB;
    const int *vin = odd ? vB : vA; int *vout = odd ? vA : vB;

    constexpr int NW = LG_THREADS / 32;
    constexpr int ITEMS = 8, SUB = LG_THREADS * ITEMS;
    __shared__ int s_base[256];              // global position of the next key of digit d
    __shared__ int s_wcnt[NW][256];          // per-warp digit counters -> per-warp offsets inside the digit's run of the sub-tile
    __shared__ int s_lstart[257];            // start of digit d's run inside the sorted sub-tile
    __shared__ int s_gdelta[256];            // global position of that run minus its local start
    __shared__ int s_scan[33];
    __shared__ unsigned s_k[SUB];            // the sub-tile in sorted order: each digit's run leaves as consecutive,
    __shared__ int s_v[SUB];                 // coalesced stores instead of one 4-byte store per lane into up to 32 sectors
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    s_base[tid] = hist[tid * gridDim.x + blockIdx.x];
    int lo, hi;
    radix_tile(d->n, gridDim.x, blockIdx.x, lo, hi);
    // Sub-tiles of 2048 keys: warp w owns the contiguous keys [w*256, (w+1)*256) of the sub-tile and ranks them
    // in 8 rounds of 32 against its PRIVATE digit counters (warp-level sync only).  Order of ranks = (sub-tile, warp,
    // round, lane) = input order: stable.
    const unsigned lt = (1u << lane) - 1u;
    for (int sub = lo; sub < hi; sub += SUB) {
#pragma unroll
        for (int k = 0; k < NW; k++) s_wcnt[k][tid] = 0;
        __syncthreads();
        unsigned key[ITEMS]; int val[ITEMS]; int rk[ITEMS]; unsigned dg[ITEMS];
#pragma unroll
        for (int r = 0; r < ITEMS; r++) {
            const int i = sub + w * (32 * ITEMS) + r * 32 + lane;
            const bool valid = i < hi;
            key[r] = valid ? kin[i] : 0u;
            val[r] = valid ? (shift == 0 ? i : vin[i]) : 0;                 // first pass: the index list is the identity
            dg[r] = valid ? ((key[r] >> shift) & 255u) : (256u + lane);    // invalid lanes never match
        }
#pragma unroll
        for (int r = 0; r < ITEMS; r++) {
            const unsigned m = match_low_bits<9>(dg[r]);      // digit, or 256 + lane for the lanes past the tile
            const int pr = __popc(m & lt);
            int cnt = 0;
            if (dg[r] < 256u) cnt = s_wcnt[w][dg[r]];
            rk[r] = cnt + pr;
            __syncwarp();
            if (dg[r] < 256u && pr == 0) s_wcnt[w][dg[r]] = cnt + __popc(m);
            __syncwarp();
        }
        __syncthreads();
        int dtot = 0;
        {   // thread tid owns digit tid: per-warp counts -> offsets inside the digit's run; dtot = keys of that digit
#pragma unroll
            for (int k = 0; k < NW; k++) { const int c = s_wcnt[k][tid]; s_wcnt[k][tid] = dtot; dtot += c; }
        }
        int stot;
        const int lstart = block_excl_scan(dtot, s_scan, stot);          // (barriers inside)
        s_lstart[tid] = lstart;
        s_gdelta[tid] = s_base[tid] - lstart;
        s_base[tid] += dtot;
        if (tid == 0) s_lstart[256] = stot;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < ITEMS; r++) {
            if (dg[r] < 256u) {
                const int lp = s_lstart[dg[r]] + s_wcnt[w][dg[r]] + rk[r];
                s_k[lp] = key[r]; s_v[lp] = val[r];
            }
        }
        __syncthreads();
        for (int e = tid; e < stot; e += LG_THREADS) {
            const unsigned kk = s_k[e];
            const int pos = s_gdelta[(kk >> shift) & 255u] + e;
            kout[pos] = kk; vout[pos] = s_v[e];
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

// heads + centroids work on tiles of HEAD_TILE sorted positions: HEAD_THREADS threads, HEAD_ITEMS positions each
// (position r * HEAD_THREADS + t of the tile: coalesced, and all loads of a thread are in flight together)
constexpr int HEAD_THREADS = 256;
constexpr int HEAD_ITEMS = 4;
constexpr int HEAD_TILE = HEAD_THREADS * HEAD_ITEMS;
constexpr int HEAD_GROUP = 2;           // tiles per CTA of the head count (the per-tile counts feed the centroid kernel)

__global__ void __launch_bounds__(HEAD_THREADS)
voxel_heads_kernel(const LargeVoxelJob *__restrict__ table, int nblk)
{
    LG_JOB(table);
    const VoxelDesc *__restrict__ d = jb.desc; const unsigned *kA = jb.kA, *kB = jb.kB; int *__restrict__ blk = jb.blk;
    const unsigned *k; const int *v;
    sorted_bufs(d, kA, kB, nullptr, nullptr, k, v);
    const int n = d->n;
#pragma unroll 1
    for (int g = 0; g < HEAD_GROUP; g++) {
        const int tile = blockIdx.x * HEAD_GROUP + g;
        if (tile >= nblk) break;                             // uniform over the CTA; tiles past n count 0 (the scan reads all nblk)
        int cnt = 0;
#pragma unroll
        for (int r = 0; r < HEAD_ITEMS; r++) {
            const int i = tile * HEAD_TILE + r * HEAD_THREADS + threadIdx.x;
            const int head = (i < n) && (i == 0 || k[i - 1] != k[i]);
            cnt += __syncthreads_count(head);
        }
        if (threadIdx.x == 0) blk[tile] = cnt;
    }
}

// One CTA per tile of sorted positions: keys and points of the tile are staged in shared memory (the points are fetched
// through the sorted index list: the gather is part of this kernel, nothing is written in between), heads are ranked
// with the per-tile counts of voxel_heads_kernel, and the thread of every head walks its voxel IN shared memory (the
// sums stay sequential in ascending input index, as PCL adds them); a voxel that runs past the tile is finished from
// global memory.  The walk used to chase one dependent global load per point.
__global__ void __launch_bounds__(HEAD_THREADS)
voxel_centroid_kernel(const LargeVoxelJob *__restrict__ table)
{
    LG_JOB(table);
    const VoxelDesc *__restrict__ d = jb.desc;
    const int n = d->n;
    const int i0 = blockIdx.x * HEAD_TILE;
    if (i0 >= n) return;                                     // the grid is sized for the longest job of the table
    const int *__restrict__ blk = jb.blk; float4 *__restrict__ out = jb.out;
    const SegIn in = jb.in;
    const int na = seg_len_a(in);
    __shared__ int s_scan[33];
    __shared__ unsigned s_key[HEAD_TILE];
    __shared__ float4 s_pt[HEAD_TILE];
    __shared__ int s_hp[HEAD_TILE];
    const unsigned *k; const int *v;
    sorted_bufs(d, jb.kA, jb.kB, jb.vA, jb.vB, k, v);
    const int t = threadIdx.x;
    const int tn = min(HEAD_TILE, n - i0);
    unsigned key[HEAD_ITEMS]; int src[HEAD_ITEMS];
#pragma unroll
    for (int r = 0; r < HEAD_ITEMS; r++) {
        const int p = r * HEAD_THREADS + t;
        key[r] = p < tn ? k[i0 + p] : 0u;
        src[r] = p < tn ? __ldg(&v[i0 + p]) : 0;
    }
    const unsigned kprev = i0 > 0 ? k[i0 - 1] : 0u;
#pragma unroll
    for (int r = 0; r < HEAD_ITEMS; r++) {
        const int p = r * HEAD_THREADS + t;
        if (p < tn) { s_key[p] = key[r]; s_pt[p] = seg_load(in, na, src[r]); }
    }
    __syncthreads();
    // heads of the tile in position order -> s_hp: the walks below then run on CONSECUTIVE threads (one per voxel)
    // instead of on the scattered head lanes of every warp (a voxel has ~4.5 points: 7 of 32 lanes were active)
    int nheads = 0;
#pragma unroll 1
    for (int r = 0; r < HEAD_ITEMS; r++) {
        const int p = r * HEAD_THREADS + t;
        const unsigned kk = p < tn ? s_key[p] : 0u;
        int head = 0;
        if (p < tn) head = (p == 0) ? (i0 == 0 || kprev != kk) : (s_key[p - 1] != kk);
        int total;
        const int lr = nheads + block_excl_scan(head, s_scan, total);
        nheads += total;
        if (head) s_hp[lr] = p;
    }
    __syncthreads();
    const int base = blk[blockIdx.x];
    for (int q = t; q < nheads; q += HEAD_THREADS) {
        const int p = s_hp[q];
        const unsigned kk = s_key[p];
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
        int j = p;
        // the keys are sorted: if the 4th key ahead still belongs to this voxel, so do the three before it.  Four
        // points per trip: the loads are independent, only the four float chains stay sequential (PCL's order)
        while (j + 4 <= tn && s_key[j + 3] == kk) {
            const float4 p0 = s_pt[j], p1 = s_pt[j + 1], p2 = s_pt[j + 2], p3 = s_pt[j + 3];
            sx += p0.x; sy += p0.y; sz += p0.z; si += p0.w;
            sx += p1.x; sy += p1.y; sz += p1.z; si += p1.w;
            sx += p2.x; sy += p2.y; sz += p2.z; si += p2.w;
            sx += p3.x; sy += p3.y; sz += p3.z; si += p3.w;
            j += 4;
        }
        while (j < tn && s_key[j] == kk) {
            const float4 pq = s_pt[j];
            sx += pq.x; sy += pq.y; sz += pq.z; si += pq.w;
            j++;
        }
        int cntp = j - p;
        if (j == HEAD_TILE) {                                // the voxel continues in the next tile(s)
            int g = i0 + HEAD_TILE;
            while (g < n && k[g] == kk) {
                const float4 pq = seg_load(in, na, v[g]);
                sx += pq.x; sy += pq.y; sz += pq.z; si += pq.w;
                g++; cntp++;
            }
        }
        const float cnt = (float)cntp;
        out[base + q] = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
    }
}

}  // namespace

void VoxelFilter::init()
{
    desc_.ensure(1);
    voxel_desc_init_kernel<<<1, 1>>>(desc_.p);
    LLB_CUDA(cudaGetLastError());
}

void VoxelFilter::reserve(int n)
{
    if (n <= SMALL_MAX) return;                              // the small paths need no global scratch
    keys_[0].ensure(n); keys_[1].ensure(n); vals_[0].ensure(n); vals_[1].ensure(n);
    hist_.ensure((size_t)256 * std::min(RADIX_MAX_BLOCKS, div_up(n, 2048)));
    blk_.ensure(div_up(n, HEAD_TILE) + 1);
}

void VoxelFilter::release()
{
    desc_.release(); keys_[0].release(); keys_[1].release(); vals_[0].release(); vals_[1].release();
    hist_.release(); blk_.release(); job_raw_.release(); job_pin_.release();
    if (job_ev_) { cudaEventDestroy(job_ev_); job_ev_ = nullptr; }
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
    SmallJobs jobs{};
    for (int j = 0; j < count; j++) {
        jobs.in[j] = to_seg(in[j]); jobs.leaf[j] = leaf[j]; jobs.out[j] = out[j]; jobs.n_out[j] = n_out_dev[j];
    }
    launch_voxel_small(jobs, count, s);
    return 1;
}

LargeVoxelJob VoxelFilter::large_job(const VoxelInput &in, float leaf, float4 *out, int *n_out_dev, const VoxelInput *bounds)
{
    const int n = std::max(in.upper(), 1);
    reserve(std::max(n, SMALL_MAX + 1));
    LargeVoxelJob j;
    j.in = to_seg(in); j.leaf = leaf; j.desc = desc_.p; j.bounds = bounds ? to_seg(*bounds) : j.in;
    j.kA = keys_[0].p; j.kB = keys_[1].p; j.vA = vals_[0].p; j.vB = vals_[1].p;
    j.hist = hist_.p; j.blk = blk_.p; j.out = out; j.n_out = n_out_dev;
    return j;
}

// `count` jobs of a device-resident table; n_upper bounds the input length of every job
int VoxelFilter::launch_large(const LargeVoxelJob *table_dev, int count, int n_upper, cudaStream_t s)
{
    const int n = std::max(n_upper, 1);
    // the histogram layout depends on the radix grid: it must be the same for sizing (reserve) and launching
    const int nblk_radix = std::min(RADIX_MAX_BLOCKS, div_up(n, radix_keys_per_block()));
    const int nblk_head = div_up(n, HEAD_TILE);
    const unsigned ny = (unsigned)std::max(count, 1);
    const int per = count > 1 ? std::max(8, 148 * 8 / count) : 148 * 8;
    const dim3 grid_stream(std::min(div_up(n, LG_THREADS), per), ny);
    int launches = 0;
    voxel_minmax_kernel<<<grid_stream, LG_THREADS, 0, s>>>(table_dev); launches++;
    voxel_setup_kernel<<<dim3(1, ny), 1, 0, s>>>(table_dev); launches++;
    voxel_keys_kernel<<<dim3(nblk_radix, ny), LG_THREADS, 0, s>>>(table_dev); launches++;     // + histogram of pass 0
    for (int pass = 0; pass < 4; pass++) {
        const int shift = pass * 8;
        if (pass > 0) { radix_hist_kernel<<<dim3(nblk_radix, ny), LG_THREADS, 0, s>>>(table_dev, shift); launches++; }
        scan_single_block_kernel<<<dim3(1, ny), 1024, 0, s>>>(table_dev, 0, 256 * nblk_radix, shift);
        radix_scatter_kernel<<<dim3(nblk_radix, ny), LG_THREADS, 0, s>>>(table_dev, shift);
        launches += 2;
    }
    voxel_heads_kernel<<<dim3(div_up(nblk_head, HEAD_GROUP), ny), HEAD_THREADS, 0, s>>>(table_dev, nblk_head); launches++;
    scan_single_block_kernel<<<dim3(1, ny), 1024, 0, s>>>(table_dev, 1, nblk_head, -1); launches++;
    voxel_centroid_kernel<<<dim3(nblk_head, ny), HEAD_THREADS, 0, s>>>(table_dev); launches++;
    LLB_CUDA(cudaGetLastError());
    return launches;
}

int VoxelFilter::run_large(const VoxelInput &in, float leaf, float4 *out, int *n_out_dev, cudaStream_t s, const VoxelInput *bounds)
{
    const LargeVoxelJob j = large_job(in, leaf, out, n_out_dev, bounds);
    job_raw_.ensure(sizeof(LargeVoxelJob));
    job_pin_.ensure(sizeof(LargeVoxelJob));
    if (!job_ev_) LLB_CUDA(cudaEventCreateWithFlags(&job_ev_, cudaEventDisableTiming));
    else LLB_CUDA(cudaEventSynchronize(job_ev_));            // the previous upload has left the pinned record (long ago)
    std::memcpy(job_pin_.p, &j, sizeof(j));
    LLB_CUDA(cudaMemcpyAsync(job_raw_.p, job_pin_.p, sizeof(j), cudaMemcpyHostToDevice, s));
    LLB_CUDA(cudaEventRecord(job_ev_, s));
    return launch_large(reinterpret_cast<const LargeVoxelJob *>(job_raw_.p), 1, std::max(in.upper(), bounds ? bounds->upper() : 0), s);
}

}  // namespace llb
