// grid_index.cu — K2 build kernels.  Both maps of a registration (corner, surf) are built by
// the SAME five launches (blockIdx.y selects the map):
//   1 bounds           (+ the last CTA to finish derives the grid: origin, cell size, dims)
//   2 per-cell counts  (one atomic per point; the value it returns is the point's slot in its cell)
//   3 tile sums of the dense count table (+ the last CTA scans the tile sums)
//   4 exclusive scan of the count table -> cell_begin
//   5 scatter into cell order (+ re-zeroes the count table for the next build)
// Everything is sized on the device; the host only knows upper bounds, so nothing synchronises.
#include "grid_index.cuh"

namespace llb {

namespace {

constexpr int TPB = 256;
constexpr int SCAN_TILE = 4096;     // 1024 threads x 4

struct GridJob {
    const float4 *pts; const int *n_dev; int n_host;
    GridDesc *desc; int *counts; int *cell_begin; int *cell_of; int *rank; int *blk; float4 *sorted;
};
struct GridJobs { GridJob j[2]; float radius; int max_cells; };

__global__ void grid_desc_init_kernel(GridDesc *d)
{
    for (int a = 0; a < 3; a++) { d->mn[a] = INT_MAX; d->mx[a] = INT_MIN; }
    d->ncell = 0; d->n = 0; d->ticket = 0;
}

__device__ void grid_setup(GridDesc *d, int n, float radius, int max_cells)
{
    float mn[3], mx[3];
    for (int a = 0; a < 3; a++) {
        mn[a] = ordered_to_float(d->mn[a]); mx[a] = ordered_to_float(d->mx[a]);
        d->mn[a] = INT_MAX; d->mx[a] = INT_MIN;               // ready for the next build
    }
    d->n = n;
    if (n <= 0) {
        d->ncell = 1; d->dim[0] = d->dim[1] = d->dim[2] = 1;
        d->org[0] = d->org[1] = d->org[2] = 0.f; d->cell = radius; d->inv_cell = 1.0f / radius;
        return;
    }
    float cell = radius * 1.001f;
    for (int it = 0; it < 64; it++) {
        float inv = 1.0f / cell;
        long long tot = 1;
        int dim[3];
        for (int a = 0; a < 3; a++) {
            dim[a] = (int)floorf((mx[a] - mn[a]) * inv) + 1;
            tot *= dim[a];
        }
        if (tot <= (long long)max_cells) {
            d->cell = cell; d->inv_cell = inv;
            for (int a = 0; a < 3; a++) { d->dim[a] = dim[a]; d->org[a] = mn[a]; }
            d->ncell = (int)tot;
            return;
        }
        cell *= 1.26f;                                       // ~ doubles the cell volume
    }
    d->cell = 3.0e38f; d->inv_cell = 0.f; d->ncell = 1;      // unreachable for finite bounds: one cell
    for (int a = 0; a < 3; a++) { d->dim[a] = 1; d->org[a] = mn[a]; }
}

__global__ void __launch_bounds__(TPB)
grid_bbox_kernel(GridJobs jobs)
{
    const GridJob &jb = jobs.j[blockIdx.y];
    GridDesc *d = jb.desc;
    __shared__ float s_red[6][TPB / 32];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int n = jb.n_dev ? *jb.n_dev : jb.n_host;
    float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int i = blockIdx.x * TPB + tid; i < n; i += gridDim.x * TPB) {
        float4 p = __ldg(&jb.pts[i]);
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
    if (tid < 3 && n > 0) {
        float m0 = s_red[tid][0], m1 = s_red[3 + tid][0];
        for (int k = 1; k < TPB / 32; k++) { m0 = fminf(m0, s_red[tid][k]); m1 = fmaxf(m1, s_red[3 + tid][k]); }
        atomicMin(&d->mn[tid], float_to_ordered(m0));
        atomicMax(&d->mx[tid], float_to_ordered(m1));
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&d->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last && tid == 0) {
        __threadfence();
        d->ticket = 0;
        grid_setup(d, n, jobs.radius, jobs.max_cells);
    }
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void __launch_bounds__(TPB)
grid_count_kernel(GridJobs jobs)
{
    const GridJob &jb = jobs.j[blockIdx.y];
    const GridDesc *d = jb.desc;
    const int n = d->n;
    const float ox = d->org[0], oy = d->org[1], oz = d->org[2], inv = d->inv_cell;
    const int dx = d->dim[0], dy = d->dim[1], dz = d->dim[2];
    for (int i = blockIdx.x * TPB + threadIdx.x; i < n; i += gridDim.x * TPB) {
        float4 p = __ldg(&jb.pts[i]);
        int cx = clampi(grid_coord(p.x, ox, inv), 0, dx - 1);
        int cy = clampi(grid_coord(p.y, oy, inv), 0, dy - 1);
        int cz = clampi(grid_coord(p.z, oz, inv), 0, dz - 1);
        int c = (cz * dy + cy) * dx + cx;
        jb.cell_of[i] = c;
        jb.rank[i] = atomicAdd(&jb.counts[c], 1);
    }
}

// tile sums of counts[0 .. ncell]; the last CTA turns them into exclusive offsets
__global__ void __launch_bounds__(1024)
scan_tile_sum_kernel(GridJobs jobs)
{
    const GridJob &jb = jobs.j[blockIdx.y];
    GridDesc *d = jb.desc;
    __shared__ int s_w[33];
    __shared__ int s_last;
    const int m = d->ncell + 1;
    const int ntiles = (m + SCAN_TILE - 1) / SCAN_TILE;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {     // grid-stride: the host does not know ncell
        const int base = tile * SCAN_TILE;
        int s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int i = base + k * 1024 + threadIdx.x;
            if (i < m) s += jb.counts[i];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            int v = s_w[threadIdx.x];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
            if (threadIdx.x == 0) jb.blk[tile] = v;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(&d->ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) d->ticket = 0;
    // exclusive scan of the tile sums by this CTA
    const int count = ntiles;
    const int per = (count + 1023) / 1024;
    const int lo = min((int)threadIdx.x * per, count), hi = min(lo + per, count);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += __ldcg(&jb.blk[i]);
    int total;
    int ex = block_excl_scan(sum, s_w, total);
    for (int i = lo; i < hi; i++) { int v = __ldcg(&jb.blk[i]); jb.blk[i] = ex; ex += v; }
}

__global__ void __launch_bounds__(1024)
scan_tile_apply_kernel(GridJobs jobs)
{
    const GridJob &jb = jobs.j[blockIdx.y];
    __shared__ int s_scan[33];
    const int m = jb.desc->ncell + 1;
    const int ntiles = (m + SCAN_TILE - 1) / SCAN_TILE;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i0 = tile * SCAN_TILE + threadIdx.x * 4;   // thread owns 4 CONSECUTIVE entries
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = (i0 + k < m) ? jb.counts[i0 + k] : 0;
        int total;
        int ex = jb.blk[tile] + block_excl_scan(v[0] + v[1] + v[2] + v[3], s_scan, total);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < m) jb.cell_begin[i0 + k] = ex;
            ex += v[k];
        }
    }
}

__global__ void __launch_bounds__(TPB)
grid_scatter_kernel(GridJobs jobs)
{
    const GridJob &jb = jobs.j[blockIdx.y];
    const GridDesc *d = jb.desc;
    const int n = d->n;
    for (int i = blockIdx.x * TPB + threadIdx.x; i < n; i += gridDim.x * TPB) {
        float4 p = __ldg(&jb.pts[i]);
        const int c = jb.cell_of[i];
        jb.sorted[jb.cell_begin[c] + jb.rank[i]] = make_float4(p.x, p.y, p.z, __int_as_float(i));
        jb.counts[c] = 0;                                    // leave the count table clean for the next build
    }
}

__global__ void grid_zero_kernel(int *p, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0;
}

}  // namespace

void GridIndex::init(int max_cells)
{
    max_cells_ = max_cells > 4096 ? max_cells : 4096;
    desc_.ensure(1);
    counts_.ensure((size_t)max_cells_ + 1);
    cell_begin_.ensure((size_t)max_cells_ + 1);
    blk_.ensure((size_t)div_up(max_cells_ + 1, SCAN_TILE) + 1);
    grid_desc_init_kernel<<<1, 1>>>(desc_.p);
    grid_zero_kernel<<<148 * 4, 256>>>(counts_.p, max_cells_ + 1);
    LLB_CUDA(cudaGetLastError());
}

void GridIndex::release()
{
    desc_.release(); sorted_.release(); counts_.release(); cell_begin_.release(); cell_of_.release(); rank_.release();
    blk_.release();
}

int GridIndex::build_pair(GridIndex &a, const float4 *pa, const int *na_dev, int na_upper,
                          GridIndex &b, const float4 *pb, const int *nb_dev, int nb_upper, float radius, cudaStream_t s)
{
    GridIndex *g[2] = { &a, &b };
    const float4 *pts[2] = { pa, pb };
    const int *ndev[2] = { na_dev, nb_dev };
    const int nup[2] = { na_upper, nb_upper };
    GridJobs jobs;
    jobs.radius = radius;
    jobs.max_cells = std::min(a.max_cells_, b.max_cells_);
    int nmax = 1;
    for (int k = 0; k < 2; k++) {
        const int n = nup[k] > 0 ? nup[k] : 1;
        nmax = std::max(nmax, n);
        g[k]->sorted_.ensure(n); g[k]->cell_of_.ensure(n); g[k]->rank_.ensure(n);
        GridJob &j = jobs.j[k];
        j.pts = pts[k]; j.n_dev = ndev[k]; j.n_host = nup[k];
        j.desc = g[k]->desc_.p; j.counts = g[k]->counts_.p; j.cell_begin = g[k]->cell_begin_.p;
        j.cell_of = g[k]->cell_of_.p; j.rank = g[k]->rank_.p; j.blk = g[k]->blk_.p; j.sorted = g[k]->sorted_.p;
    }
    const dim3 grid_pts(std::min(div_up(nmax, TPB), 148 * 4), 2);
    const dim3 grid_scan(std::min(div_up(jobs.max_cells + 1, SCAN_TILE), 148 * 2), 2);
    grid_bbox_kernel<<<grid_pts, TPB, 0, s>>>(jobs);
    grid_count_kernel<<<grid_pts, TPB, 0, s>>>(jobs);
    scan_tile_sum_kernel<<<grid_scan, 1024, 0, s>>>(jobs);
    scan_tile_apply_kernel<<<grid_scan, 1024, 0, s>>>(jobs);
    grid_scatter_kernel<<<grid_pts, TPB, 0, s>>>(jobs);
    LLB_CUDA(cudaGetLastError());
    return 5;
}

}  // namespace llb
