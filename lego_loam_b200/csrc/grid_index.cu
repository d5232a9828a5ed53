// grid_index.cu — K2 build kernels: bounds -> grid set-up -> per-cell counts (one atomic
// per point, the returned value is the point's slot inside its cell) -> exclusive scan of
// the dense cell table -> scatter into cell order.  Everything is sized on the device;
// the host only knows upper bounds, so nothing synchronises.
#include "grid_index.cuh"

namespace llb {

namespace {

constexpr int TPB = 256;
constexpr int SCAN_TILE = 4096;     // 1024 threads x 4

__global__ void grid_desc_init_kernel(GridDesc *d)
{
    for (int a = 0; a < 3; a++) { d->mn[a] = INT_MAX; d->mx[a] = INT_MIN; }
    d->ncell = 0; d->n = 0;
}

__global__ void __launch_bounds__(TPB)
grid_bbox_kernel(const float4 *__restrict__ pts, const int *n_dev, int n_host, GridDesc *__restrict__ d)
{
    __shared__ float s_red[6][TPB / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int n = n_dev ? *n_dev : n_host;
    float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int i = blockIdx.x * TPB + tid; i < n; i += gridDim.x * TPB) {
        float4 p = __ldg(&pts[i]);
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
}

__global__ void grid_setup_kernel(GridDesc *d, const int *n_dev, int n_host, float radius, int max_cells)
{
    const int n = n_dev ? *n_dev : n_host;
    float mn[3], mx[3];
    for (int a = 0; a < 3; a++) {
        mn[a] = ordered_to_float(d->mn[a]); mx[a] = ordered_to_float(d->mx[a]);
        d->mn[a] = INT_MAX; d->mx[a] = INT_MIN;
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
        cell *= 1.26f;                                   // ~ doubles the cell volume
    }
    // unreachable for finite bounds; fall back to one cell (brute force)
    d->cell = 3.0e38f; d->inv_cell = 0.f; d->ncell = 1;
    for (int a = 0; a < 3; a++) { d->dim[a] = 1; d->org[a] = mn[a]; }
}

__global__ void __launch_bounds__(TPB)
grid_clear_kernel(const GridDesc *__restrict__ d, int *__restrict__ table)
{
    const int m = d->ncell + 1;
    for (int i = blockIdx.x * TPB + threadIdx.x; i < m; i += gridDim.x * TPB) table[i] = 0;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void __launch_bounds__(TPB)
grid_count_kernel(const float4 *__restrict__ pts, const GridDesc *__restrict__ d, int *__restrict__ counts,
                  int *__restrict__ cell_of, int *__restrict__ rank)
{
    const int n = d->n;
    const float ox = d->org[0], oy = d->org[1], oz = d->org[2], inv = d->inv_cell;
    const int dx = d->dim[0], dy = d->dim[1], dz = d->dim[2];
    for (int i = blockIdx.x * TPB + threadIdx.x; i < n; i += gridDim.x * TPB) {
        float4 p = __ldg(&pts[i]);
        int cx = clampi(grid_coord(p.x, ox, inv), 0, dx - 1);
        int cy = clampi(grid_coord(p.y, oy, inv), 0, dy - 1);
        int cz = clampi(grid_coord(p.z, oz, inv), 0, dz - 1);
        int c = (cz * dy + cy) * dx + cx;
        cell_of[i] = c;
        rank[i] = atomicAdd(&counts[c], 1);
    }
}

__global__ void __launch_bounds__(1024)
scan_tile_sum_kernel(const GridDesc *__restrict__ d, const int *__restrict__ data, int *__restrict__ blk)
{
    const int m = d->ncell + 1;
    const int base = blockIdx.x * SCAN_TILE;
    if (base >= m) {
        if (threadIdx.x == 0) blk[blockIdx.x] = 0;
        return;
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int i = base + k * 1024 + threadIdx.x;
        if (i < m) s += data[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    __shared__ int s_w[32];
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = s_w[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (threadIdx.x == 0) blk[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(1024)
scan_blocks_kernel(int *__restrict__ blk, int count)
{
    __shared__ int s_scan[33];
    const int per = (count + 1023) / 1024;
    const int lo = min((int)threadIdx.x * per, count), hi = min(lo + per, count);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += blk[i];
    int total;
    int base = block_excl_scan(sum, s_scan, total);
    for (int i = lo; i < hi; i++) { int v = blk[i]; blk[i] = base; base += v; }
}

__global__ void __launch_bounds__(1024)
scan_tile_apply_kernel(const GridDesc *__restrict__ d, int *__restrict__ data, const int *__restrict__ blk)
{
    __shared__ int s_scan[33];
    const int m = d->ncell + 1;
    const int base = blockIdx.x * SCAN_TILE;
    if (base >= m) return;
    // thread owns 4 CONSECUTIVE entries
    const int i0 = base + threadIdx.x * 4;
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (i0 + k < m) ? data[i0 + k] : 0;
    int total;
    int ex = blk[blockIdx.x] + block_excl_scan(v[0] + v[1] + v[2] + v[3], s_scan, total);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (i0 + k < m) data[i0 + k] = ex;
        ex += v[k];
    }
}

__global__ void __launch_bounds__(TPB)
grid_scatter_kernel(const float4 *__restrict__ pts, const GridDesc *__restrict__ d, const int *__restrict__ cell_begin,
                    const int *__restrict__ cell_of, const int *__restrict__ rank, float4 *__restrict__ sorted)
{
    const int n = d->n;
    for (int i = blockIdx.x * TPB + threadIdx.x; i < n; i += gridDim.x * TPB) {
        float4 p = __ldg(&pts[i]);
        int pos = cell_begin[cell_of[i]] + rank[i];
        sorted[pos] = make_float4(p.x, p.y, p.z, __int_as_float(i));
    }
}

}  // namespace

void GridIndex::init(int max_cells)
{
    max_cells_ = max_cells > 4096 ? max_cells : 4096;
    desc_.ensure(1);
    cell_begin_.ensure((size_t)max_cells_ + 1);
    blk_.ensure((size_t)div_up(max_cells_ + 1, SCAN_TILE) + 1);
    grid_desc_init_kernel<<<1, 1>>>(desc_.p);
    LLB_CUDA(cudaGetLastError());
}

void GridIndex::release()
{
    desc_.release(); sorted_.release(); cell_begin_.release(); cell_of_.release(); rank_.release(); blk_.release();
}

int GridIndex::build(const float4 *pts, const int *n_dev, int n_upper, float radius, cudaStream_t s)
{
    const int n = n_upper > 0 ? n_upper : 1;
    sorted_.ensure(n); cell_of_.ensure(n); rank_.ensure(n);
    const int grid_pts = std::min(div_up(n, TPB), 148 * 8);
    const int nblk_scan = div_up(max_cells_ + 1, SCAN_TILE);
    grid_bbox_kernel<<<grid_pts, TPB, 0, s>>>(pts, n_dev, n_upper, desc_.p);
    grid_setup_kernel<<<1, 1, 0, s>>>(desc_.p, n_dev, n_upper, radius, max_cells_);
    grid_clear_kernel<<<148 * 8, TPB, 0, s>>>(desc_.p, cell_begin_.p);
    grid_count_kernel<<<grid_pts, TPB, 0, s>>>(pts, desc_.p, cell_begin_.p, cell_of_.p, rank_.p);
    scan_tile_sum_kernel<<<nblk_scan, 1024, 0, s>>>(desc_.p, cell_begin_.p, blk_.p);
    scan_blocks_kernel<<<1, 1024, 0, s>>>(blk_.p, nblk_scan);
    scan_tile_apply_kernel<<<nblk_scan, 1024, 0, s>>>(desc_.p, cell_begin_.p, blk_.p);
    grid_scatter_kernel<<<grid_pts, TPB, 0, s>>>(pts, desc_.p, cell_begin_.p, cell_of_.p, rank_.p, sorted_.p);
    LLB_CUDA(cudaGetLastError());
    return 8;
}

}  // namespace llb
