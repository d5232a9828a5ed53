// grid_index.cu — K2 build kernels.  Both maps of a registration (corner, surf) are built by
// the SAME five launches (blockIdx.y selects the map):
//   1 bounds           (+ the last CTA to finish derives the grid: origin, cell size, dims)
//   2 per-cell counts  (one atomic per point; the value it returns is the point's slot in its cell)
//   3 exclusive scan of the per-ROW point counts (a row = the dimx cells of one (y,z); one CTA per map)
//   4 cell_begin of the OCCUPIED rows (one warp per row: warp scan of the row's cell counts); empty rows are
//     skipped - queries consult row_begin first - so the dense tables are only touched where points are
//     (a 100k-point local map occupies ~1 % of its ~1M cells, profiles/r01c_batch.md)
//   5 scatter into cell order (+ re-zeroes the touched cell / row counts for the next build)
// Everything is sized on the device; the host only knows upper bounds, so nothing synchronises.
#include "grid_index.cuh"

namespace llb {

namespace {

constexpr int TPB = 256;

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
grid_bbox_kernel(GridJobs jobs, const GridJob *__restrict__ table)
{
    const GridJob &jb = table ? table[blockIdx.y] : jobs.j[blockIdx.y];
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
grid_count_kernel(GridJobs jobs, const GridJob *__restrict__ table)
{
    const GridJob &jb = table ? table[blockIdx.y] : jobs.j[blockIdx.y];
    const GridDesc *d = jb.desc;
    const int n = d->n;
    const float ox = d->org[0], oy = d->org[1], oz = d->org[2], inv = d->inv_cell;
    const int dx = d->dim[0], dy = d->dim[1], dz = d->dim[2];
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    constexpr int U = 4;                                     // independent load -> atomic chains per thread
    for (int i0 = blockIdx.x * TPB * U; i0 < n; i0 += gridDim.x * TPB * U) {
        int c[U], row[U], base[U]; unsigned mc[U];
#pragma unroll
        for (int k = 0; k < U; k++) {
            const int i = i0 + k * TPB + threadIdx.x;
            c[k] = -1 - lane; row[k] = -1 - lane;            // invalid lanes match nobody
            if (i < n) {
                const float4 p = __ldg(&jb.pts[i]);
                const int cx = clampi(grid_coord(p.x, ox, inv), 0, dx - 1);
                const int cy = clampi(grid_coord(p.y, oy, inv), 0, dy - 1);
                const int cz = clampi(grid_coord(p.z, oz, inv), 0, dz - 1);
                row[k] = cz * dy + cy;
                c[k] = row[k] * dx + cx;
            }
        }
        // map clouds arrive in voxel order, so neighbouring lanes mostly share a cell (and almost always a row):
        // ONE atomic per distinct cell / row per warp; the leader's return value + the lane's rank among its
        // peers is the point's slot inside its cell
#pragma unroll
        for (int k = 0; k < U; k++) {
            mc[k] = __match_any_sync(FULL, c[k]);
            base[k] = 0;
            if (c[k] >= 0 && lane == __ffs(mc[k]) - 1) base[k] = atomicAdd(&jb.counts[c[k]], __popc(mc[k]));
            const unsigned mr = __match_any_sync(FULL, row[k]);
            if (row[k] >= 0 && lane == __ffs(mr) - 1) atomicAdd(&jb.row_cnt[row[k]], __popc(mr));
        }
#pragma unroll
        for (int k = 0; k < U; k++) {
            const int i = i0 + k * TPB + threadIdx.x;
            const int b = __shfl_sync(FULL, base[k], __ffs(mc[k]) - 1);
            if (i < n) {
                jb.cell_of[i] = c[k];
                jb.rank[i] = b + __popc(mc[k] & lt);
            }
        }
    }
}

// exclusive scan of the per-row point counts -> row_begin[0 .. nrows]; ONE CTA per map (blockIdx.y)
__global__ void __launch_bounds__(1024)
grid_row_scan_kernel(GridJobs jobs, const GridJob *__restrict__ table)
{
    const GridJob &jb = table ? table[blockIdx.y] : jobs.j[blockIdx.y];
    const GridDesc *d = jb.desc;
    __shared__ int s_w[33];
    __shared__ int s_carry;
    const int nrows = d->dim[1] * d->dim[2];
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nrows; base += 4096) {         // 1024 threads x 4 consecutive rows
        const int i0 = base + threadIdx.x * 4;
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = (i0 + k < nrows) ? jb.row_cnt[i0 + k] : 0;
        int total;
        int ex = s_carry + block_excl_scan(v[0] + v[1] + v[2] + v[3], s_w, total);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < nrows) jb.row_begin[i0 + k] = ex;
            ex += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) jb.row_begin[nrows] = s_carry;
}

// cell_begin of every occupied row: one warp per row.  Lane l owns 8 CONSECUTIVE cells of a 256-cell stretch (its loads
// fill whole 32-byte sectors), adds them up, ONE warp scan over the 32 lane sums gives every lane its base: ~45
// instructions per stretch (the first version scanned 8 interleaved strips with 8 warp scans: ~150, and this kernel
// was a quarter of all instructions of a batch step, profiles/r02_knnfit.md).
__global__ void __launch_bounds__(TPB)
grid_row_apply_kernel(GridJobs jobs, const GridJob *__restrict__ table)
{
    const GridJob &jb = table ? table[blockIdx.y] : jobs.j[blockIdx.y];
    const GridDesc *d = jb.desc;
    const int dx = d->dim[0], nrows = d->dim[1] * d->dim[2];
    const int lane = threadIdx.x & 31;
    const int wpb = TPB / 32;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < nrows; r += gridDim.x * wpb) {
        const int rb = jb.row_begin[r], rn = jb.row_begin[r + 1] - rb;
        if (rn == 0) continue;                               // empty rows are never read by a query
        int *__restrict__ cb = jb.cell_begin + (size_t)r * dx;
        const int *__restrict__ cnt = jb.counts + (size_t)r * dx;
        int run = rb;
        for (int x0 = 0; x0 < dx; x0 += 256) {
            const int xl = x0 + lane * 8;
            int v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = xl + j < dx ? cnt[xl + j] : 0;
            int s = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) s += v[j];
            const int inc = warp_incl_scan(s);
            int o = run + inc - s;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (xl + j < dx) cb[xl + j] = o;
                o += v[j];
            }
            run += __shfl_sync(FULL, inc, 31);
        }
        if (lane == 0) cb[dx] = run;                         // end of the row's last cell (= next row's first entry)
    }
}

__global__ void __launch_bounds__(TPB)
grid_scatter_kernel(GridJobs jobs, const GridJob *__restrict__ table)
{
    const GridJob &jb = table ? table[blockIdx.y] : jobs.j[blockIdx.y];
    const GridDesc *d = jb.desc;
    const int n = d->n;
    const int dimx = d->dim[0];
    constexpr int U = 4;
    for (int i0 = blockIdx.x * TPB * U; i0 < n; i0 += gridDim.x * TPB * U) {
        float4 p[U]; int c[U], rk[U], cb[U];
#pragma unroll
        for (int k = 0; k < U; k++) {
            const int i = i0 + k * TPB + threadIdx.x;
            c[k] = -1;
            if (i < n) { p[k] = __ldg(&jb.pts[i]); c[k] = jb.cell_of[i]; rk[k] = jb.rank[i]; }
        }
#pragma unroll
        for (int k = 0; k < U; k++) if (c[k] >= 0) cb[k] = jb.cell_begin[c[k]];
#pragma unroll
        for (int k = 0; k < U; k++) {
            const int i = i0 + k * TPB + threadIdx.x;
            if (c[k] >= 0) {
                jb.sorted[cb[k] + rk[k]] = make_float4(p[k].x, p[k].y, p[k].z, __int_as_float(i));
                jb.counts[c[k]] = 0;                         // leave the count tables clean for the next build
                jb.row_cnt[c[k] / dimx] = 0;
            }
        }
    }
}

__global__ void grid_zero_kernel(int *p, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0;
}

}  // namespace

void GridIndex::init(int max_cells, cudaStream_t s)
{
    max_cells_ = max_cells > 4096 ? max_cells : 4096;
    desc_.ensure(1);
    counts_.ensure((size_t)max_cells_ + 1);
    cell_begin_.ensure((size_t)max_cells_ + 1);
    row_cnt_.ensure((size_t)max_cells_ + 1);
    row_begin_.ensure((size_t)max_cells_ + 2);
    grid_desc_init_kernel<<<1, 1, 0, s>>>(desc_.p);
    grid_zero_kernel<<<148 * 4, 256, 0, s>>>(counts_.p, max_cells_ + 1);
    grid_zero_kernel<<<148 * 4, 256, 0, s>>>(row_cnt_.p, max_cells_ + 1);
    LLB_CUDA(cudaGetLastError());
}

void GridIndex::release()
{
    desc_.release(); sorted_.release(); counts_.release(); cell_begin_.release(); cell_of_.release(); rank_.release();
    row_cnt_.release(); row_begin_.release();
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
        j.cell_of = g[k]->cell_of_.p; j.rank = g[k]->rank_.p; j.row_cnt = g[k]->row_cnt_.p; j.row_begin = g[k]->row_begin_.p;
        j.sorted = g[k]->sorted_.p;
    }
    const dim3 grid_pts(std::min(div_up(nmax, TPB), 148 * 4), 2);
    grid_bbox_kernel<<<grid_pts, TPB, 0, s>>>(jobs, nullptr);
    grid_count_kernel<<<grid_pts, TPB, 0, s>>>(jobs, nullptr);
    grid_row_scan_kernel<<<dim3(1, 2), 1024, 0, s>>>(jobs, nullptr);
    grid_row_apply_kernel<<<dim3(148 * 2, 2), TPB, 0, s>>>(jobs, nullptr);
    grid_scatter_kernel<<<grid_pts, TPB, 0, s>>>(jobs, nullptr);
    LLB_CUDA(cudaGetLastError());
    return 5;
}

GridJob GridIndex::job(const float4 *pts, const int *n_dev, int n_upper)
{
    const int n = n_upper > 0 ? n_upper : 1;
    sorted_.ensure(n); cell_of_.ensure(n); rank_.ensure(n);
    GridJob j;
    j.pts = pts; j.n_dev = n_dev; j.n_host = n_upper;
    j.desc = desc_.p; j.counts = counts_.p; j.cell_begin = cell_begin_.p;
    j.cell_of = cell_of_.p; j.rank = rank_.p; j.row_cnt = row_cnt_.p; j.row_begin = row_begin_.p; j.sorted = sorted_.p;
    return j;
}

int GridIndex::build_table(const GridJob *table_dev, int count, int n_upper_max, float radius, int max_cells,
                           int ctas_per_map, cudaStream_t s)
{
    if (count <= 0) return 0;
    GridJobs jobs{};
    jobs.radius = radius; jobs.max_cells = max_cells;
    // the whole table shares each launch: a few CTAs per map keep the grid near one wave on 148 SMs
    const int per = std::max(1, ctas_per_map);
    const dim3 grid_pts(std::min(div_up(std::max(n_upper_max, 1), TPB), per), count);
    grid_bbox_kernel<<<grid_pts, TPB, 0, s>>>(jobs, table_dev);
    grid_count_kernel<<<grid_pts, TPB, 0, s>>>(jobs, table_dev);
    grid_row_scan_kernel<<<dim3(1, count), 1024, 0, s>>>(jobs, table_dev);
    // one warp per row with two dependent round trips each: 4x the CTAs of the streaming kernels keeps the chains short
    // (16x / 32x were measured too: 0.185 / 0.196 ms for the index stage of a 32-slot step instead of 0.175)
    grid_row_apply_kernel<<<dim3(per * 4, count), TPB, 0, s>>>(jobs, table_dev);
    grid_scatter_kernel<<<grid_pts, TPB, 0, s>>>(jobs, table_dev);
    LLB_CUDA(cudaGetLastError());
    return 5;
}

}  // namespace llb
