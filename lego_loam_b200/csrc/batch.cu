// batch.cu — kernels of the batched multi-registration engine (see batch.cuh): every launch covers
// ALL slots of the batch (blockIdx.y = slot), so a step costs the same number of launches for 1 or
// 64 independent sequences.  The arithmetic is the single-registration path's (s2m_dev.cuh, knn.cuh):
// same expressions, same association order, fp64 accumulation of exact float products in a fixed order.
#include "batch.cuh"
#include "s2m_dev.cuh"
#include "knn.cuh"
#include "cta_radix.cuh"
#include <cstdlib>

namespace llb {

namespace {

constexpr int KNN_NW = BATCH_KNN_THREADS / 32;
constexpr int FIT_NW = BATCH_FIT_THREADS / 32;

__global__ void __launch_bounds__(256)
batch_unpack_kernel(const BatchUnpack *__restrict__ jobs)
{
    const BatchUnpack jb = jobs[blockIdx.y];
    // src: pcl::PointXYZI stride (8 floats): x y z w intensity c1 c2 c3
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < jb.n; i += gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(jb.src32) + 2 * i);
        const float inten = __ldg(jb.src32 + 8 * i + 4);
        jb.dst[i] = make_float4(a.x, a.y, a.z, inten);
    }
}

__global__ void __launch_bounds__(256)
batch_copy_kernel(const BatchCopy *__restrict__ jobs)
{
    const BatchCopy jb = jobs[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < jb.n; i += gridDim.x * blockDim.x) jb.dst[i] = __ldg(&jb.src[i]);
}

__global__ void batch_state_init_kernel(S2mState *st, int B)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    S2mState *s = st + b;
    for (int i = 0; i < 6; i++) { s->T[i] = 0.f; s->cs[i] = (i & 1) ? 0.f : 1.f; s->AtB[i] = 0.f; s->X[i] = 0.f; }
    for (int i = 0; i < 36; i++) { s->matP[i] = 0.f; s->AtA[i] = 0.f; s->AtA0[i] = 0.f; }
    s->matP_valid = 1;                                       // the reference starts with matP = 0 (MO:361)
    s->converged = 0; s->iters = 0; s->n_corr = 0; s->is_degenerate = 0; s->skipped = 0; s->ticket = 0;
}

// pose hand-over, sin/cos of MO:498-506, guard MO:1331, flags: one thread per slot
__global__ void batch_prepare_kernel(const BatchReg *__restrict__ regs, const float *__restrict__ poses, int B, S2mParams prm)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const BatchReg r = regs[b];
    S2mState *st = r.st;
    for (int i = 0; i < 6; i++) st->T[i] = poses[6 * b + i];
    update_sincos(st);
    st->converged = 0;
    st->iters = 0;
    st->n_corr = 0;
    st->ticket = 0;
    st->skipped = !(r.cmap.desc->n > prm.corner_map_min && r.smap.desc->n > prm.surf_map_min);
}

// ---- query binning, once per registration: the queries of a slot are ordered by the MAP cell they fall into at the
// initial pose (stable radix sort in shared memory, one CTA per slot and query kind).  The kNN kernel hands thread j
// the j-th query of that order, so the lanes of a warp search the same cell runs: uniform loop trip counts and
// warp-broadcast candidate loads.  The pose moves by centimetres between LM iterations - far less than a cell - so
// the order stays good for the whole registration; it only affects speed, never results.
constexpr int QS_THREADS = 512;
constexpr int QS_ITEMS = 4;
#ifndef LLB_QS_SHIFT
#define LLB_QS_SHIFT 1      // cost bucket = 2 candidates (8 measured the same)
#endif

__global__ void __launch_bounds__(QS_THREADS)
batch_qsort_kernel(const BatchReg *__restrict__ regs, int cap, int by_count)
{
    extern __shared__ __align__(16) unsigned char qs_smem[];
    unsigned *kin = reinterpret_cast<unsigned *>(qs_smem), *kout = kin + cap;
    unsigned short *vin = reinterpret_cast<unsigned short *>(kout + cap), *vout = vin + cap;
    __shared__ int s_wcnt[QS_THREADS / 32][256];
    __shared__ int s_base[256];
    __shared__ int s_scan[33];
    const BatchReg r = regs[blockIdx.y];
    const S2mState *st = r.st;
    if (st->skipped) return;
    const int which = blockIdx.x;                            // 0: corner queries, 1: surf queries
    const int nc = *r.nc_dev, ns = *r.ns_dev;
    const int n = which == 0 ? nc : ns;
    if (n <= 0 || nc + ns > r.cap) return;
    int *__restrict__ perm = r.qperm + (which == 0 ? 0 : nc);
    const int tid = threadIdx.x;
    if (n > cap || n > 65535) {                              // does not fit the shared-memory sort: keep scan order
        for (int i = tid; i < n; i += QS_THREADS) perm[i] = i;
        return;
    }
    const float crx = st->cs[0], srx = st->cs[1], cry = st->cs[2], sry = st->cs[3], crz = st->cs[4], srz = st->cs[5];
    const float tX = st->T[3], tY = st->T[4], tZ = st->T[5];
    const GridDesc *g = which == 0 ? r.cmap.desc : r.smap.desc;
    const float4 *__restrict__ qs = which == 0 ? r.corner : r.surf;
    const int dimx = g->dim[0], dimy = g->dim[1], dimz = g->dim[2];
    const float inv = g->inv_cell, ox = g->org[0], oy = g->org[1], oz = g->org[2];
    for (int i = tid; i < n; i += QS_THREADS) {
        float sx, sy, sz;
        associate_to_map(crx, srx, cry, sry, crz, srz, tX, tY, tZ, __ldg(&qs[i]), sx, sy, sz);
        const int cx = min(max(grid_coord(sx, ox, inv), 0), dimx - 1);
        const int cy = min(max(grid_coord(sy, oy, inv), 0), dimy - 1);
        const int cz = min(max(grid_coord(sz, oz, inv), 0), dimz - 1);
        unsigned key = (unsigned)((cz * dimy + cy) * dimx + cx);
        if (by_count) {
            // key = number of candidates in the 3x3x3 cell neighbourhood (what the kNN thread of this query will walk):
            // threads of a warp then get queries of (nearly) equal cost
            const int *__restrict__ cell_begin = which == 0 ? r.cmap.cell_begin : r.smap.cell_begin;
            const int *__restrict__ row_begin = which == 0 ? r.cmap.row_begin : r.smap.row_begin;
            const int x0 = max(cx - 1, 0), x1 = min(cx + 1, dimx - 1);
            int tot = 0;
            // the three rows of the query's own z-slab (12 independent loads): a cheap proxy of the 9-row total
#pragma unroll
            for (int dy = -1; dy <= 1; dy++) {
                const int y = cy + dy;
                if (y < 0 || y >= dimy) continue;
                const int ry = cz * dimy + y;
                const int r0 = __ldg(&row_begin[ry]), r1 = __ldg(&row_begin[ry + 1]);
                const int c0 = __ldg(&cell_begin[ry * dimx + x0]), c1 = __ldg(&cell_begin[ry * dimx + x1 + 1]);
                tot += (r1 > r0) ? (c1 - c0) : 0;
            }
            key = (unsigned)min(tot >> LLB_QS_SHIFT, 255);    // one 8-bit radix pass is enough to equalise warps
        }
        kin[i] = key;
        vin[i] = (unsigned short)i;
    }
    int nbits = 1;
    if (by_count) nbits = 8;
    else while (nbits < 32 && ((unsigned)(g->ncell - 1) >> nbits) != 0u) nbits++;
    cta_radix_sort<QS_THREADS, QS_ITEMS>(kin, kout, vin, vout, n, nbits, s_wcnt, s_base, s_scan);
    for (int i = tid; i < n; i += QS_THREADS) perm[i] = (int)vin[i];
}

// ---- iteration, step 1: pointAssociateToMap + radius-bounded exact 5-NN, one THREAD per query, two phases.
// Throughput form of knn.cuh's search.  Phase 1 walks the (up to) nine cell runs around the query and appends every
// candidate INSIDE the gate (d2 < 1, ~15 % of them) to a per-thread list in shared memory: ~14 straight-line
// instructions per candidate, no divergent path.  Phase 2 pushes the ~20 listed candidates through a branch-free
// sorted insertion and keeps the five smallest (distance, original index) pairs - the oracle's tie rule.  (A first
// single-phase version kept the top five while scanning: its rarely-taken insertion path ran with 1-3 active lanes
// and ncu showed 8.5 active lanes per instruction on average; profiles/r01c_batch.md.)
// Consecutive queries are voxel-ordered DS points, i.e. spatial neighbours: the lanes of a warp read the same cells
// and the float4 candidate loads hit L1 (81 %).  Rows whose slab is farther than the gate radius are skipped (1 %
// margin: cell membership is decided by the same floorf expression for map points and queries, rounding
// differences are orders of magnitude below the margin); empty rows are skipped through the row directory.
constexpr int KNN_LCAP = 32;                                 // list entries per thread (compressed to 5 when full)

// five smallest (d, original index) of the thread's list, ascending, by branch-free sorted insertion.  The list holds
// positions in `sorted` only (4 B per entry keeps shared memory small and L1 large); distance and original index (the
// tie-breaker) are re-derived from the candidate itself - the same float expression, hence the same bits.
__device__ __forceinline__ void knn_select5_thread(const int (*li)[BATCH_KNN_THREADS], int cnt, int tid,
                                                   const float4 *__restrict__ sorted, float qx, float qy, float qz,
                                                   float (&bd)[5], int (&bi)[5], int (&bp)[5])
{
#pragma unroll
    for (int k = 0; k < 5; k++) { bd[k] = __int_as_float(0x7f800000); bi[k] = 0x7fffffff; bp[k] = -1; }
    for (int j = 0; j < cnt; j++) {
        int pos = li[j][tid];
        const float4 p = __ldg(&sorted[pos]);
        float d = l2_simple(qx, qy, qz, p);
        int oi = __float_as_int(p.w);
#pragma unroll
        for (int k = 0; k < 5; k++) {                        // (d, oi, pos) sinks to its place, the rest shifts down
            const bool lt = d < bd[k] || (d == bd[k] && oi < bi[k]);
            const float td = lt ? bd[k] : d; const int ti = lt ? bi[k] : oi; const int tp = lt ? bp[k] : pos;
            bd[k] = lt ? d : bd[k]; bi[k] = lt ? oi : bi[k]; bp[k] = lt ? pos : bp[k];
            d = td; oi = ti; pos = tp;
        }
    }
}

__global__ void __launch_bounds__(BATCH_KNN_THREADS, 4)
batch_knn_kernel(const BatchReg *__restrict__ regs, S2mParams prm)
{
    __shared__ int s_li[KNN_LCAP][BATCH_KNN_THREADS];
    const BatchReg r = regs[blockIdx.y];
    const S2mState *st = r.st;
    if (__ldcg(&st->skipped) || __ldcg(&st->converged)) return;
    const int tid = threadIdx.x;
    const float crx = __ldcg(&st->cs[0]), srx = __ldcg(&st->cs[1]), cry = __ldcg(&st->cs[2]),
                sry = __ldcg(&st->cs[3]), crz = __ldcg(&st->cs[4]), srz = __ldcg(&st->cs[5]);
    const float tX = __ldcg(&st->T[3]), tY = __ldcg(&st->T[4]), tZ = __ldcg(&st->T[5]);
    const int nc = *r.nc_dev, ns = *r.ns_dev;
    const int nq = min(nc + ns, r.cap);
    const float max_sq = prm.knn_max_sqdist;
    const float prune_sq = max_sq * 1.01f;
    for (int j = blockIdx.x * BATCH_KNN_THREADS + tid; j < nq; j += gridDim.x * BATCH_KNN_THREADS) {
        const bool is_corner = j < nc;
        const int q = (is_corner ? 0 : nc) + __ldg(&r.qperm[j]);      // cell-ordered query (batch_qsort_kernel)
        const float4 po = is_corner ? __ldg(&r.corner[q]) : __ldg(&r.surf[q - nc]);
        float sx, sy, sz;
        associate_to_map(crx, srx, cry, sry, crz, srz, tX, tY, tZ, po, sx, sy, sz);
        const GridDesc *g = is_corner ? r.cmap.desc : r.smap.desc;
        const int *__restrict__ cell_begin = is_corner ? r.cmap.cell_begin : r.smap.cell_begin;
        const int *__restrict__ row_begin = is_corner ? r.cmap.row_begin : r.smap.row_begin;
        const float4 *__restrict__ sorted = is_corner ? r.cmap.sorted : r.smap.sorted;
        const int dimx = g->dim[0], dimy = g->dim[1], dimz = g->dim[2];
        const float inv = g->inv_cell, cell = g->cell;
        const float fy = (sy - g->org[1]) * inv, fz = (sz - g->org[2]) * inv;
        const int cx = grid_coord(sx, g->org[0], inv), cy = (int)floorf(fy), cz = (int)floorf(fz);
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, dimx - 1);
        // distance from the query to the neighbouring rows' slabs, in metres (0 for the query's own row)
        const float ly = (fy - (float)cy) * cell, lz = (fz - (float)cz) * cell;
        int found = 0;
        int *my_list = &s_li[0][tid];
        if (x0 <= x1) {
#pragma unroll 1
            for (int rr = 0; rr < 9; rr++) {
                const int dy = (rr % 3) - 1, dz = (rr / 3) - 1;
                const int y = cy + dy, z = cz + dz;
                if (y < 0 || y >= dimy || z < 0 || z >= dimz) continue;
                const float gy = dy < 0 ? ly : (dy > 0 ? cell - ly : 0.f);
                const float gz = dz < 0 ? lz : (dz > 0 ? cell - lz : 0.f);
                if (gy * gy + gz * gz > prune_sq) continue;
                const int ry = z * dimy + y;
                if (__ldg(&row_begin[ry + 1]) == __ldg(&row_begin[ry])) continue;     // empty row: no cell table there
                const int row = ry * dimx;
                const int rb = __ldg(&cell_begin[row + x0]), re = __ldg(&cell_begin[row + x1 + 1]);
#pragma unroll 4
                for (int i = rb; i < re; i++) {
                    const float4 p = __ldg(&sorted[i]);
                    const float d = l2_simple(sx, sy, sz, p);
                    const bool in = d < max_sq;
                    if (in && found < KNN_LCAP) my_list[found * BATCH_KNN_THREADS] = i;   // predicated store
                    found += in ? 1 : 0;
                }
            }
        }
        int cnt = min(found, KNN_LCAP);
        if (found > KNN_LCAP) {
            // more in-gate candidates than list entries (dense corner clusters, rare): second walk that keeps the list
            // compressed to its best five whenever it fills up
            cnt = 0;
#pragma unroll 1
            for (int rr = 0; rr < 9; rr++) {
                const int y = cy + (rr % 3) - 1, z = cz + (rr / 3) - 1;
                if (y < 0 || y >= dimy || z < 0 || z >= dimz) continue;
                const int ry = z * dimy + y;
                if (__ldg(&row_begin[ry + 1]) == __ldg(&row_begin[ry])) continue;
                const int row = ry * dimx;
                const int rb = __ldg(&cell_begin[row + x0]), re = __ldg(&cell_begin[row + x1 + 1]);
#pragma unroll 1
                for (int i = rb; i < re; i++) {
                    const float d = l2_simple(sx, sy, sz, __ldg(&sorted[i]));
                    if (d < max_sq) {
                        if (cnt == KNN_LCAP) {
                            float cd[5]; int ci[5], cp[5];
                            knn_select5_thread(s_li, cnt, tid, sorted, sx, sy, sz, cd, ci, cp);
#pragma unroll
                            for (int k = 0; k < 5; k++) s_li[k][tid] = cp[k];
                            cnt = 5;
                        }
                        s_li[cnt][tid] = i;
                        cnt++;
                    }
                }
            }
        }
        float bd[5]; int bi[5], bp[5];
        knn_select5_thread(s_li, cnt, tid, sorted, sx, sy, sz, bd, bi, bp);
#pragma unroll
        for (int k = 0; k < 5; k++) r.nn[(size_t)k * r.cap + q] = found >= 5 ? bp[k] : -1;
        r.d5[q] = found >= 5 ? bd[4] : -1.f;
    }
}

// kNN variant 1 (LLB_KNN_VARIANT=1): single phase, the five best (distance, original index) pairs live in registers
// while the thread scans its candidates row by row, queries in scan order.  Measured on B200 (64 slots x ~3.5k
// queries, profiles/r01c_batch.md): 118 us per iteration vs 157 us for the two-phase list variant above
// (LLB_KNN_VARIANT=2), with or without cell-ordered queries: both are bound by unequal candidate counts of neighbouring
// lanes (8-14 active lanes per instruction).  Variant 3 below (the default) removes that imbalance.
__global__ void __launch_bounds__(256, 4)
batch_knn1_kernel(const BatchReg *__restrict__ regs, S2mParams prm)
{
    const BatchReg r = regs[blockIdx.y];
    const S2mState *st = r.st;
    if (__ldcg(&st->skipped) || __ldcg(&st->converged)) return;
    const float crx = __ldcg(&st->cs[0]), srx = __ldcg(&st->cs[1]), cry = __ldcg(&st->cs[2]),
                sry = __ldcg(&st->cs[3]), crz = __ldcg(&st->cs[4]), srz = __ldcg(&st->cs[5]);
    const float tX = __ldcg(&st->T[3]), tY = __ldcg(&st->T[4]), tZ = __ldcg(&st->T[5]);
    const int nc = *r.nc_dev, ns = *r.ns_dev;
    const int nq = min(nc + ns, r.cap);
    const float max_sq = prm.knn_max_sqdist;
    const float prune_sq = max_sq * 1.01f;
    for (int j = blockIdx.x * 256 + threadIdx.x; j < nq; j += gridDim.x * 256) {
        const bool is_corner = j < nc;
        const int q = j;                                     // scan order (cell-ordered queries measured no faster)
        const float4 po = is_corner ? __ldg(&r.corner[q]) : __ldg(&r.surf[q - nc]);
        float sx, sy, sz;
        associate_to_map(crx, srx, cry, sry, crz, srz, tX, tY, tZ, po, sx, sy, sz);
        const GridDesc *g = is_corner ? r.cmap.desc : r.smap.desc;
        const int *__restrict__ cell_begin = is_corner ? r.cmap.cell_begin : r.smap.cell_begin;
        const int *__restrict__ row_begin = is_corner ? r.cmap.row_begin : r.smap.row_begin;
        const float4 *__restrict__ sorted = is_corner ? r.cmap.sorted : r.smap.sorted;
        const int dimx = g->dim[0], dimy = g->dim[1], dimz = g->dim[2];
        const float inv = g->inv_cell, cell = g->cell;
        const float fy = (sy - g->org[1]) * inv, fz = (sz - g->org[2]) * inv;
        const int cx = grid_coord(sx, g->org[0], inv), cy = (int)floorf(fy), cz = (int)floorf(fz);
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, dimx - 1);
        // distance from the query to the neighbouring rows' slabs, in metres (0 for the query's own row)
        const float ly = (fy - (float)cy) * cell, lz = (fz - (float)cz) * cell;
        // the list starts as five sentinels (gate distance, index -1): "d < bd[4]" is then the gate test AND the top-5
        // test in one compare, and five real neighbours were found iff the last sentinel has been pushed out
        float bd[5]; int bi[5], bp[5];
#pragma unroll
        for (int k = 0; k < 5; k++) { bd[k] = max_sq; bi[k] = -1; bp[k] = -1; }
        if (x0 <= x1) {
#pragma unroll 1
            for (int ro = 0; ro < 9; ro++) {
                // own row first, then the four face neighbours, then the corners: the 5th-best distance tightens early
                // and most later candidates fail the single compare (fewer trips through the insertion path)
                const int rr = (int)((0x862075314ull >> (4 * ro)) & 15ull);
                const int dy = (rr % 3) - 1, dz = (rr / 3) - 1;
                const int y = cy + dy, z = cz + dz;
                if (y < 0 || y >= dimy || z < 0 || z >= dimz) continue;
                const float gy = dy < 0 ? ly : (dy > 0 ? cell - ly : 0.f);
                const float gz = dz < 0 ? lz : (dz > 0 ? cell - lz : 0.f);
                if (gy * gy + gz * gz > prune_sq) continue;
                const int ry = z * dimy + y;
                if (__ldg(&row_begin[ry + 1]) == __ldg(&row_begin[ry])) continue;     // empty row: no cell table there
                const int row = ry * dimx;
                const int rb = __ldg(&cell_begin[row + x0]), re = __ldg(&cell_begin[row + x1 + 1]);
#pragma unroll 2
                for (int i = rb; i < re; i++) {
                    const float4 p = __ldg(&sorted[i]);
                    const float d = l2_simple(sx, sy, sz, p);
                    const int oi = __float_as_int(p.w);
                    if (d < bd[4] || (d == bd[4] && oi < bi[4])) {   // sentinels have index -1: never beaten on a tie
                        // insert (d, oi) into the ascending list, dropping the last entry
                        bd[4] = d; bi[4] = oi; bp[4] = i;
#pragma unroll
                        for (int k = 4; k > 0; k--) {
                            const bool sw = bd[k] < bd[k - 1] || (bd[k] == bd[k - 1] && bi[k] < bi[k - 1]);
                            if (sw) {
                                const float td = bd[k]; bd[k] = bd[k - 1]; bd[k - 1] = td;
                                const int ti = bi[k]; bi[k] = bi[k - 1]; bi[k - 1] = ti;
                                const int tp = bp[k]; bp[k] = bp[k - 1]; bp[k - 1] = tp;
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 5; k++) r.nn[(size_t)k * r.cap + q] = bp[4] >= 0 ? bp[k] : -1;
        r.d5[q] = bp[4] >= 0 ? bd[4] : -1.f;
    }
}

// DEFAULT kNN kernel, variant 3 (LLB_KNN_VARIANT=3): the register top-5 search of batch_knn1_kernel with the two causes of its low lane
// utilisation (8.7 active lanes per instruction) removed: (a) the queries are handed out in order of their candidate
// count (batch_qsort_kernel, by_count) so the threads of a warp have equal work, (b) the nine cell runs of a query are
// compacted into a per-thread list in shared memory and walked as ONE flattened loop with the next candidate's load
// issued before the current one is processed, so a warp iterates max(total) times instead of sum over rows of max(row).
__global__ void __launch_bounds__(256, 4)
batch_knn3_kernel(const BatchReg *__restrict__ regs, S2mParams prm)
{
    __shared__ int s_rb[9][256], s_re[9][256];
    const BatchReg r = regs[blockIdx.y];
    const S2mState *st = r.st;
    if (__ldcg(&st->skipped) || __ldcg(&st->converged)) return;
    const int tid = threadIdx.x;
    const float crx = __ldcg(&st->cs[0]), srx = __ldcg(&st->cs[1]), cry = __ldcg(&st->cs[2]),
                sry = __ldcg(&st->cs[3]), crz = __ldcg(&st->cs[4]), srz = __ldcg(&st->cs[5]);
    const float tX = __ldcg(&st->T[3]), tY = __ldcg(&st->T[4]), tZ = __ldcg(&st->T[5]);
    const int nc = *r.nc_dev, ns = *r.ns_dev;
    const int nq = min(nc + ns, r.cap);
    const float max_sq = prm.knn_max_sqdist;
    const float prune_sq = max_sq * 1.01f;
    for (int j = blockIdx.x * 256 + tid; j < nq; j += gridDim.x * 256) {
        const bool is_corner = j < nc;
        const int q = (is_corner ? 0 : nc) + __ldg(&r.qperm[j]);
        const float4 po = is_corner ? __ldg(&r.corner[q]) : __ldg(&r.surf[q - nc]);
        float sx, sy, sz;
        associate_to_map(crx, srx, cry, sry, crz, srz, tX, tY, tZ, po, sx, sy, sz);
        const GridDesc *g = is_corner ? r.cmap.desc : r.smap.desc;
        const int *__restrict__ cell_begin = is_corner ? r.cmap.cell_begin : r.smap.cell_begin;
        const int *__restrict__ row_begin = is_corner ? r.cmap.row_begin : r.smap.row_begin;
        const float4 *__restrict__ sorted = is_corner ? r.cmap.sorted : r.smap.sorted;
        const int dimx = g->dim[0], dimy = g->dim[1], dimz = g->dim[2];
        const float inv = g->inv_cell, cell = g->cell;
        const float fy = (sy - g->org[1]) * inv, fz = (sz - g->org[2]) * inv;
        const int cx = grid_coord(sx, g->org[0], inv), cy = (int)floorf(fy), cz = (int)floorf(fz);
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, dimx - 1);
        const float ly = (fy - (float)cy) * cell, lz = (fz - (float)cz) * cell;
        // ---- the non-empty, non-pruned runs of this query, own row first (see batch_knn1_kernel)
        int nr = 0;
        if (x0 <= x1) {
#pragma unroll
            for (int ro = 0; ro < 9; ro++) {
                const int rr = (int)((0x862075314ull >> (4 * ro)) & 15ull);
                const int dy = (rr % 3) - 1, dz = (rr / 3) - 1;
                const int y = cy + dy, z = cz + dz;
                const float gy = dy < 0 ? ly : (dy > 0 ? cell - ly : 0.f);
                const float gz = dz < 0 ? lz : (dz > 0 ? cell - lz : 0.f);
                if (y >= 0 && y < dimy && z >= 0 && z < dimz && gy * gy + gz * gz <= prune_sq) {
                    const int ry = z * dimy + y;
                    const int r0 = __ldg(&row_begin[ry]), r1 = __ldg(&row_begin[ry + 1]);
                    const int rb = __ldg(&cell_begin[ry * dimx + x0]), re = __ldg(&cell_begin[ry * dimx + x1 + 1]);
                    if (r1 > r0 && re > rb) { s_rb[nr][tid] = rb; s_re[nr][tid] = re; nr++; }
                }
            }
        }
        float bd[5]; int bi[5], bp[5];
#pragma unroll
        for (int k = 0; k < 5; k++) { bd[k] = max_sq; bi[k] = -1; bp[k] = -1; }
        // ---- flattened walk, software-pipelined by one candidate
        int run = 0, i = 0, e = 0;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nr > 0) { i = s_rb[0][tid]; e = s_re[0][tid]; p = __ldg(&sorted[i]); }
        while (run < nr) {
            int ni = i + 1, nrun = run, ne = e;
            if (ni >= e) { nrun = run + 1; if (nrun < nr) { ni = s_rb[nrun][tid]; ne = s_re[nrun][tid]; } }
            float4 pn = p;
            if (nrun < nr) pn = __ldg(&sorted[ni]);
            const float d = l2_simple(sx, sy, sz, p);
            const int oi = __float_as_int(p.w);
            if (d < bd[4] || (d == bd[4] && oi < bi[4])) {
                bd[4] = d; bi[4] = oi; bp[4] = i;
#pragma unroll
                for (int k = 4; k > 0; k--) {
                    const bool sw = bd[k] < bd[k - 1] || (bd[k] == bd[k - 1] && bi[k] < bi[k - 1]);
                    if (sw) {
                        const float td = bd[k]; bd[k] = bd[k - 1]; bd[k - 1] = td;
                        const int ti = bi[k]; bi[k] = bi[k - 1]; bi[k - 1] = ti;
                        const int tp = bp[k]; bp[k] = bp[k - 1]; bp[k - 1] = tp;
                    }
                }
            }
            p = pn; i = ni; run = nrun; e = ne;
        }
#pragma unroll
        for (int k = 0; k < 5; k++) r.nn[(size_t)k * r.cap + q] = bp[4] >= 0 ? bp[k] : -1;
        r.d5[q] = bp[4] >= 0 ? bd[4] : -1.f;
    }
}

// ---- iteration, step 2: gate, line / plane fit, residual, Jacobian row (one THREAD per query; corner and
// surf queries live in different warps), then the 28 fp64 products of the rows of this CTA
__global__ void __launch_bounds__(BATCH_FIT_THREADS, 3)
batch_fit_kernel(const BatchReg *__restrict__ regs, int iter, S2mParams prm)
{
    __shared__ float s_row[8][BATCH_FIT_THREADS + 1];        // +1: lane k reads row ia(k), same column -> distinct banks
    __shared__ double s_acc[FIT_NW][32];
    const BatchReg r = regs[blockIdx.y];
    const S2mState *st = r.st;
    if (__ldcg(&st->skipped) || __ldcg(&st->converged)) return;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const float crx = __ldcg(&st->cs[0]), srx = __ldcg(&st->cs[1]), cry = __ldcg(&st->cs[2]),
                sry = __ldcg(&st->cs[3]), crz = __ldcg(&st->cs[4]), srz = __ldcg(&st->cs[5]);
    const float tX = __ldcg(&st->T[3]), tY = __ldcg(&st->T[4]), tZ = __ldcg(&st->T[5]);
    const int nc = *r.nc_dev, ns = *r.ns_dev;
    const int nq = min(nc + ns, r.cap);
    const int nc_pad = (nc + 31) & ~31;                      // surf rows start on a warp boundary
    const int total = nc_pad + (nq - min(nc, nq));
    int ia, ib;
    pair_of(lane, ia, ib);
    double acc = 0.0;
    for (int base = blockIdx.x * BATCH_FIT_THREADS; base < total; base += gridDim.x * BATCH_FIT_THREADS) {
        const int s = base + tid;
        const bool is_corner = s < nc_pad;
        const int q = is_corner ? s : s - nc_pad + nc;
        const bool live = is_corner ? (s < nc && s < nq) : (q < nq);
        float v[8] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
        if (live) {
            const float d5 = r.d5[q];
            if ((d5 >= 0.f) && ((double)d5 < (double)prm.knn_max_sqdist)) {       // MO:1101 / MO:1183
                const float4 po = is_corner ? __ldg(&r.corner[q]) : __ldg(&r.surf[q - nc]);
                float sx, sy, sz;
                associate_to_map(crx, srx, cry, sry, crz, srz, tX, tY, tZ, po, sx, sy, sz);
                const float4 *sorted = is_corner ? r.cmap.sorted : r.smap.sorted;
                float nx[5], ny[5], nz[5];
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const float4 p = __ldg(&sorted[r.nn[(size_t)k * r.cap + q]]);
                    nx[k] = p.x; ny[k] = p.y; nz[k] = p.z;
                }
                float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
                const bool ok = is_corner ? corner_fit(nx, ny, nz, sx, sy, sz, coeff) : surf_fit(nx, ny, nz, sx, sy, sz, coeff);
                if (ok) jacobian_row(crx, srx, cry, sry, crz, srz, po.x, po.y, po.z, coeff, v);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) s_row[k][tid] = v[k];
        __syncthreads();
        if (lane < S2M_ACC)
            for (int rr = w; rr < BATCH_FIT_THREADS; rr += FIT_NW) acc += (double)s_row[ia][rr] * (double)s_row[ib][rr];
        __syncthreads();
    }
    s_acc[w][lane] = acc;
    __syncthreads();
    if (tid < S2M_ACC) {
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < FIT_NW; k++) sum += s_acc[k][tid];
        r.partials[(size_t)blockIdx.x * S2M_ACC + tid] = sum;
    }
    // ---- the LAST CTA of this slot to get here adds the CTA partials in a fixed order and performs the
    // LMOptimization tail (MO:1273-1326): no separate launch, no host round trip
    __shared__ int s_last;
    __shared__ double s_tot[32];
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&r.st->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid < 32) {
        double s = 0.0;
        if (lane < S2M_ACC)
            for (int b = 0; b < (int)gridDim.x; b++) s += __ldcg(&r.partials[(size_t)b * S2M_ACC + lane]);
        s_tot[lane] = s;
        __syncwarp();
        S2mState *stw = r.st;
        if (lane == 0) { stw->ticket = 0; lm_solve(stw, s_tot, iter, prm, false); }
        __syncwarp();
        if (lane < 6) {                                      // sin/cos of the new pose, one per lane
            stw->cs[lane] = pose_trig(stw->T, lane);
        }
    }
}

__global__ void batch_collect_kernel(const BatchReg *__restrict__ regs, int B, BatchResult *__restrict__ out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const BatchReg r = regs[b];
    const S2mState *st = r.st;
    BatchResult o;
    for (int i = 0; i < 6; i++) o.T[i] = st->T[i];
    o.iters = st->iters; o.converged = st->converged; o.n_corr = st->n_corr; o.is_degenerate = st->is_degenerate;
    o.skipped = st->skipped; o.nc = *r.nc_dev; o.ns = *r.ns_dev; o.pad = 0;
    for (int k = 0; k < 4; k++) o.ds[k] = r.nc_dev[k];       // nc_dev points at the slot's four DS counts
    out[b] = o;
}

}  // namespace

void launch_batch_copy(const BatchCopy *jobs_dev, int count, int n_max, cudaStream_t s)
{
    if (count <= 0) return;
    const dim3 grid(std::max(1, std::min(div_up(std::max(n_max, 1), 256), 32)), count);
    batch_copy_kernel<<<grid, 256, 0, s>>>(jobs_dev);
    LLB_CUDA(cudaGetLastError());
}

void launch_batch_unpack(const BatchUnpack *jobs_dev, int count, int n_max, cudaStream_t s)
{
    if (count <= 0) return;
    const dim3 grid(std::max(1, std::min(div_up(std::max(n_max, 1), 256), 64)), count);
    batch_unpack_kernel<<<grid, 256, 0, s>>>(jobs_dev);
    LLB_CUDA(cudaGetLastError());
}

void launch_batch_state_init(S2mState *st, int B, cudaStream_t s)
{
    batch_state_init_kernel<<<div_up(B, 128), 128, 0, s>>>(st, B);
    LLB_CUDA(cudaGetLastError());
}

void launch_batch_prepare(const BatchReg *regs, const float *poses_dev, int B, const S2mParams &prm, cudaStream_t s)
{
    batch_prepare_kernel<<<div_up(B, 128), 128, 0, s>>>(regs, poses_dev, B, prm);
    LLB_CUDA(cudaGetLastError());
}

int batch_knn_variant()
{
    static const int variant = getenv("LLB_KNN_VARIANT") ? atoi(getenv("LLB_KNN_VARIANT")) : 3;
    return variant;
}

void launch_batch_qsort(const BatchReg *regs, int B, int cap, cudaStream_t s)
{
    const int bytes = cap * 12;
    if (bytes > 48 * 1024)                                   // opt-in above the default limit (idempotent, per device)
        LLB_CUDA(cudaFuncSetAttribute(batch_qsort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 12));
    batch_qsort_kernel<<<dim3(2, B), QS_THREADS, bytes, s>>>(regs, cap, batch_knn_variant() == 3 ? 1 : 0);
    LLB_CUDA(cudaGetLastError());
}

void launch_batch_knn(const BatchReg *regs, int B, int ctas_per_slot, const S2mParams &prm, cudaStream_t s)
{
    if (batch_knn_variant() == 3) {
        batch_knn3_kernel<<<dim3(std::max(1, ctas_per_slot * BATCH_KNN_THREADS / 256), B), 256, 0, s>>>(regs, prm);
        LLB_CUDA(cudaGetLastError());
        return;
    }
    if (batch_knn_variant() != 2) {
        batch_knn1_kernel<<<dim3(std::max(1, ctas_per_slot * BATCH_KNN_THREADS / 256), B), 256, 0, s>>>(regs, prm);
        LLB_CUDA(cudaGetLastError());
        return;
    }
    batch_knn_kernel<<<dim3(std::max(1, ctas_per_slot), B), BATCH_KNN_THREADS, 0, s>>>(regs, prm);
    LLB_CUDA(cudaGetLastError());
}

void launch_batch_fit(const BatchReg *regs, int B, int fit_blocks, int iter, const S2mParams &prm, cudaStream_t s)
{
    batch_fit_kernel<<<dim3(std::max(1, fit_blocks), B), BATCH_FIT_THREADS, 0, s>>>(regs, iter, prm);
    LLB_CUDA(cudaGetLastError());
}

void launch_batch_collect(const BatchReg *regs, int B, BatchResult *out, cudaStream_t s)
{
    batch_collect_kernel<<<div_up(B, 128), 128, 0, s>>>(regs, B, out);
    LLB_CUDA(cudaGetLastError());
}

}  // namespace llb
