// batch.cu — kernels of the batched multi-registration engine (see batch.cuh): every launch covers ALL slots of the
// batch, so a step costs the same number of launches for 1 or 64 independent sequences, and scan2MapOptimization of all
// slots is ONE persistent kernel (batch_lm_kernel: chunks of 32 queries handed out through per-slot control words, the
// warp that delivers the last chunk of a (slot, iteration) takes the LM step and releases the next iteration).  The
// arithmetic is the single-registration path's (s2m_dev.cuh, knn.cuh): same expressions, same association order, fp64
// accumulation of exact float products in a fixed order.  Measured history: profiles/r02_knnfit.md.
#include "batch.cuh"
#include "s2m_dev.cuh"
#include "knn.cuh"
#include "cta_radix.cuh"
#include <cstdlib>

namespace llb {

namespace {

constexpr int ITER_NW = BATCH_ITER_THREADS / 32;
constexpr int KNN_RUNS = 8;                                 // neighbour rows of a query's own (y,z) row of cells

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
// chunks of 32 queries of a slot: corner queries padded to a warp boundary, then the surf queries
__device__ __forceinline__ int batch_chunks(int nc, int ns, int cap)
{
    const int nq = min(nc + ns, cap);
    const int nc_pad = (nc + 31) & ~31;
    return max(1, (nc_pad + (nq - min(nc, nq)) + 31) >> 5);   // an empty sweep still gets its (empty) LM steps, as in the reference
}

__global__ void batch_prepare_kernel(const BatchReg *__restrict__ regs, const float *__restrict__ poses, int B, S2mParams prm,
                                     int max_iter, BatchQueue *__restrict__ queue)
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
    // work hand-out of the registration kernel: per-slot control word (iteration << 16 | next chunk)
    const int nc = *r.nc_dev, ns = *r.ns_dev, n_chunks = batch_chunks(nc, ns, r.cap);
    const bool runs = !st->skipped && max_iter > 0;
    queue->slot[b] = BatchSlotInfo{ nc, ns, runs ? 0 : 1, n_chunks };
    queue->ctl[b] = runs ? 0u : 0xffff8000u;                 // iteration 0, chunk 0 | nothing to hand out (CTL_DONE)
    if (runs) atomicAdd(&queue->live, 1);
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

// ---- THE registration kernel: scan2MapOptimization (MO:1329-1350) of ALL slots in ONE persistent launch: every LM
// iteration of every slot, i.e. cornerOptimization + surfOptimization + the products of LMOptimization + the LM step; the
// neighbour positions never leave the registers (96 B per query-iteration is the contract of the fused
// form).  Work item = 32 queries of one slot (a "chunk"), executed by one autonomous warp, one THREAD per query (corner
// and surf queries in different warps, handed out in the cost order of batch_qsort_kernel):
//   1. pointAssociateToMap (MO:513-527)
//   2. exact 5-NN inside a BOUND that starts at the reference's gate (MO:1101 / MO:1183: d2[4] < 1.0) or, from the second
//      iteration on, at (r5_prev + |q - q_prev|)^2 -- the five neighbours of the previous iteration are still that close,
//      so nothing farther can be among the five nearest now.  The query's own (y,z) row of cells is ranked first and the
//      neighbour rows' runs are narrowed to the cells that intersect the 5th-best ball: ~20 candidates instead of the
//      ~36 of the 3x3x3 neighbourhood (fewer once the bound is seeded).  Exact for ANY bound: five neighbours found
//      inside the bound are the five nearest; otherwise the search is repeated with the gate.
//   3. line / plane fit, residual, Jacobian row (s2m_dev.cuh), 4. the chunk's 28 fp64 products, 5. the warp that delivers
//      the LAST chunk of (slot, iteration) adds the chunk partials in chunk order and performs the LM step
//      (MO:1273-1326), then RELEASES the slot's next iteration.
// Slots advance through their iterations independently (per-slot control word: iteration | next chunk): no launch and
// no grid-wide barrier per iteration, and a slot that converges early or late costs exactly its own work.  (One launch
// per iteration was bound by the latency of a single item plus the serial LM step at its end: ~60 us per launch even
// with 4 of 32 slots live, profiles/r02_knnfit.md.)
__device__ __forceinline__ void knn5_insert(float d, int oi, int pos, float (&bd)[5], int (&bi)[5], int (&bp)[5])
{
    bd[4] = d; bi[4] = oi; bp[4] = pos;
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

// exact 5 nearest (distance, original index) of (sx,sy,sz) among the map points with squared distance < bound_sq
// (strictly; sentinels carry index -1 and lose every tie).  bp[4] >= 0 <=> five were found.
__device__ __forceinline__ void knn5_bounded(const MapIndexView &m, float sx, float sy, float sz, float bound_sq,
                                             float (&bd)[5], int (&bi)[5], int (&bp)[5])
{
    const GridDesc *g = m.desc;
    const int dimx = g->dim[0], dimy = g->dim[1], dimz = g->dim[2];
    const float inv = g->inv_cell, cell = g->cell;
    const float fx = (sx - g->org[0]) * inv, fy = (sy - g->org[1]) * inv, fz = (sz - g->org[2]) * inv;
    const int cx = (int)floorf(fx), cy = (int)floorf(fy), cz = (int)floorf(fz);
#pragma unroll
    for (int k = 0; k < 5; k++) { bd[k] = bound_sq; bi[k] = -1; bp[k] = -1; }
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, dimx - 1);
    if (x0 > x1) return;
    // distances from the query to the faces of its own cell, in metres.  Cell membership of map points and queries comes
    // from the same floorf expression; the rounding of these offsets (~1e-5 m) is far below the pruning margins
    const float lx = (fx - (float)cx) * cell, ly = (fy - (float)cy) * cell, lz = (fz - (float)cz) * cell;
    const bool x_in = cx >= 0 && cx < dimx;                 // narrowing by x needs the query inside the grid's x range
    const float gl = lx * lx, gr = (cell - lx) * (cell - lx);
    const int *__restrict__ row_begin = m.row_begin;
    const int *__restrict__ cell_begin = m.cell_begin;
    const float4 *__restrict__ sorted = m.sorted;
#pragma unroll 1
    for (int ro = 0; ro < 9; ro++) {
        // own row first, then the four face neighbours, then the four edge neighbours: the 5th-best distance tightens early
        const int rr = (int)((0x862075314ull >> (4 * ro)) & 15ull);
        const int dy = (rr % 3) - 1, dz = (rr / 3) - 1;
        const int y = cy + dy, z = cz + dz;
        if (y < 0 || y >= dimy || z < 0 || z >= dimz) continue;
        const float gy = dy < 0 ? ly : (dy > 0 ? cell - ly : 0.f);
        const float gz = dz < 0 ? lz : (dz > 0 ? cell - lz : 0.f);
        const float g2 = gy * gy + gz * gz;
        const float lim = bd[4] * 1.001f + 1e-4f;            // current 5th-best (or the bound), with margin
        if (g2 > lim) continue;
        const int ry = z * dimy + y;
        // cells of this row that intersect the ball: the own x cell always, a neighbour only when its face is inside
        int xa = x0, xb = x1;
        if (x_in) {
            xa = (cx - 1 >= 0 && g2 + gl <= lim) ? cx - 1 : cx;
            xb = (cx + 1 < dimx && g2 + gr <= lim) ? cx + 1 : cx;
        }
        const int r0 = __ldg(&row_begin[ry]), r1 = __ldg(&row_begin[ry + 1]);
        const int rb = __ldg(&cell_begin[ry * dimx + xa]), re = __ldg(&cell_begin[ry * dimx + xb + 1]);
        if (r1 <= r0) continue;                              // empty row: its cell table is not materialised
#pragma unroll 2
        for (int i = rb; i < re; i++) {
            const float4 p = __ldg(&sorted[i]);
            const float d = l2_simple(sx, sy, sz, p);
            const int oi = __float_as_int(p.w);
            if (d < bd[4] || (d == bd[4] && oi < bi[4])) knn5_insert(d, oi, i, bd, bi, bp);
        }
    }
}

// one branch-free insertion step of knn5_fast: (d, pos) sinks to its place in the ascending list, the rest shifts down, the
// smallest distance that is not kept goes to `rej`.  Written in PTX (setp / selp) because the compiler turns the C++
// selects into a tree of branches, which is exactly the divergence this step exists to avoid.
__device__ __forceinline__ void knn5_push(float d, int pos, float (&bd)[5], int (&bp)[5], float &rej)
{
    asm("{\n\t"
        ".reg .pred c0, c1, c2, c3, c4;\n\t"
        ".reg .f32 t;\n\t"
        ".reg .b32 u;\n\t"
        "setp.lt.f32 c0, %11, %0;\n\t"
        "setp.lt.f32 c1, %11, %1;\n\t"
        "setp.lt.f32 c2, %11, %2;\n\t"
        "setp.lt.f32 c3, %11, %3;\n\t"
        "setp.lt.f32 c4, %11, %4;\n\t"
        "selp.f32 t, %4, %11, c4;\n\t"                       // what drops out (or d itself when it is not kept)
        "min.f32 %10, %10, t;\n\t"
        "selp.f32 t, %11, %4, c4;\n\t selp.f32 %4, %3, t, c3;\n\t"
        "selp.b32 u, %12, %9, c4;\n\t selp.b32 %9, %8, u, c3;\n\t"
        "selp.f32 t, %11, %3, c3;\n\t selp.f32 %3, %2, t, c2;\n\t"
        "selp.b32 u, %12, %8, c3;\n\t selp.b32 %8, %7, u, c2;\n\t"
        "selp.f32 t, %11, %2, c2;\n\t selp.f32 %2, %1, t, c1;\n\t"
        "selp.b32 u, %12, %7, c2;\n\t selp.b32 %7, %6, u, c1;\n\t"
        "selp.f32 t, %11, %1, c1;\n\t selp.f32 %1, %0, t, c0;\n\t"
        "selp.b32 u, %12, %6, c1;\n\t selp.b32 %6, %5, u, c0;\n\t"
        "selp.f32 %0, %11, %0, c0;\n\t"
        "selp.b32 %5, %12, %5, c0;\n\t"
        "}"
        : "+f"(bd[0]), "+f"(bd[1]), "+f"(bd[2]), "+f"(bd[3]), "+f"(bd[4]),
          "+r"(bp[0]), "+r"(bp[1]), "+r"(bp[2]), "+r"(bp[3]), "+r"(bp[4]), "+f"(rej)
        : "f"(d), "r"(pos));
}

// per-warp scratch of the iteration kernel
template <int STAGE>                // STAGE = candidates a lane stages per pass
struct IterWarpSmem {
    union {
        float4 stage[STAGE][32];     // candidate points of the lanes' cell runs, copied by cp.async (lane-interleaved)
        struct {                     // after the search (the staging buffer is free then):
            float row[8][33];        //   Jacobian rows of the warp's 32 queries ({arx, ary, arz, cx, cy, cz, b, 1}; padded)
            double tot[32];          //   the 28 sums of a slot's LM step
        } fit;
    };
    int run_b[KNN_RUNS][32];         // narrowed cell runs of the neighbour rows, per lane: [begin, end) in `sorted`
    int run_e[KNN_RUNS][32];
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Ranks the candidates of the lane's runs [run_b[k], run_e[k]), k < n_runs.  A lane's candidate loads are all in flight
// at once: it issues one 16-byte cp.async per candidate into ITS column of the warp's staging buffer (no registers
// held, LDGSTS: ~8 cycles per warp-wide instruction whatever the depth), waits once and then ranks from shared memory.
// The first version walked the runs with plain loads, one ahead: the 32 lanes advance in lockstep, so every step
// waited for the slowest lane's L2 / DRAM round trip - ~60 dependent round trips, 60 us per launch even with 4 of 32
// slots live (profiles/r02_knnfit.md).
template <int STAGE>
__device__ __forceinline__ void knn5_rank_runs(const float4 *__restrict__ sorted, float sx, float sy, float sz, int n_runs,
                                               IterWarpSmem<STAGE> &ws, int lane, float (&bd)[5], int (&bp)[5], float &rej)
{
    int run = 0, i = 0, e = 0;
    if (n_runs > 0) { i = ws.run_b[0][lane]; e = ws.run_e[0][lane]; }
    while (run < n_runs) {                                   // passes of <= STAGE candidates
        int fr = run, fi = i, fe = e, cnt = 0;
        while (fr < n_runs && cnt < STAGE) {
            cp_async16(&ws.stage[cnt][lane], &sorted[fi]);
            cnt++; fi++;
            if (fi >= fe) { fr++; if (fr < n_runs) { fi = ws.run_b[fr][lane]; fe = ws.run_e[fr][lane]; } }
        }
        cp_async_wait_all();
        for (int k = 0; k < cnt; k++) {                      // same progression: i is the candidate's position
            const float4 p = ws.stage[k][lane];
            knn5_push(l2_simple(sx, sy, sz, p), i, bd, bp, rej);
            i++;
            if (i >= e) { run++; if (run < n_runs) { i = ws.run_b[run][lane]; e = ws.run_e[run][lane]; } }
        }
    }
}

// exact 5 nearest of (sx,sy,sz) with squared distance < bound_sq among the map points, distance keys only (see knn5_fast):
//   A  the query's own (y,z) row of cells: its run of <= 3 cells, narrowed when the bound is already tight;
//   B  the 8 neighbour rows against the 5th-best distance after A: pruned rows cost nothing, the surviving ones get
//      their run narrowed to the cells that intersect the ball; all their index loads are issued together;
//   C  the surviving runs ranked as one flattened list.
template <int STAGE>
__device__ __forceinline__ void knn5_walk(const MapIndexView &m, float sx, float sy, float sz, float bound_sq,
                                          IterWarpSmem<STAGE> &ws, int lane, float (&bd)[5], int (&bp)[5], float &rej,
                                          long long *tk = nullptr)
{
    const GridDesc *g = m.desc;
    const int dimx = g->dim[0], dimy = g->dim[1], dimz = g->dim[2];
    const float inv = g->inv_cell, cell = g->cell;
    const float fx = (sx - g->org[0]) * inv, fy = (sy - g->org[1]) * inv, fz = (sz - g->org[2]) * inv;
    const int cx = (int)floorf(fx), cy = (int)floorf(fy), cz = (int)floorf(fz);
#pragma unroll
    for (int k = 0; k < 5; k++) { bd[k] = bound_sq; bp[k] = -1; }
    rej = __int_as_float(0x7f800000);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, dimx - 1);
    if (x0 > x1) return;
    const float lx = (fx - (float)cx) * cell, ly = (fy - (float)cy) * cell, lz = (fz - (float)cz) * cell;
    const bool x_in = cx >= 0 && cx < dimx;
    const float gl = lx * lx, gr = (cell - lx) * (cell - lx);
    const int *__restrict__ row_begin = m.row_begin;
    const int *__restrict__ cell_begin = m.cell_begin;
    const float4 *__restrict__ sorted = m.sorted;
    // ---- A: own row
    if (cy >= 0 && cy < dimy && cz >= 0 && cz < dimz) {
        const float lim = bound_sq * 1.001f + 1e-4f;
        int xa = x0, xb = x1;
        if (x_in) {
            xa = (cx - 1 >= 0 && gl <= lim) ? cx - 1 : cx;
            xb = (cx + 1 < dimx && gr <= lim) ? cx + 1 : cx;
        }
        const int ry = cz * dimy + cy;
        const int r0 = __ldg(&row_begin[ry]), r1 = __ldg(&row_begin[ry + 1]);
        const int rb = __ldg(&cell_begin[ry * dimx + xa]), re = __ldg(&cell_begin[ry * dimx + xb + 1]);
        if (r1 > r0 && re > rb) {
            ws.run_b[0][lane] = rb; ws.run_e[0][lane] = re;
            knn5_rank_runs(sorted, sx, sy, sz, 1, ws, lane, bd, bp, rej);
        }
    }
    if (tk) tk[3] = clock64();
    // ---- B: neighbour rows that can still hold one of the five
    const float lim = bd[4] * 1.001f + 1e-4f;
    int n_runs = 0;
#pragma unroll
    for (int ro = 1; ro < 9; ro++) {
        const int rr = (int)((0x862075314ull >> (4 * ro)) & 15ull);          // face neighbours first, then edge neighbours
        const int dy = (rr % 3) - 1, dz = (rr / 3) - 1;
        const int y = cy + dy, z = cz + dz;
        const float gy = dy < 0 ? ly : (dy > 0 ? cell - ly : 0.f);
        const float gz = dz < 0 ? lz : (dz > 0 ? cell - lz : 0.f);
        const float g2 = gy * gy + gz * gz;
        if (y >= 0 && y < dimy && z >= 0 && z < dimz && g2 <= lim) {
            int xa = x0, xb = x1;
            if (x_in) {
                xa = (cx - 1 >= 0 && g2 + gl <= lim) ? cx - 1 : cx;
                xb = (cx + 1 < dimx && g2 + gr <= lim) ? cx + 1 : cx;
            }
            const int ry = z * dimy + y;
            const int r0 = __ldg(&row_begin[ry]), r1 = __ldg(&row_begin[ry + 1]);
            const int rb = __ldg(&cell_begin[ry * dimx + xa]), re = __ldg(&cell_begin[ry * dimx + xb + 1]);
            if (r1 > r0 && re > rb) { ws.run_b[n_runs][lane] = rb; ws.run_e[n_runs][lane] = re; n_runs++; }
        }
    }
    if (tk) tk[4] = clock64();
    // ---- C
    knn5_rank_runs(sorted, sx, sy, sz, n_runs, ws, lane, bd, bp, rej);
    if (tk) tk[5] = clock64();
}

// The LM step of a slot once all its chunks of this iteration are in: the chunk partials are added in chunk order (a
// fixed order, whatever warp produced them), then MO:1273-1326 on the 28 sums.  One warp.
__device__ __forceinline__ void batch_lm_step(const BatchReg &r, int n_chunks, int iter, const S2mParams &prm, int lane,
                                              double *s_tot)
{
    // lane L adds the partials of chunks L, L + 32, ... (all its loads independent, one round trip instead of a chain of
    // them), then a butterfly over the lanes: a fixed order whatever warp produced which chunk
#pragma unroll
    for (int g = 0; g < S2M_ACC; g += 14) {
        double acc[14];
#pragma unroll
        for (int k = 0; k < 14; k++) acc[k] = 0.0;
        for (int c = lane; c < n_chunks; c += 32) {
            const double *__restrict__ pp = r.partials + (size_t)c * S2M_ACC + g;
#pragma unroll
            for (int k = 0; k < 14; k++) acc[k] += __ldcg(pp + k);
        }
#pragma unroll
        for (int k = 0; k < 14; k++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(FULL, acc[k], o);
            if (lane == 0) s_tot[g + k] = acc[k];
        }
    }
    __syncwarp();
    S2mState *stw = r.st;
    if (lane == 0) { stw->ticket = 0; lm_solve(stw, s_tot, iter, prm, false); }
    __syncwarp();
    if (lane < 6) stw->cs[lane] = pose_trig(stw->T, lane);          // sin/cos of the new pose, one per lane
    __syncwarp();
}

// control word of a slot that hands out nothing more: a chunk field no slot reaches, and room for the late increments
// of warps that looked at the word just before it was set (they must not wrap it into a valid item)
constexpr unsigned CTL_DONE = 0xffff8000u;

template <int CTAS_PER_SM, int STAGE>
__global__ void __launch_bounds__(BATCH_ITER_THREADS, CTAS_PER_SM)
batch_lm_kernel(const BatchReg *__restrict__ regs, int B, S2mParams prm, int max_iter, int seed_bound, BatchQueue *__restrict__ queue)
{
    extern __shared__ __align__(16) unsigned char s_iter_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    IterWarpSmem<STAGE> &ws = reinterpret_cast<IterWarpSmem<STAGE> *>(s_iter_raw)[w];
    const float max_sq = prm.knn_max_sqdist;
    int ia, ib;
    pair_of(lane, ia, ib);
    long long tkb[12], tacc[12];
    long long *tk = queue->prof ? tkb : nullptr;             // debug (LLB_ITER_PROF): per-phase cycles of this warp's items
    int tk_items = 0;
    const long long tk_start = clock64();
    for (int k = 0; k < 12; k++) tacc[k] = 0;
    unsigned *ctl = queue->ctl;
    // warps of one CTA (and, with the usual round-robin placement, of one SM) start at the same slot: they share its map
    // in L1; a warp moves on to the next slot only when its slot has nothing to hand out
    int pref = blockIdx.x % B;
    unsigned idle = 0;
    for (;;) {
        if (tk) { const long long c = clock64(); for (int k = 0; k < 11; k++) tk[k] = c; }
        // ---- take a chunk: lanes look at 32 slots' control words at once
        int slot = -1; unsigned got = 0; int n_chunks = 0, nc = 0, ns = 0;
        for (int g0 = 0; g0 < B && slot < 0; g0 += 32) {
            const int sl = (pref + g0 + lane) % B;
            unsigned cw = CTL_DONE;
            if (g0 + lane < B) cw = __ldcg(&ctl[sl]);
            const int4 info = (g0 + lane < B) ? __ldcg(reinterpret_cast<const int4 *>(&queue->slot[sl])) : make_int4(0, 0, 1, 0);
            unsigned avail = __ballot_sync(FULL, (int)(cw & 0xffffu) < info.w);
            while (avail && slot < 0) {
                const int src = __ffs(avail) - 1;
                avail &= avail - 1;
                unsigned old = 0;
                if (lane == src) old = atomicAdd(&ctl[sl], 1u);
                old = __shfl_sync(FULL, old, src);
                const int nch = __shfl_sync(FULL, info.w, src);
                if ((int)(old & 0xffffu) < nch) {            // ours (a late increment of an exhausted word changes nothing)
                    slot = __shfl_sync(FULL, sl, src); got = old; n_chunks = nch;
                    nc = __shfl_sync(FULL, info.x, src); ns = __shfl_sync(FULL, info.y, src);
                }
            }
        }
        if (slot < 0) {                                      // nothing to hand out right now: LM steps in flight, or all done
            if (__ldcg(&queue->live) <= 0) break;
            __nanosleep(idle < 8 ? 100 : 400);
            idle++;
            continue;
        }
        idle = 0;
        pref = slot;
        __threadfence();                                     // acquire: the pose the slot's last LM step published
        const int iter = (int)(got >> 16);
        const int chunk = n_chunks - 1 - (int)(got & 0xffffu);   // the expensive chunks first (cost-ordered queries)
        const BatchReg &r = regs[slot];
        const S2mState *st = r.st;
        if (tk) tk[1] = clock64();
        const int nq = min(nc + ns, r.cap);
        const int nc_pad = (nc + 31) & ~31;
        const int s = (chunk << 5) + lane;
        const bool is_corner = s < nc_pad;
        const int j = is_corner ? s : s - nc_pad + nc;       // rank in the cost order (corner part, then surf part)
        const bool live = is_corner ? (s < nc && s < nq) : (j < nq);
        float v[8] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
        if (live) {
            const int q = (is_corner ? 0 : nc) + __ldg(&r.qperm[j]);
            const float4 *qp = is_corner ? &r.corner[q] : &r.surf[q - nc];
            const MapIndexView &m = is_corner ? r.cmap : r.smap;
            float sx, sy, sz;
            {
                const float4 po = __ldg(qp);
                associate_to_map(__ldcg(&st->cs[0]), __ldcg(&st->cs[1]), __ldcg(&st->cs[2]), __ldcg(&st->cs[3]), __ldcg(&st->cs[4]),
                                 __ldcg(&st->cs[5]), __ldcg(&st->T[3]), __ldcg(&st->T[4]), __ldcg(&st->T[5]), po, sx, sy, sz);
            }
            // the search only needs the mapped point: pose and query are re-read for the Jacobian row afterwards instead of
            // staying in registers (the compiler otherwise re-derives sx, sy, sz inside the candidate loop to save three)
            asm volatile("" : "+f"(sx), "+f"(sy), "+f"(sz));
            // bound of the search: the gate, or what the previous iteration's neighbours guarantee
            float bound = max_sq;
            if (seed_bound && iter > 0) {
                const float4 pv = __ldcg(&r.qprev[q]);       // written by another SM in the previous iteration: L2, not L1
                if (pv.w >= 0.f) {
                    const float ex = sx - pv.x, ey = sy - pv.y, ez = sz - pv.z;
                    const float reach = sqrtf(pv.w) + sqrtf(ex * ex + ey * ey + ez * ez);
                    bound = fminf(max_sq, reach * reach * 1.0001f + 1e-7f);
                }
            }
            if (tk) tk[2] = clock64();
            float bd[5], rej; int bp[5];
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {           // second pass (gate as the bound): never observed, keeps it exact
                knn5_walk(m, sx, sy, sz, bound, ws, lane, bd, bp, rej, tk);
                if (bp[4] >= 0 || bound >= max_sq) break;
                bound = max_sq;
            }
            bool found = bp[4] >= 0;
            if (found && (bd[0] == bd[1] || bd[1] == bd[2] || bd[2] == bd[3] || bd[3] == bd[4] || bd[4] == rej)) {
                int bi[5];                                   // equal distances: the exact (distance, original index) order
                knn5_bounded(m, sx, sy, sz, max_sq, bd, bi, bp);
                found = bp[4] >= 0;
            }
            if (seed_bound) __stcg(&r.qprev[q], make_float4(sx, sy, sz, found ? bd[4] : -1.f));
            if (found && (double)bd[4] < (double)max_sq) {   // MO:1101 / MO:1183
                float nx[5], ny[5], nz[5];
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const float4 p = __ldg(&m.sorted[bp[k]]);
                    nx[k] = p.x; ny[k] = p.y; nz[k] = p.z;
                }
                if (tk) tk[6] = clock64();
                float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
                const bool ok = is_corner ? corner_fit(nx, ny, nz, sx, sy, sz, coeff) : surf_fit(nx, ny, nz, sx, sy, sz, coeff);
                if (ok) {
                    const float4 po = __ldg(qp);
                    jacobian_row(__ldcg(&st->cs[0]), __ldcg(&st->cs[1]), __ldcg(&st->cs[2]), __ldcg(&st->cs[3]), __ldcg(&st->cs[4]),
                                 __ldcg(&st->cs[5]), po.x, po.y, po.z, coeff, v);
                }
            }
        }
        // ---- the 28 fp64 products of the chunk's rows, lane k = product k, rows in lane order
        __syncwarp();
        if (tk) tk[7] = clock64();
#pragma unroll
        for (int k = 0; k < 8; k++) ws.fit.row[k][lane] = v[k];
        __syncwarp();
        if (lane < S2M_ACC) {
            double acc = 0.0;
#pragma unroll 8
            for (int rr = 0; rr < 32; rr++) acc += (double)ws.fit.row[ia][rr] * (double)ws.fit.row[ib][rr];
            __stcg(&r.partials[(size_t)chunk * S2M_ACC + lane], acc);
        }
        // ---- the warp that delivers the LAST chunk of (slot, iteration) performs the LM step and releases the next iteration
        if (tk) tk[8] = clock64();
        __threadfence();
        __syncwarp();
        unsigned done = 0;
        if (lane == 0) done = atomicAdd(&r.st->ticket, 1u);
        done = __shfl_sync(FULL, done, 0);
        if (tk) tk[9] = clock64();
        if (done == (unsigned)(n_chunks - 1)) {
            __threadfence();
            __syncwarp();
            batch_lm_step(r, n_chunks, iter, prm, lane, ws.fit.tot);
            const bool more = !__ldcg(&r.st->converged) && iter + 1 < max_iter;      // MO:1336, MO:1344-1345
            __threadfence();                                 // release: pose, sin/cos and flags before the control word
            __syncwarp();
            if (lane == 0) {
                if (more) atomicExch(&ctl[slot], (unsigned)(iter + 1) << 16);
                else { atomicExch(&ctl[slot], CTL_DONE); atomicSub(&queue->live, 1); }
            }
            if (tk) tk[10] = clock64();
        }
        if (tk) {                                            // a phase this lane skipped keeps the previous stamp (0 cycles)
            for (int k = 1; k < 11; k++) if (tk[k] < tk[k - 1]) tk[k] = tk[k - 1];
            for (int k = 0; k < 10; k++) tacc[k] += tk[k + 1] - tk[k];
            tk_items++;
        }
    }
    if (tk && lane == 0) {
        long long *o = queue->prof + (size_t)(blockIdx.x * ITER_NW + w) * 12;
        for (int k = 0; k < 10; k++) o[k] = tacc[k];
        o[10] = clock64() - tk_start; o[11] = tk_items;
    }
}

__global__ void batch_collect_kernel(const BatchReg *__restrict__ regs, int B, BatchResult *__restrict__ out,
                                     BatchQueue *__restrict__ queue)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) queue->live = 0;                             // (every slot is done here; the prepare kernel counts them again)
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

void launch_batch_prepare(const BatchReg *regs, const float *poses_dev, int B, const S2mParams &prm, int max_iter, BatchQueue *queue,
                          cudaStream_t s)
{
    batch_prepare_kernel<<<div_up(B, 128), 128, 0, s>>>(regs, poses_dev, B, prm, max_iter, queue);
    LLB_CUDA(cudaGetLastError());
}

void launch_batch_qsort(const BatchReg *regs, int B, int cap, cudaStream_t s)
{
    const int bytes = cap * 12;
    if (bytes > 48 * 1024)                                   // opt-in above the default limit (idempotent, per device)
        LLB_CUDA(cudaFuncSetAttribute(batch_qsort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 12));
    static const int by_count = getenv("LLB_QS_BY_COUNT") ? atoi(getenv("LLB_QS_BY_COUNT")) : 1;
    batch_qsort_kernel<<<dim3(2, B), QS_THREADS, bytes, s>>>(regs, cap, by_count);
    LLB_CUDA(cudaGetLastError());
}

// variants of the registration kernel: (CTAs of 4 warps per SM, candidates staged per lane and pass); more resident warps
// hide more of the dependent round trips of an item, fewer leave more registers / staging per warp
static int batch_lm_variant()
{
    static const int v = getenv("LLB_ITER_OCC") ? atoi(getenv("LLB_ITER_OCC")) : BATCH_ITER_CTAS_PER_SM;
    return v;
}

template <int OCC, int STAGE>
static int lm_grid_of(int sms)
{
    int per_sm = 0;
    const int smem = (int)sizeof(IterWarpSmem<STAGE>) * ITER_NW;
    LLB_CUDA(cudaFuncSetAttribute(batch_lm_kernel<OCC, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    LLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, batch_lm_kernel<OCC, STAGE>, BATCH_ITER_THREADS, smem));
    return std::max(1, sms * std::max(1, per_sm));
}

template <int OCC, int STAGE>
static void lm_launch_of(const BatchReg *regs, int B, int grid, const S2mParams &prm, int max_iter, int seed, BatchQueue *queue, cudaStream_t s)
{
    const int smem = (int)sizeof(IterWarpSmem<STAGE>) * ITER_NW;
    batch_lm_kernel<OCC, STAGE><<<grid, BATCH_ITER_THREADS, smem, s>>>(regs, B, prm, max_iter, seed, queue);
}

// CTAs of the persistent registration kernel: all of them must be co-resident (idle warps poll for released work)
int batch_lm_grid()
{
    int dev = 0, sms = 0;
    LLB_CUDA(cudaGetDevice(&dev));
    LLB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int grid = 0;
    switch (batch_lm_variant()) {
    case 6: grid = lm_grid_of<6, 12>(sms); break;
    case 8: grid = lm_grid_of<8, 8>(sms); break;
    default: grid = lm_grid_of<4, 20>(sms); break;
    }
    // two thirds of what fits: the kernel is latency-bound, and a second batch's kernels (bench.py alternates two batches
    // per GPU) find room beside it; measured 78.5k registrations/s vs 66.5k with the full grid and 76.8k with half of it
    grid = std::max(1, grid * 2 / 3);
    if (getenv("LLB_ITER_GRID")) grid = std::max(1, std::min(grid * 3 / 2, atoi(getenv("LLB_ITER_GRID"))));
    return grid;
}

void launch_batch_lm(const BatchReg *regs, int B, int grid, const S2mParams &prm, int max_iter, BatchQueue *queue, cudaStream_t s)
{
    static const int seed = getenv("LLB_KNN_SEED_BOUND") ? atoi(getenv("LLB_KNN_SEED_BOUND")) : 1;
    if (max_iter > 0xfff0) throw std::runtime_error("s2m_max_iterations exceeds the batch control word");
    switch (batch_lm_variant()) {
    case 6: lm_launch_of<6, 12>(regs, B, grid, prm, max_iter, seed, queue, s); break;
    case 8: lm_launch_of<8, 8>(regs, B, grid, prm, max_iter, seed, queue, s); break;
    default: lm_launch_of<4, 20>(regs, B, grid, prm, max_iter, seed, queue, s); break;
    }
    LLB_CUDA(cudaGetLastError());
}

void launch_batch_collect(const BatchReg *regs, int B, BatchResult *out, BatchQueue *queue, cudaStream_t s)
{
    batch_collect_kernel<<<div_up(B, 128), 128, 0, s>>>(regs, B, out, queue);
    LLB_CUDA(cudaGetLastError());
}

}  // namespace llb
