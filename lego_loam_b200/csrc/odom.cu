// odom.cu — K5: both LM loops of updateTransformation (FA:1666-1695) in ONE persistent launch: an 8-CTA thread-block
// cluster (odom_kernel; every CTA runs the deterministic iteration loop redundantly on its own shared-memory copy of
// the state, the correspondence search of the refresh iterations is split by feature over the CTAs between two cluster
// barriers) or, batched, one CTA per slot (odom_batch_kernel).  One warp per feature point: the exact 1-NN over the
// previous sweep goes through a uniform-grid index built at llb_odom_set_last (cells of the 5 m gate, lexicographic
// (distance, index) minimum), the +-2.5-ring neighbour scans of FA:1061-1099 / FA:1172-1220 are evaluated 32 candidates
// at a time with ballot logic that reproduces the sequential "first strict minimum before the first break" semantics,
// and lane k accumulates the k-th product of the 3x3 normal equations in fp64.  Quirk C20 (stale trees when the last
// sweep has exactly 10 / 100 points, FA:1668 vs FA:1785) is reproduced by keeping the previous index in that case.
#include "odom.cuh"
#include "linalg.cuh"
#include "glibc_sincosf.cuh"
#include "grid_index.cuh"
#include <cmath>
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace llb {

namespace {

constexpr int OD_THREADS = 1024;
constexpr int OD_NW = OD_THREADS / 32;
constexpr int OD_ACC = 10;          // 6 upper-tri AtA + 3 AtB + count

struct OdomData {
    const float4 *sharp, *flat, *cornerLast, *surfLast;
    int nsharp, nflat, ncl, nsl;
    float *cInd1, *cInd2, *sInd1, *sInd2, *sInd3;
    float4 *dbg_coeff; int *dbg_valid;
    MapIndexView cgrid, sgrid;          // uniform-grid indices of the previous sweep's clouds (sorted == nullptr: none)
};

__device__ __forceinline__ float sqdist_ref(const float4 &a, float x, float y, float z)
{
    // FA:1067-1072: (a.x - x)^2 + (a.y - y)^2 + (a.z - z)^2, left to right
    return (a.x - x) * (a.x - x) + (a.y - y) * (a.y - y) + (a.z - z) * (a.z - z);
}

// TransformToStart FA:860-883; sin / cos of the float angles are glibc's sinf / cosf (glibc_sincosf.cuh), as in the reference build
__device__ __forceinline__ void transform_to_start(const float *T, const float4 &pi, float &ox, float &oy, float &oz)
{
    const float s = 10 * (pi.w - (int)pi.w);
    const float rx = s * T[0], ry = s * T[1], rz = s * T[2];
    const float tx = s * T[3], ty = s * T[4], tz = s * T[5];
    const float crz = glibcm::cosf_(rz), srz = glibcm::sinf_(rz);
    const float crx = glibcm::cosf_(rx), srx = glibcm::sinf_(rx);
    const float cry = glibcm::cosf_(ry), sry = glibcm::sinf_(ry);

    const float x1 = crz * (pi.x - tx) + srz * (pi.y - ty);
    const float y1 = -srz * (pi.x - tx) + crz * (pi.y - ty);
    const float z1 = (pi.z - tz);
    const float y2 = crx * y1 + srx * z1;
    const float z2 = -srx * y1 + crx * z1;
    ox = cry * x1 - sry * z2;
    oy = y2;
    oz = sry * x1 + cry * z2;
}

// exact 1-NN (flann L2_Simple, ties -> smaller index); every lane gets the result
__device__ __forceinline__ void nn1_warp(const float4 *__restrict__ pts, int n, float x, float y, float z, int lane,
                                         float &best_d, int &best_i)
{
    float bd = __int_as_float(0x7f800000); int bi = INT_MAX;
    for (int j = lane; j < n; j += 32) {
        const float4 p = __ldg(&pts[j]);
        float diff = x - p.x; float d = diff * diff;
        diff = y - p.y; d += diff * diff;
        diff = z - p.z; d += diff * diff;
        if (d < bd) { bd = d; bi = j; }            // ascending j per lane: ties keep the smaller index
    }
    const unsigned db = __float_as_uint(bd);
    const unsigned mind = __reduce_min_sync(FULL, db);
    const unsigned ci = (db == mind) ? (unsigned)bi : 0xffffffffu;
    best_i = (int)__reduce_min_sync(FULL, ci);
    best_d = __uint_as_float(mind);
}

// The same through the uniform-grid index of the cloud (cells slightly larger than the gate radius sqrt(25) = 5 m,
// UT:125): the 3x3x3 cells around the query contain every point the reference could accept (FA:1057: d2 < 25), so the
// exact nearest neighbour among them is the kd-tree's answer whenever the answer is used; otherwise best_d >= gate.
// Ties: smaller ORIGINAL index, as nn1_warp.
__device__ __forceinline__ void nn1_grid_warp(const MapIndexView &m, float x, float y, float z, int lane,
                                              float &best_d, int &best_i)
{
    const GridDesc *g = m.desc;
    const int dimx = g->dim[0], dimy = g->dim[1], dimz = g->dim[2];
    const float inv = g->inv_cell;
    const int cx = grid_coord(x, g->org[0], inv), cy = grid_coord(y, g->org[1], inv), cz = grid_coord(z, g->org[2], inv);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, dimx - 1);
    int b0 = 0, e0 = 0;
    {
        const int k = lane < 9 ? lane : 0;
        const int yy = cy + (k % 3) - 1, zz = cz + (k / 3) - 1;
        if (lane < 9 && yy >= 0 && yy < dimy && zz >= 0 && zz < dimz && x0 <= x1) {
            const int ry = zz * dimy + yy;
            if (__ldg(&m.row_begin[ry + 1]) > __ldg(&m.row_begin[ry])) {
                b0 = __ldg(&m.cell_begin[ry * dimx + x0]);
                e0 = __ldg(&m.cell_begin[ry * dimx + x1 + 1]);
            }
        }
    }
    float bd = __int_as_float(0x7f800000); int bi = INT_MAX;
#pragma unroll 1
    for (int r = 0; r < 9; r++) {
        const int rb = __shfl_sync(FULL, b0, r), re = __shfl_sync(FULL, e0, r);
        for (int j = rb + lane; j < re; j += 32) {
            const float4 p = __ldg(&m.sorted[j]);
            float diff = x - p.x; float d = diff * diff;
            diff = y - p.y; d += diff * diff;
            diff = z - p.z; d += diff * diff;
            const int oi = __float_as_int(p.w);
            if (d < bd || (d == bd && oi < bi)) { bd = d; bi = oi; }
        }
    }
    const unsigned db = __float_as_uint(bd);
    const unsigned mind = __reduce_min_sync(FULL, db);
    const unsigned ci = (db == mind) ? (unsigned)bi : 0xffffffffu;
    best_i = (int)__reduce_min_sync(FULL, ci);
    best_d = __uint_as_float(mind);
}

// One direction of the neighbour scan.  Visits j = start, start+step, ... while in [lo, hi)
// and until the first element whose ring violates the bound; among visited elements with
// cls(j) == k (k = 0, 1) keeps the first strict minimum below bd[k].
//   ringBreak: +1 forward (break if ring >= scan + 3), -1 backward (break if ring <= scan - 3)
template <bool SURF>
__device__ __forceinline__ void window_scan(const float4 *__restrict__ pts, int start, int step, int lo, int hi,
                                            int closestScan, float x, float y, float z, int lane,
                                            float (&bd)[2], int (&bj)[2])
{
    // lane-local running minima over the chunks (strict "<": within a lane the earlier visited element wins); ONE
    // ballot per chunk finds the first break; the warp-wide choice is made once, after the loop
    float lbd[2] = { bd[0], bd[1] };
    int lbj[2] = { -1, -1 };
    // the walk is a chain of dependent "load 32 candidates -> look for the break" steps (a surf window spans ~2500
    // points = 80 steps of one load latency each): the loads of four steps are issued together; the ones behind the
    // break are simply not used
    constexpr int U = 4;
    bool done = false;
    for (int base = start; !done; base += 32 * U * step) {
        float4 pv[U]; bool inv[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = base + (u * 32 + lane) * step;
            inv[u] = (j >= lo) && (j < hi);
            pv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (inv[u]) pv[u] = __ldg(&pts[j]);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (done) break;
            const int j = base + (u * 32 + lane) * step;
            const float4 p = pv[u];
            const bool in = inv[u];
            const int ring = (int)p.w;
            const bool brk = !in || (step > 0 ? ring >= closestScan + 3 : ring <= closestScan - 3);
            const unsigned bmask = __ballot_sync(FULL, brk);
            const int fb = bmask ? (__ffs(bmask) - 1) : 32;
            const bool vis = lane < fb;
            const float d = sqdist_ref(p, x, y, z);
            int cls;
            if (SURF) cls = (step > 0) ? ((ring <= closestScan) ? 0 : 1) : ((ring >= closestScan) ? 0 : 1);
            else cls = (step > 0) ? ((ring > closestScan) ? 0 : -1) : ((ring < closestScan) ? 0 : -1);
#pragma unroll
            for (int k = 0; k < (SURF ? 2 : 1); k++)
                if (vis && cls == k && d < lbd[k]) { lbd[k] = d; lbj[k] = j; }
            if (bmask) done = true;
        }
    }
    // the sequential loop keeps the FIRST strict minimum below the incoming bd[k]: smallest distance, and among equal
    // distances the element visited first (smallest j going forward, largest j going backward)
#pragma unroll
    for (int k = 0; k < (SURF ? 2 : 1); k++) {
        const bool cand = lbj[k] >= 0;
        const unsigned db = cand ? __float_as_uint(lbd[k]) : 0xffffffffu;       // distances are >= 0: bit order == value order
        const unsigned mind = __reduce_min_sync(FULL, db);
        if (mind != 0xffffffffu) {                                               // warp-uniform
            const bool win = cand && db == mind;
            const unsigned order = win ? (unsigned)(step > 0 ? lbj[k] : (INT_MAX - lbj[k])) : 0xffffffffu;
            const unsigned first = __reduce_min_sync(FULL, order);
            bd[k] = __uint_as_float(mind);
            bj[k] = step > 0 ? (int)first : INT_MAX - (int)first;
        }
    }
}
struct Trig { float srx, crx, sry, cry, srz, crz, tx, ty, tz; };

// 3x3 LM step shared by both solvers (FA:1324-1376 / FA:1425-1477).  Returns `more`.
// T, deg and matP are the CTA's shared-memory copies of the persistent state (loaded once per launch, mirrored to the
// global OdomState with plain stores): the sequential iterations never wait for a global load of their own results
__device__ int solve3(OdomState *st, bool writer, float *T, int *deg_s, float *matP, const double *sum, int iter, int which,
                      const OdomParams &prm)
{
    float AtA[9], AtB[3], A[9], B[3], X[3];
    AtA[0] = (float)sum[0]; AtA[1] = AtA[3] = (float)sum[1]; AtA[2] = AtA[6] = (float)sum[2];
    AtA[4] = (float)sum[3]; AtA[5] = AtA[7] = (float)sum[4]; AtA[8] = (float)sum[5];
    AtB[0] = (float)sum[6]; AtB[1] = (float)sum[7]; AtB[2] = (float)sum[8];
    for (int i = 0; i < 9; i++) A[i] = AtA[i];
    for (int i = 0; i < 3; i++) B[i] = AtB[i];
    cv_solve_qr<3, 3>(A, B, X);
    if (iter == 0) {
        float E[3], V[9], V2[9], Vinv[9];
        cv_eigen3(AtA[0], AtA[1], AtA[2], AtA[4], AtA[5], AtA[8], E, V);
        for (int i = 0; i < 9; i++) V2[i] = V[i];
        int deg = 0;
        for (int i = 2; i >= 0; i--) {
            if (E[i] < prm.degeneracy_thresh) {
                for (int j = 0; j < 3; j++) V2[i * 3 + j] = 0.f;
                deg = 1;
            } else break;
        }
        *deg_s = deg; if (writer) st->is_degenerate = deg;
        cv_inv3(V, Vinv);
        cv_gemm<3, 3, 3>(Vinv, V2, matP);
        if (writer) for (int i = 0; i < 9; i++) st->matP[i] = matP[i];
    }
    if (*deg_s) {
        float X2[3] = { X[0], X[1], X[2] };
        cv_gemm<3, 3, 1>(matP, X2, X);
    }
    float deltaR, deltaT;
    if (which == 0) {                 // surf: rx, rz, ty
        T[0] += X[0]; T[2] += X[1]; T[4] += X[2];
        double r0 = (double)X[0] * 180.0 / 3.14159265358979323846, r1 = (double)X[1] * 180.0 / 3.14159265358979323846;
        double t0 = (double)(X[2] * 100);
        deltaR = (float)sqrt(r0 * r0 + r1 * r1);
        deltaT = (float)sqrt(t0 * t0);
    } else {                          // corner: ry, tx, tz
        T[1] += X[0]; T[3] += X[1]; T[5] += X[2];
        double r0 = (double)X[0] * 180.0 / 3.14159265358979323846;
        double t0 = (double)(X[1] * 100), t1 = (double)(X[2] * 100);
        deltaR = (float)sqrt(r0 * r0);
        deltaT = (float)sqrt(t0 * t0 + t1 * t1);
    }
    for (int i = 0; i < 6; i++) { if (isnan(T[i])) T[i] = 0.f; if (writer) st->T[i] = T[i]; }   // C14
    if ((double)deltaR < (double)prm.converge_deg && (double)deltaT < (double)prm.converge_cm) return 0;
    return 1;
}

// mode 0: full updateTransformation; mode 1: one surf iteration `iter0`; mode 2: one corner iteration
//
// Per iteration, over tiles of 1024 features:
//   P1 thread/feature : TransformToStart (six fp64 sin/cos per point, FA:871-881) -> shared memory
//   P2 warp/feature   : only when iter % 5 == 0 (C4): exact 1-NN + the two ring-window scans -> index arrays
//   P3 thread/feature : line / plane coefficients from the stored indices, weight, Jacobian row -> shared memory
//   P4 lane k of warp w accumulates product k of the 10 normal-equation terms over rows w, w+32, ...
// then a fixed-order reduction and the 3x3 LM step by thread 0.
// CL > 1: the kernel runs as a thread-block CLUSTER of CL CTAs.  Every CTA executes the whole (deterministic) iteration
// loop redundantly on its own shared-memory copy of the state, so all control flow stays in lockstep without any
// communication; only the correspondence SEARCH of the refresh iterations (P2: 1-NN + ring-window scans, issue-bound
// on one SM: 200k cycles for 192 surf features) is split by feature over the CTAs, its results go to the global
// index arrays every CTA reads in P3 anyway, and one cluster barrier makes them visible.  CTA 0 alone writes OdomState.
template <int CL>
__device__ __forceinline__ void odom_body(const OdomParams &prm, const OdomData &dat, OdomState *__restrict__ st, int mode,
                                          int iter0)
{
    int crank = 0;
    if (CL > 1) crank = (int)cg::this_cluster().block_rank();
    const bool writer = crank == 0;
    __shared__ double s_acc[OD_NW][OD_ACC];
    __shared__ double s_tot[OD_ACC];
    __shared__ float s_T[6];
    __shared__ float s_matP[9];
    __shared__ int s_deg;
    __shared__ Trig s_trig;
    __shared__ int s_more;
    __shared__ float s_sel[3][OD_THREADS];                  // de-skewed feature of the tile
    __shared__ float s_row[5][OD_THREADS];                  // J0 J1 J2 b valid

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    if (tid == 0 && writer) {
        st->skipped = (dat.ncl < 10 || dat.nsl < 100) ? 1 : 0;               // FA:1668
        if (mode == 0) { st->iters[0] = st->iters[1] = 0; st->converged[0] = st->converged[1] = 0; }
        st->n_corr = 0; st->more = 1;
    }
    if (dat.ncl < 10 || dat.nsl < 100) return;
    if (tid == 0) {                                                          // persistent state -> shared memory, once
        for (int i = 0; i < 6; i++) s_T[i] = st->T[i];
        for (int i = 0; i < 9; i++) s_matP[i] = st->matP[i];
        s_deg = st->is_degenerate;
    }
    if (CL > 1) cg::this_cluster().sync();                                   // nobody overwrites the state before all have read it

    const int ia_tab[10] = { 0, 0, 0, 1, 1, 2, 0, 1, 2, 4 };
    const int ib_tab[10] = { 0, 1, 2, 1, 2, 2, 3, 3, 3, 4 };
    const int ia = lane < OD_ACC ? ia_tab[lane] : 0, ib = lane < OD_ACC ? ib_tab[lane] : 0;

    for (int which = 0; which < 2; which++) {
        if (mode == 1 && which != 0) continue;
        if (mode == 2 && which != 1) continue;
        const int nq = which == 0 ? dat.nflat : dat.nsharp;
        const float4 *qpts = which == 0 ? dat.flat : dat.sharp;
        const float4 *last = which == 0 ? dat.surfLast : dat.cornerLast;
        const int nlast = which == 0 ? dat.nsl : dat.ncl;
        const int fwdEnd = min(nq, nlast);                                   // quirk C1

        const int it_begin = mode == 0 ? 0 : iter0, it_end = mode == 0 ? prm.max_iter : iter0 + 1;
        for (int iter = it_begin; iter < it_end; iter++) {
            __syncthreads();
            if (tid == 0) { s_trig.tx = s_T[3]; s_trig.ty = s_T[4]; s_trig.tz = s_T[5]; }
            __syncthreads();
            if (tid < 6) {                                                   // six threads, one sinf / cosf each
                const float a = s_T[tid >> 1];
                const float v = (tid & 1) ? glibcm::cosf_(a) : glibcm::sinf_(a);
                if (tid == 0) s_trig.srx = v; if (tid == 1) s_trig.crx = v;
                if (tid == 2) s_trig.sry = v; if (tid == 3) s_trig.cry = v;
                if (tid == 4) s_trig.srz = v; if (tid == 5) s_trig.crz = v;
            }
            __syncthreads();
            const float srx = s_trig.srx, crx = s_trig.crx, sry = s_trig.sry, cry = s_trig.cry,
                        srz = s_trig.srz, crz = s_trig.crz, tx = s_trig.tx, ty = s_trig.ty, tz = s_trig.tz;
            double acc = 0.0;

            for (int t0 = 0; t0 < nq; t0 += OD_THREADS) {
                const int tn = min(OD_THREADS, nq - t0);
                // ---- P1
                float4 pi = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tid < tn) {
                    pi = __ldg(&qpts[t0 + tid]);
                    float x0, y0, z0;
                    transform_to_start(s_T, pi, x0, y0, z0);
                    s_sel[0][tid] = x0; s_sel[1][tid] = y0; s_sel[2][tid] = z0;
                }
                __syncthreads();
                // ---- P2
                if (iter % 5 == 0) {                                          // C4
                    // no CTA may overwrite indices another CTA is still reading in P3 of an earlier iteration
                    if (CL > 1) cg::this_cluster().sync();
                    for (int s = w + OD_NW * crank; s < tn; s += OD_NW * CL) {
                        const int i = t0 + s;
                        const float x0 = s_sel[0][s], y0 = s_sel[1][s], z0 = s_sel[2][s];
                        float d1; int closest;
                        const MapIndexView &gv = which == 0 ? dat.sgrid : dat.cgrid;
                        if (gv.sorted) nn1_grid_warp(gv, x0, y0, z0, lane, d1, closest);
                        else nn1_warp(last, nlast, x0, y0, z0, lane, d1, closest);
                        float bd[2] = { prm.nearest_sqdist, prm.nearest_sqdist };
                        int bj[2] = { -1, -1 };
                        if (d1 < prm.nearest_sqdist && closest < nlast) {         // (closest >= nlast: stale index, C20)
                            const int closestScan = (int)__ldg(&last[closest]).w;
                            if (which == 0) {
                                window_scan<true>(last, closest + 1, +1, 0, fwdEnd, closestScan, x0, y0, z0, lane, bd, bj);
                                window_scan<true>(last, closest - 1, -1, 0, nlast, closestScan, x0, y0, z0, lane, bd, bj);
                            } else {
                                window_scan<false>(last, closest + 1, +1, 0, fwdEnd, closestScan, x0, y0, z0, lane, bd, bj);
                                window_scan<false>(last, closest - 1, -1, 0, nlast, closestScan, x0, y0, z0, lane, bd, bj);
                            }
                        } else closest = -1;
                        if (lane == 0) {
                            if (which == 0) { dat.sInd1[i] = (float)closest; dat.sInd2[i] = (float)bj[0]; dat.sInd3[i] = (float)bj[1]; }
                            else { dat.cInd1[i] = (float)closest; dat.cInd2[i] = (float)bj[0]; }
                        }
                    }
                    if (CL > 1) { __threadfence(); cg::this_cluster().sync(); }   // every CTA's share of the indices is visible
                    else __syncthreads();
                }
                // ---- P3
                {
                    float v[5] = { 0.f, 0.f, 0.f, 0.f, 0.f };
                    if (tid < tn) {
                        const int i = t0 + tid;
                        const float x0 = s_sel[0][tid], y0 = s_sel[1][tid], z0 = s_sel[2][tid];
                        bool ok = false;
                        float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (which == 0) {
                            const float f1 = dat.sInd1[i], f2 = dat.sInd2[i], f3 = dat.sInd3[i];
                            if (f2 >= 0 && f3 >= 0) {
                                const float4 t1 = __ldg(&last[(int)f1]), t2 = __ldg(&last[(int)f2]), t3 = __ldg(&last[(int)f3]);
                                float pa = (t2.y - t1.y) * (t3.z - t1.z) - (t3.y - t1.y) * (t2.z - t1.z);
                                float pb = (t2.z - t1.z) * (t3.x - t1.x) - (t3.z - t1.z) * (t2.x - t1.x);
                                float pc = (t2.x - t1.x) * (t3.y - t1.y) - (t3.x - t1.x) * (t2.y - t1.y);
                                float pd = -(pa * t1.x + pb * t1.y + pc * t1.z);
                                const float ps = sqrtf(pa * pa + pb * pb + pc * pc);
                                pa /= ps; pb /= ps; pc /= ps; pd /= ps;
                                const float pd2 = pa * x0 + pb * y0 + pc * z0 + pd;
                                float s = 1;
                                if (iter >= 5)
                                    s = (float)(1.0 - 1.8 * (double)fabsf(pd2) / (double)sqrtf(sqrtf(x0 * x0 + y0 * y0 + z0 * z0)));
                                if ((double)s > 0.1 && pd2 != 0) {
                                    coeff = make_float4(s * pa, s * pb, s * pc, s * pd2);
                                    ok = true;
                                }
                            }
                        } else {
                            const float f1 = dat.cInd1[i], f2 = dat.cInd2[i];
                            if (f2 >= 0) {
                                const float4 t1 = __ldg(&last[(int)f1]), t2 = __ldg(&last[(int)f2]);
                                const float x1 = t1.x, y1 = t1.y, z1 = t1.z, x2 = t2.x, y2 = t2.y, z2 = t2.z;
                                const float m11 = ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1));
                                const float m22 = ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1));
                                const float m33 = ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1));
                                const float a012 = sqrtf(m11 * m11 + m22 * m22 + m33 * m33);
                                const float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
                                const float la = ((y1 - y2) * m11 + (z1 - z2) * m22) / a012 / l12;
                                const float lb = -((x1 - x2) * m11 - (z1 - z2) * m33) / a012 / l12;
                                const float lc = -((x1 - x2) * m22 + (y1 - y2) * m33) / a012 / l12;
                                const float ld2 = a012 / l12;
                                float s = 1;
                                if (iter >= 5) s = (float)(1.0 - 1.8 * (double)fabsf(ld2));
                                if ((double)s > 0.1 && ld2 != 0) {
                                    coeff = make_float4(s * la, s * lb, s * lc, s * ld2);
                                    ok = true;
                                }
                            }
                        }
                        if (dat.dbg_coeff) { dat.dbg_coeff[i] = coeff; dat.dbg_valid[i] = ok ? 1 : 0; }
                        if (ok) {
                            if (which == 0) {                                         // FA:1291-1321
                                const float a1 = crx * sry * srz, a2 = crx * crz * sry, a3 = srx * sry, a4 = tx * a1 - ty * a2 - tz * a3;
                                const float a5 = srx * srz, a6 = crz * srx, a7 = ty * a6 - tz * crx - tx * a5;
                                const float a8 = crx * cry * srz, a9 = crx * cry * crz, a10 = cry * srx, a11 = tz * a10 + ty * a9 - tx * a8;
                                const float b1 = -crz * sry - cry * srx * srz, b2 = cry * crz * srx - sry * srz;
                                const float b5 = cry * crz - srx * sry * srz, b6 = cry * srz + crz * srx * sry;
                                const float c1 = -b6, c2 = b5, c3 = tx * b6 - ty * b5, c4 = -crx * crz, c5 = crx * srz, c6 = ty * c5 + tx * -c4;
                                const float c7 = b2, c8 = -b1, c9 = tx * -b2 - ty * -b1;
                                v[0] = (-a1 * pi.x + a2 * pi.y + a3 * pi.z + a4) * coeff.x
                                     + (a5 * pi.x - a6 * pi.y + crx * pi.z + a7) * coeff.y
                                     + (a8 * pi.x - a9 * pi.y - a10 * pi.z + a11) * coeff.z;
                                v[1] = (c1 * pi.x + c2 * pi.y + c3) * coeff.x
                                     + (c4 * pi.x - c5 * pi.y + c6) * coeff.y
                                     + (c7 * pi.x + c8 * pi.y + c9) * coeff.z;
                                v[2] = -b6 * coeff.x + c4 * coeff.y + b2 * coeff.z;
                            } else {                                                  // FA:1400-1422
                                const float b1 = -crz * sry - cry * srx * srz, b2 = cry * crz * srx - sry * srz, b3 = crx * cry,
                                            b4 = tx * -b1 + ty * -b2 + tz * b3;
                                const float b5 = cry * crz - srx * sry * srz, b6 = cry * srz + crz * srx * sry, b7 = crx * sry,
                                            b8 = tz * b7 - ty * b6 - tx * b5;
                                const float c5 = crx * srz;
                                v[0] = (b1 * pi.x + b2 * pi.y - b3 * pi.z + b4) * coeff.x
                                     + (b5 * pi.x + b6 * pi.y - b7 * pi.z + b8) * coeff.z;
                                v[1] = -b5 * coeff.x + c5 * coeff.y + b1 * coeff.z;
                                v[2] = b7 * coeff.x - srx * coeff.y - b3 * coeff.z;
                            }
                            v[3] = (float)(-0.05 * (double)coeff.w);
                            v[4] = 1.f;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 5; k++) s_row[k][tid] = v[k];
                }
                __syncthreads();
                // ---- P4
                if (lane < OD_ACC)
                    for (int r = w; r < tn; r += OD_NW) acc += (double)s_row[ia][r] * (double)s_row[ib][r];
                __syncthreads();
            }

            if (lane < OD_ACC) s_acc[w][lane] = acc;
            __syncthreads();
            if (tid < OD_ACC) {
                double s = 0.0;
                for (int k = 0; k < OD_NW; k++) s += s_acc[k][tid];
                s_tot[tid] = s;
            }
            __syncthreads();
            if (tid == 0) {
                const int n_corr = (int)s_tot[9];
                if (writer) { st->n_corr = n_corr; st->iters[which] = iter + 1; }
                int more = 1;
                if (n_corr >= prm.min_corr) {                                 // FA:1677 / FA:1690
                    more = solve3(st, writer, s_T, &s_deg, s_matP, s_tot, iter, which, prm);
                    if (!more && writer) st->converged[which] = 1;
                }
                if (writer) st->more = more;
                s_more = more;
            }
            __syncthreads();
            if (!s_more) break;
        }
    }
}

// one sweep pair (llb_ctx)
constexpr int OD_CLUSTER = 8;
__global__ void __cluster_dims__(OD_CLUSTER, 1, 1) __launch_bounds__(OD_THREADS, 1)
odom_kernel(OdomParams prm, OdomData dat, OdomState *__restrict__ st, int mode, int iter0)
{
    odom_body<OD_CLUSTER>(prm, dat, st, mode, iter0);
}

// B independent sweep pairs, one CTA each (llb_batch): the <= 50 sequential iterations of a pair cannot be spread
// over more than one CTA, but 148 SMs run 148 pairs side by side
__global__ void __launch_bounds__(OD_THREADS, 1)
odom_batch_kernel(OdomParams prm, const OdomBatchJob *__restrict__ jobs)
{
    const OdomBatchJob jb = jobs[blockIdx.x];
    OdomData dat;
    dat.sharp = jb.sharp; dat.flat = jb.flat; dat.cornerLast = jb.cornerLast; dat.surfLast = jb.surfLast;
    dat.nsharp = jb.nsharp; dat.nflat = jb.nflat; dat.ncl = jb.ncl; dat.nsl = jb.nsl;
    dat.cInd1 = jb.ind; dat.cInd2 = jb.ind + jb.cap; dat.sInd1 = jb.ind + 2 * jb.cap; dat.sInd2 = jb.ind + 3 * jb.cap;
    dat.sInd3 = jb.ind + 4 * jb.cap;
    dat.dbg_coeff = nullptr; dat.dbg_valid = nullptr;
    dat.cgrid = MapIndexView{ nullptr, nullptr, nullptr, nullptr }; dat.sgrid = dat.cgrid;   // brute-force 1-NN
    odom_body<1>(prm, dat, jb.st, 0, 0);
}

__global__ void odom_state_init_kernel(OdomState *st)
{
    for (int i = 0; i < 6; i++) st->T[i] = 0.f;
    for (int i = 0; i < 9; i++) st->matP[i] = 0.f;
    st->is_degenerate = 0; st->iters[0] = st->iters[1] = 0; st->converged[0] = st->converged[1] = 0;
    st->n_corr = 0; st->skipped = 0; st->more = 1;
}

__global__ void fill_float_kernel(float *p, int n, float v)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

void launch_odom_batch(const OdomParams &prm, const OdomBatchJob *jobs_dev, int count, cudaStream_t s)
{
    if (count <= 0) return;
    odom_batch_kernel<<<count, OD_THREADS, 0, s>>>(prm, jobs_dev);
    LLB_CUDA(cudaGetLastError());
}

__global__ void odom_batch_set_pose_kernel(const OdomBatchJob *__restrict__ jobs, const float *__restrict__ poses, int count)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= count) return;
    OdomState *st = jobs[b].st;
    for (int i = 0; i < 6; i++) st->T[i] = poses[6 * b + i];
}

void launch_odom_batch_set_pose(const OdomBatchJob *jobs_dev, const float *poses_dev, int count, cudaStream_t s)
{
    odom_batch_set_pose_kernel<<<div_up(count, 128), 128, 0, s>>>(jobs_dev, poses_dev, count);
    LLB_CUDA(cudaGetLastError());
}

void launch_odom_state_init(OdomState *st, int count, cudaStream_t s)
{
    for (int i = 0; i < count; i++) odom_state_init_kernel<<<1, 1, 0, s>>>(st + i);
    LLB_CUDA(cudaGetLastError());
}

void launch_odom_fill(float *p, int n, float v, cudaStream_t s)
{
    fill_float_kernel<<<div_up(n, 256), 256, 0, s>>>(p, n, v);
    LLB_CUDA(cudaGetLastError());
}

void OdomSolver::init(const OdomParams &p)
{
    prm_ = p;
    state_.ensure(1);
    odom_state_init_kernel<<<1, 1>>>(state_.p);
    LLB_CUDA(cudaGetLastError());
}

void OdomSolver::release()
{
    state_.release(); cornerLast_.release(); surfLast_.release(); sharp_.release(); flat_.release();
    ind_.release(); dbg_coeff_.release(); dbg_valid_.release();
    if (grids_init_) { gridCorner_.release(); gridSurf_.release(); grids_init_ = false; grids_built_ = false; }
}

int OdomSolver::set_last(int ncl, int nsl, cudaStream_t s)
{
    ncl_ = ncl; nsl_ = nsl; last_set_ = true;
    // spatial index of the previous sweep's clouds (the reference rebuilds its two kd-trees here, FA:1615-1616 /
    // FA:1786-1787): cells of the gate radius, so the 27-cell neighbourhood holds every acceptable nearest neighbour
    // (first call: the tables are zeroed on the same stream the build runs on - on the default stream the build on a
    // non-blocking stream raced with them: an intermittently wrong first index, seen as a flaky first odometry test)
    if (!grids_init_) { gridCorner_.init(1 << 18, s); gridSurf_.init(1 << 18, s); grids_init_ = true; }
    // Quirk C20: the reference rebuilds its two kd-trees only when the new clouds have MORE than 10 / 100 points
    // (FA:1785-1788) while updateTransformation runs from 10 / 100 points on (FA:1668): with exactly 10 corner or 100
    // surf points (or a smaller cloud on the other side) the next matcher searches the PREVIOUS sweep's trees - which
    // hold their own copy of that sweep's points - and uses the returned indices in the NEW clouds.  The grid indices
    // keep a copy of their points too, so "do not rebuild" reproduces it; an index that points beyond the new cloud
    // (undefined behaviour in the reference) counts as "no neighbour" in the kernel.
    if (!(ncl > 10 && nsl > 100)) return 0;
    const int n = GridIndex::build_pair(gridCorner_, cornerLast_.p, nullptr, ncl, gridSurf_, surfLast_.p, nullptr, nsl,
                                        std::sqrt(prm_.nearest_sqdist), s);
    grids_built_ = true;
    return n;
}

void OdomSolver::ensure_work()
{
    const int need = std::max(std::max(nsharp_, nflat_), 1);
    if (need > cap_) {
        const int cap = need + 64;
        ind_.ensure((size_t)cap * 5);
        fill_float_kernel<<<div_up(cap * 5, 256), 256>>>(ind_.p, cap * 5, -1.f);
        LLB_CUDA(cudaGetLastError());
        LLB_CUDA(cudaDeviceSynchronize());
        dbg_coeff_.ensure(cap); dbg_valid_.ensure(cap);
        cap_ = cap;
    }
}

void OdomSolver::set_features(int nsharp, int nflat)
{
    nsharp_ = nsharp; nflat_ = nflat; feat_set_ = true;
    ensure_work();
}

static OdomData make_data(const float4 *sharp, const float4 *flat, const float4 *cl, const float4 *sl, int nsharp,
                          int nflat, int ncl, int nsl, float *ind, int cap, float4 *dbg_coeff, int *dbg_valid)
{
    OdomData d;
    d.cgrid = MapIndexView{ nullptr, nullptr, nullptr, nullptr }; d.sgrid = d.cgrid;
    d.sharp = sharp; d.flat = flat; d.cornerLast = cl; d.surfLast = sl;
    d.nsharp = nsharp; d.nflat = nflat; d.ncl = ncl; d.nsl = nsl;
    d.cInd1 = ind; d.cInd2 = ind + cap; d.sInd1 = ind + 2 * cap; d.sInd2 = ind + 3 * cap; d.sInd3 = ind + 4 * cap;
    d.dbg_coeff = dbg_coeff; d.dbg_valid = dbg_valid;
    return d;
}

int OdomSolver::optimize(const float *T, cudaStream_t s)
{
    LLB_CUDA(cudaMemcpyAsync(state_.p, T, 6 * sizeof(float), cudaMemcpyHostToDevice, s));   // OdomState starts with T[6]
    OdomData d = make_data(sharp_.p, flat_.p, cornerLast_.p, surfLast_.p, nsharp_, nflat_, ncl_, nsl_, ind_.p, cap_,
                           nullptr, nullptr);
    if (grids_built_) { d.cgrid = gridCorner_.view(); d.sgrid = gridSurf_.view(); }
    odom_kernel<<<OD_CLUSTER, OD_THREADS, 0, s>>>(prm_, d, state_.p, 0, 0);
    LLB_CUDA(cudaGetLastError());
    dbg_which_ = -1;
    return 1;
}

int OdomSolver::iterate(int which, const float *T, int iter, cudaStream_t s)
{
    LLB_CUDA(cudaMemcpyAsync(state_.p, T, 6 * sizeof(float), cudaMemcpyHostToDevice, s));
    OdomData d = make_data(sharp_.p, flat_.p, cornerLast_.p, surfLast_.p, nsharp_, nflat_, ncl_, nsl_, ind_.p, cap_,
                           dbg_coeff_.p, dbg_valid_.p);
    if (grids_built_) { d.cgrid = gridCorner_.view(); d.sgrid = gridSurf_.view(); }
    odom_kernel<<<OD_CLUSTER, OD_THREADS, 0, s>>>(prm_, d, state_.p, which == 0 ? 1 : 2, iter);
    LLB_CUDA(cudaGetLastError());
    dbg_which_ = which;
    return 1;
}

void OdomSolver::download_correspondences(std::vector<float4> &ori, std::vector<float4> &coeff, cudaStream_t s)
{
    ori.clear(); coeff.clear();
    if (dbg_which_ < 0) return;
    const int n = dbg_which_ == 0 ? nflat_ : nsharp_;
    if (n <= 0) return;
    std::vector<float4> q(n), c(n); std::vector<int> v(n);
    LLB_CUDA(cudaMemcpyAsync(q.data(), dbg_which_ == 0 ? flat_.p : sharp_.p, sizeof(float4) * n, cudaMemcpyDeviceToHost, s));
    LLB_CUDA(cudaMemcpyAsync(c.data(), dbg_coeff_.p, sizeof(float4) * n, cudaMemcpyDeviceToHost, s));
    LLB_CUDA(cudaMemcpyAsync(v.data(), dbg_valid_.p, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    LLB_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < n; i++)
        if (v[i]) { ori.push_back(q[i]); coeff.push_back(c[i]); }
}

void OdomSolver::download_search_ind(int which, std::vector<float> &i1, std::vector<float> &i2, std::vector<float> &i3,
                                     cudaStream_t s)
{
    // which: 0 = surf (pointSearchSurfInd1/2/3), 1 = corner (pointSearchCornerInd1/2), as everywhere else in the C ABI
    const int n = which == 0 ? nflat_ : nsharp_;
    i1.assign(n, -1.f); i2.assign(n, -1.f); i3.assign(n, -1.f);
    if (n <= 0) return;
    const float *b1 = which == 0 ? ind_.p + 2 * cap_ : ind_.p;
    const float *b2 = which == 0 ? ind_.p + 3 * cap_ : ind_.p + cap_;
    LLB_CUDA(cudaMemcpyAsync(i1.data(), b1, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
    LLB_CUDA(cudaMemcpyAsync(i2.data(), b2, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
    if (which == 0) LLB_CUDA(cudaMemcpyAsync(i3.data(), ind_.p + 4 * cap_, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
    LLB_CUDA(cudaStreamSynchronize(s));
}

}  // namespace llb
