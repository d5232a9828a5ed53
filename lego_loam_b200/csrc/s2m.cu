// s2m.cu — K3 (transform -> radius-bounded exact 5-NN -> line / plane fit -> residual +
// Jacobian row) and K4 (J^T J / J^T r reduction, LM step, degeneracy projection,
// convergence test) as ONE persistent cooperative kernel for the whole loop of
// scan2MapOptimization (MO:1336-1346): an iteration never returns to the host and never
// pays a launch.
//
// Work decomposition per iteration (each CTA owns a contiguous slice of the queries):
//   phase A  one WARP per query: the 9 contiguous cell runs around the query are read
//            cooperatively (lane = candidate, coalesced float4 loads, all loads of the
//            non-empty runs in flight before the first use), each lane keeps a private
//            sorted top-5 in registers and five rounds of warp-wide lexicographic
//            (distance, original index) arg-min give the exact 5-NN with the oracle's tie
//            rule; the 5 neighbours go to shared memory.
//   phase B  one THREAD per query: gate, 3x3 Jacobi eigen fit (corner) or 5x3 Householder
//            least squares (surf), residual, weight, Jacobian row -> shared memory.
//   phase C  lane k of warp w accumulates product k of the 28 normal-equation terms over
//            rows w, w+16, ... in fp64 (exact float x float products, like cv::gemm).
//   then     fixed-order CTA reduction -> global partials -> grid.sync -> CTA 0 adds the
//            partials in a fixed order, rounds to fp32 exactly like cv::gemm, and performs
//            the 6x6 LM step (QR solve, iteration-0 eigen degeneracy test, matP projection,
//            pose update, sin/cos of the new pose, convergence flag) -> grid.sync.
//
// Numerics: IEEE fp32 without contraction (-fmad=false) in the reference's association
// order; the double-promoted sub-expressions of SURVEY.md Appendix B are evaluated in
// fp64; pose sin/cos are (float)sin((double)x) (correctly rounded float).
#include "s2m.cuh"
#include "linalg.cuh"
#include "knn.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace llb {

namespace {

// 512 threads x 128 registers = one CTA per SM (an 80-register / two-CTA variant spilled and was slower,
// profiles/r01_s2m_phases.md); the grid is sized so that every SM gets a CTA before a CTA gets a second
// round of queries.
constexpr int S2M_THREADS = 512;
constexpr int S2M_CTAS_PER_SM = 1;
constexpr int S2M_NW = S2M_THREADS / 32;
constexpr int S2M_TILE = 64;             // queries a CTA stages per pass
constexpr int S2M_QPB = S2M_NW;          // target queries per CTA when sizing the grid (one per warp)

// cornerOptimization body for one query (MO:1102-1170).  Returns true when the row is accepted.
__device__ __forceinline__ bool corner_fit(const float (&nx)[5], const float (&ny)[5], const float (&nz)[5],
                                           float x0, float y0, float z0, float4 &coeff)
{
    float cx = 0, cy = 0, cz = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) { cx += nx[j]; cy += ny[j]; cz += nz[j]; }
    cx /= 5; cy /= 5; cz /= 5;

    float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        float ax = nx[j] - cx, ay = ny[j] - cy, az = nz[j] - cz;
        a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
        a22 += ay * ay; a23 += ay * az;
        a33 += az * az;
    }
    a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;

    float D1[3], V1[9];
    cv_eigen3(a11, a12, a13, a22, a23, a33, D1, V1);
    if (!(D1[0] > 3 * D1[1])) return false;

    float x1 = (float)((double)cx + 0.1 * (double)V1[0]);
    float y1 = (float)((double)cy + 0.1 * (double)V1[1]);
    float z1 = (float)((double)cz + 0.1 * (double)V1[2]);
    float x2 = (float)((double)cx - 0.1 * (double)V1[0]);
    float y2 = (float)((double)cy - 0.1 * (double)V1[1]);
    float z2 = (float)((double)cz - 0.1 * (double)V1[2]);

    float m11 = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1);
    float m22 = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1);
    float m33 = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
    float a012 = sqrtf(m11 * m11 + m22 * m22 + m33 * m33);
    float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
    float la = ((y1 - y2) * m11 + (z1 - z2) * m22) / a012 / l12;
    float lb = -((x1 - x2) * m11 - (z1 - z2) * m33) / a012 / l12;
    float lc = -((x1 - x2) * m22 + (y1 - y2) * m33) / a012 / l12;
    float ld2 = a012 / l12;

    float s = (float)(1.0 - 0.9 * (double)fabsf(ld2));
    coeff = make_float4(s * la, s * lb, s * lc, s * ld2);
    return (double)s > 0.1;
}

// surfOptimization body for one query (MO:1184-1223)
__device__ __forceinline__ bool surf_fit(const float (&nx)[5], const float (&ny)[5], const float (&nz)[5],
                                         float x0, float y0, float z0, float4 &coeff)
{
    float A0[15], B0[5] = { -1.f, -1.f, -1.f, -1.f, -1.f }, X0[3];
#pragma unroll
    for (int j = 0; j < 5; j++) { A0[3 * j] = nx[j]; A0[3 * j + 1] = ny[j]; A0[3 * j + 2] = nz[j]; }
    cv_solve_qr<5, 3>(A0, B0, X0);

    float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1;
    float ps = sqrtf(pa * pa + pb * pb + pc * pc);
    pa /= ps; pb /= ps; pc /= ps; pd /= ps;

    bool planeValid = true;
#pragma unroll
    for (int j = 0; j < 5; j++)
        if ((double)fabsf(pa * nx[j] + pb * ny[j] + pc * nz[j] + pd) > 0.2) planeValid = false;
    if (!planeValid) return false;

    float pd2 = pa * x0 + pb * y0 + pc * z0 + pd;
    float s = (float)(1.0 - 0.9 * (double)fabsf(pd2) / (double)sqrtf(sqrtf(x0 * x0 + y0 * y0 + z0 * z0)));
    coeff = make_float4(s * pa, s * pb, s * pc, s * pd2);
    return (double)s > 0.1;
}

__device__ void update_sincos(S2mState *st)
{
    st->cs[0] = (float)cos((double)st->T[0]); st->cs[1] = (float)sin((double)st->T[0]);
    st->cs[2] = (float)cos((double)st->T[1]); st->cs[3] = (float)sin((double)st->T[1]);
    st->cs[4] = (float)cos((double)st->T[2]); st->cs[5] = (float)sin((double)st->T[2]);
}

// LMOptimization tail MO:1273-1326 on the 28 reduced sums (one thread)
// Certificate that every eigenvalue of the symmetric matrix A (6x6, float) exceeds `bound`:
// Cholesky of (A - bound*I) in fp64 succeeds iff that matrix is positive definite.
__device__ bool all_eigenvalues_above(const float *A, double bound)
{
    double L[36];
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = (double)A[j * 6 + j] - bound;
#pragma unroll
        for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k];
        if (!(d > 0.0)) return false;
        const double r = sqrt(d);
        L[j * 6 + j] = r;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            double s = (double)A[i * 6 + j];
#pragma unroll
            for (int k = 0; k < j; k++) s -= L[i * 6 + k] * L[j * 6 + k];
            L[i * 6 + j] = s / r;
        }
    }
    return true;
}

// iteration-0 degeneracy analysis MO:1278-1299 on AtA: cv::eigen, zero the rows of the small
// eigenvalues, matP = V^-1 * V2.  ~80 us for one thread (Jacobi in local memory): only run when the
// cheap certificate below cannot rule degeneracy out, or when matP is asked for.
__device__ void degeneracy_full(S2mState *st, const float *AtA, float thresh)
{
    float A[36], E[6], V[36], V2[36], Vinv[36];
    for (int i = 0; i < 36; i++) A[i] = AtA[i];
    cv_eigen<6>(A, E, V);
    for (int i = 0; i < 36; i++) V2[i] = V[i];
    int deg = 0;
    for (int i = 5; i >= 0; i--) {
        if (E[i] < thresh) {
            for (int j = 0; j < 6; j++) V2[i * 6 + j] = 0.f;
            deg = 1;
        } else break;
    }
    st->is_degenerate = deg;
    cv_inv_lu<6>(V, Vinv);
    cv_gemm<6, 6, 6>(Vinv, V2, st->matP);
    st->matP_valid = 1;
}

__device__ void lm_solve(S2mState *st, const double *sum, int iter, const S2mParams &prm, bool do_trig)
{
    const int n_corr = (int)sum[27];
    st->n_corr = n_corr;
    st->iters = iter + 1;
    if (n_corr < prm.min_corr) return;                       // MO:1238: pose untouched, not converged

    float AtA[36], AtB[6], A[36], B[6], X[6];
    int k = 0;
    for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++, k++) { float v = (float)sum[k]; AtA[i * 6 + j] = v; AtA[j * 6 + i] = v; }
    for (int i = 0; i < 6; i++) AtB[i] = (float)sum[21 + i];
    for (int i = 0; i < 36; i++) { A[i] = AtA[i]; st->AtA[i] = AtA[i]; }
    for (int i = 0; i < 6; i++) { B[i] = AtB[i]; st->AtB[i] = AtB[i]; }
    cv_solve_qr<6, 6>(A, B, X);

    if (iter == 0) {
        // isDegenerate <=> the float Jacobi reports an eigenvalue < thresh.  Jacobi's absolute eigenvalue
        // error is O(n eps ||A||) <= ~1e-5 trace(A); if A - (thresh + 1e-4 trace) I is positive definite
        // every Jacobi eigenvalue is >= thresh for sure: not degenerate, and matP is never read
        // (MO:1301), so the eigen-decomposition is deferred until somebody asks for matP.
        double tr = 0.0;
        for (int i = 0; i < 6; i++) { tr += (double)AtA[i * 6 + i]; }
        for (int i = 0; i < 36; i++) st->AtA0[i] = AtA[i];
        if (all_eigenvalues_above(AtA, (double)prm.degeneracy_thresh + 1e-4 * tr + 1.0)) {
            st->is_degenerate = 0;
            st->matP_valid = 0;
        } else {
            degeneracy_full(st, AtA, prm.degeneracy_thresh);
        }
    }
    if (st->is_degenerate) {
        float X2[6];
        for (int i = 0; i < 6; i++) X2[i] = X[i];
        cv_gemm<6, 6, 1>(st->matP, X2, X);
    }
    for (int i = 0; i < 6; i++) { st->T[i] += X[i]; st->X[i] = X[i]; }
    if (do_trig) update_sincos(st);

    double r0 = (double)(X[0] * 57.29578f), r1 = (double)(X[1] * 57.29578f), r2 = (double)(X[2] * 57.29578f);
    double t0 = (double)(X[3] * 100), t1 = (double)(X[4] * 100), t2 = (double)(X[5] * 100);
    float deltaR = (float)sqrt(r0 * r0 + r1 * r1 + r2 * r2);
    float deltaT = (float)sqrt(t0 * t0 + t1 * t1 + t2 * t2);
    if ((double)deltaR < (double)prm.converge_deg && (double)deltaT < (double)prm.converge_cm) st->converged = 1;
}

struct PoseArg { float T[6]; };

__global__ void s2m_prepare_kernel(S2mState *st, PoseArg pose, const float *T_dev, const GridDesc *cd,
                                   const GridDesc *sd, S2mParams prm)
{
    for (int i = 0; i < 6; i++) st->T[i] = T_dev ? T_dev[i] : pose.T[i];
    update_sincos(st);
    st->converged = 0;
    st->iters = 0;
    st->n_corr = 0;
    st->ticket = 0;
    st->skipped = !(cd->n > prm.corner_map_min && sd->n > prm.surf_map_min);   // MO:1331
}

__global__ void s2m_state_init_kernel(S2mState *st)
{
    for (int i = 0; i < 6; i++) { st->T[i] = 0.f; st->cs[i] = (i & 1) ? 0.f : 1.f; st->AtB[i] = 0.f; st->X[i] = 0.f; }
    for (int i = 0; i < 36; i++) { st->matP[i] = 0.f; st->AtA[i] = 0.f; st->AtA0[i] = 0.f; }
    st->matP_valid = 1;                                      // the reference starts with matP = 0 (MO:361)
    st->converged = 0; st->iters = 0; st->n_corr = 0; st->is_degenerate = 0; st->skipped = 0; st->ticket = 0;
}

// index pairs of the 28 accumulated products of v = {arx, ary, arz, cx, cy, cz, b, 1}
__device__ __forceinline__ void pair_of(int k, int &ia, int &ib)
{
    ia = 7; ib = 7;                                          // k == 27: row count
    int c = 0;
    for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++, c++)
            if (c == k) { ia = i; ib = j; }
    if (k >= 21 && k < 27) { ia = k - 21; ib = 6; }
}

__global__ void __launch_bounds__(S2M_THREADS, S2M_CTAS_PER_SM)
s2m_loop_kernel(S2mParams prm, int it_begin, int it_end, S2mQueries q, MapIndexView cmap, MapIndexView smap,
                S2mState *st, double *partials, double *acc_out, S2mDebug dbg, int rank, int world, int do_solve,
                int prof_off)
{
    cg::grid_group grid = cg::this_grid();

    __shared__ float s_nn[15][S2M_TILE];                      // 5 neighbours x (x,y,z), SoA
    __shared__ float s_q[7][S2M_TILE];                        // po.xyz, sel.xyz, 5th squared distance (or -1)
    __shared__ float s_row[8][S2M_TILE];                      // Jacobian row, -residual, valid
    __shared__ double s_acc[S2M_NW][32];
    __shared__ double s_tot[32];
    __shared__ unsigned long long s_wkey[S2M_NW][KNN_CAP];    // per-warp in-gate candidate lists (phase A)
    __shared__ int s_wpos[S2M_NW][KNN_CAP];
    __shared__ int s_rb[9][S2M_TILE], s_re[9][S2M_TILE];      // cell-run ranges per query (phase A1 -> A2)

    if (__ldcg(&st->skipped)) return;                         // uniform over the grid (guard MO:1331)

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nc = q.nc_dev ? *q.nc_dev : q.nc_upper;
    const int ns = q.ns_dev ? *q.ns_dev : q.ns_upper;
    const int nq = nc + ns;
    // this rank's queries are qi = rank + world * j, j in [0, nloc): the first ncl are corner queries.
    // CTA b owns a contiguous range [j0, j1) (spatial locality of the cell reads) of equal WEIGHT: a corner
    // query counts 3x (dense 0.2 m cells -> long runs, Jacobi fit), see profiles/r01_s2m_phases.md
    const int nloc = nq > rank ? (nq - rank + world - 1) / world : 0;
    const int ncl = nc > rank ? (nc - rank + world - 1) / world : 0;
    const int G = (int)gridDim.x, bx = (int)blockIdx.x;
    constexpr long long CW = 3;
    const long long Wt = CW * ncl + (nloc - ncl);
    auto boundary = [&](int b) -> int {
        const long long p = Wt * b / G;
        return p < CW * ncl ? (int)(p / CW) : (int)(ncl + (p - CW * ncl));
    };
    const int j0 = boundary(bx), j1 = boundary(bx + 1);
    const int cnt = j1 - j0;

    int ia, ib;
    pair_of(lane, ia, ib);

    for (int iter = it_begin; iter < it_end; iter++) {
        const float crx = __ldcg(&st->cs[0]), srx = __ldcg(&st->cs[1]), cry = __ldcg(&st->cs[2]),
                    sry = __ldcg(&st->cs[3]), crz = __ldcg(&st->cs[4]), srz = __ldcg(&st->cs[5]);
        const float tX = __ldcg(&st->T[3]), tY = __ldcg(&st->T[4]), tZ = __ldcg(&st->T[5]);
        double acc = 0.0;
        long long pa = 0, pb = 0, pc = 0;
        const bool prof = (blockIdx.x == 0 && tid == 0);
        const long long t_start = clock64();
        if (prof) st->prof[iter % 10][0] = t_start;

        for (int t0 = 0; t0 < cnt; t0 += S2M_TILE) {
            const int tn = min(S2M_TILE, cnt - t0);
            // slots [0, ncs) of this tile are corner queries, [ncs, tn) surf queries
            const int ncs = max(0, min(tn, ncl - (j0 + t0)));
            // ---------------- phase A1: thread per query: transform + the 9 cell-run ranges (18 independent loads)
            if (tid < tn) {
                const int s = tid;
                const int qi = rank + world * (j0 + t0 + s);
                const bool is_corner = qi < nc;
                const float4 po = is_corner ? __ldg(&q.corner[qi]) : __ldg(&q.surf[qi - nc]);
                // pointAssociateToMap MO:513-527
                const float x1 = crz * po.x - srz * po.y;
                const float y1 = srz * po.x + crz * po.y;
                const float z1 = po.z;
                const float y2 = crx * y1 - srx * z1;
                const float z2 = srx * y1 + crx * z1;
                const float sx = cry * x1 + sry * z2 + tX;
                const float sy = y2 + tY;
                const float sz = -sry * x1 + cry * z2 + tZ;
                s_q[0][s] = po.x; s_q[1][s] = po.y; s_q[2][s] = po.z;
                s_q[3][s] = sx; s_q[4][s] = sy; s_q[5][s] = sz;
                int rb[9], re[9];
                knn_ranges(is_corner ? cmap : smap, sx, sy, sz, rb, re);
#pragma unroll
                for (int r = 0; r < 9; r++) { s_rb[r][s] = rb[r]; s_re[r][s] = re[r]; }
            }
            __syncthreads();
            // ---------------- phase A2: warp per query: candidates -> in-gate list -> 5 smallest
            for (int s = w; s < tn; s += S2M_NW) {
                const int qi = rank + world * (j0 + t0 + s);
                const bool is_corner = qi < nc;
                const float sx = s_q[3][s], sy = s_q[4][s], sz = s_q[5][s];
                int rbv[9], rev[9];
#pragma unroll
                for (int r = 0; r < 9; r++) { rbv[r] = s_rb[r][s]; rev[r] = s_re[r][s]; }

                int npos[5]; float nd[5]; int ni[5];
                const MapIndexView &mv = is_corner ? cmap : smap;
                const int found = knn5_warp(mv, sx, sy, sz, prm.knn_max_sqdist, lane, rbv, rev, s_wkey[w], s_wpos[w], npos, nd, ni);
                // lane j < 5 fetches neighbour j (an L1/L2 hit: the warp just read it) and stores it
                int myp = npos[0]; int myi = ni[0]; float myd = nd[0];
#pragma unroll
                for (int k = 1; k < 5; k++) { myp = (lane == k) ? npos[k] : myp; myi = (lane == k) ? ni[k] : myi; myd = (lane == k) ? nd[k] : myd; }
                if (lane < 5) {
                    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (myp >= 0) p = __ldg(&mv.sorted[myp]);
                    s_nn[lane * 3 + 0][s] = p.x; s_nn[lane * 3 + 1][s] = p.y; s_nn[lane * 3 + 2][s] = p.z;
                    if (dbg.knn_idx) { dbg.knn_idx[qi * 5 + lane] = myi; dbg.knn_d2[qi * 5 + lane] = myd; }
                }
                if (lane == 5) s_q[6][s] = (found == 5) ? nd[4] : -1.f;
            }
            __syncthreads();
            if (tid == 0) pa = clock64();
            // ---------------- phase B: thread per query
            // corner slots are served by warp 0, surf slots by warps 1..: the Jacobi and the QR code paths
            // never diverge inside a warp (TILE <= THREADS - 32 surf slots always fit)
            const int c32 = min(ncs, 32);                    // corner slots beyond 32 ride with the surf warps
            const int s = tid < 32 ? (tid < c32 ? tid : -1) : (c32 + tid - 32);
            if (s >= 0 && s < tn) {
                const int qi = rank + world * (j0 + t0 + s);
                const bool is_corner = qi < nc;
                const float d5 = s_q[6][s];
                bool ok = (d5 >= 0.f) && ((double)d5 < (double)prm.knn_max_sqdist);    // MO:1101 / MO:1183
                float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
                const float px = s_q[0][s], py = s_q[1][s], pz = s_q[2][s];
                if (ok) {
                    float nx[5], ny[5], nz[5];
#pragma unroll
                    for (int k = 0; k < 5; k++) { nx[k] = s_nn[3 * k][s]; ny[k] = s_nn[3 * k + 1][s]; nz[k] = s_nn[3 * k + 2][s]; }
                    const float sx = s_q[3][s], sy = s_q[4][s], sz = s_q[5][s];
                    ok = is_corner ? corner_fit(nx, ny, nz, sx, sy, sz, coeff) : surf_fit(nx, ny, nz, sx, sy, sz, coeff);
                }
                if (dbg.coeff) { dbg.coeff[qi] = coeff; dbg.valid[qi] = ok ? 1 : 0; }
                float v[8] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
                if (ok) {
                    // Jacobian row MO:1252-1271
                    v[0] = (crx * sry * srz * px + crx * crz * sry * py - srx * sry * pz) * coeff.x
                         + (-srx * srz * px - crz * srx * py - crx * pz) * coeff.y
                         + (crx * cry * srz * px + crx * cry * crz * py - cry * srx * pz) * coeff.z;
                    v[1] = ((cry * srx * srz - crz * sry) * px + (sry * srz + cry * crz * srx) * py + crx * cry * pz) * coeff.x
                         + ((-cry * crz - srx * sry * srz) * px + (cry * srz - crz * srx * sry) * py - crx * sry * pz) * coeff.z;
                    v[2] = ((crz * srx * sry - cry * srz) * px + (-cry * crz - srx * sry * srz) * py) * coeff.x
                         + (crx * crz * px - crx * srz * py) * coeff.y
                         + ((sry * srz + cry * crz * srx) * px + (crz * sry - cry * srx * srz) * py) * coeff.z;
                    v[3] = coeff.x; v[4] = coeff.y; v[5] = coeff.z;
                    v[6] = -coeff.w;
                    v[7] = 1.f;
                }
#pragma unroll
                for (int k = 0; k < 8; k++) s_row[k][s] = v[k];
            }
            __syncthreads();
            if (tid == 0) pb = clock64();
            // ---------------- phase C: lane k accumulates product k over rows w, w+16, ...
            if (lane < S2M_ACC)
                for (int r = w; r < tn; r += S2M_NW) acc += (double)s_row[ia][r] * (double)s_row[ib][r];
            __syncthreads();
            if (tid == 0) pc = clock64();
        }
        if (prof) { st->prof[iter % 10][1] = pa; st->prof[iter % 10][2] = pb; st->prof[iter % 10][3] = pc; }

        // ---- CTA partial (fixed order over warps)
        s_acc[w][lane] = acc;
        __syncthreads();
        if (tid < S2M_ACC) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < S2M_NW; k++) s += s_acc[k][tid];
            partials[(size_t)blockIdx.x * S2M_ACC + tid] = s;
        }
        grid.sync();
        if (prof) st->prof[iter % 10][4] = clock64();
        if (tid == 0) {                                      // per-CTA phase cycles of this iteration (diagnostics)
            double *cp = partials + (size_t)prof_off + (size_t)blockIdx.x * 4;
            cp[0] = (double)(pa - t_start); cp[1] = (double)(pb - pa); cp[2] = (double)(pc - pb);
            cp[3] = (double)(clock64() - pc);
        }

        // ---- CTA 0: deterministic grid reduction (16 interleaved slices, fixed order) + LM step
        if (blockIdx.x == 0) {
            double s = 0.0;
            if (lane < S2M_ACC)
                for (int b = w; b < (int)gridDim.x; b += S2M_NW) s += __ldcg(&partials[(size_t)b * S2M_ACC + lane]);
            s_acc[w][lane] = s;
            __syncthreads();
            if (tid < S2M_ACC) {
                double tsum = 0.0;
#pragma unroll
                for (int k = 0; k < S2M_NW; k++) tsum += s_acc[k][tid];
                s_tot[tid] = tsum;
                if (!do_solve) acc_out[tid] = tsum;
            }
            __syncthreads();
            if (prof) st->prof[iter % 10][5] = clock64();
            if (tid == 0 && do_solve) lm_solve(st, s_tot, iter, prm, false);
            __syncthreads();
            // the six sin/cos of the new pose, one per thread (fp64 libm calls are the long pole of the step)
            if (tid < 6 && do_solve) {
                const double a = (double)st->T[tid >> 1];
                st->cs[tid] = (tid & 1) ? (float)sin(a) : (float)cos(a);
            }
            if (prof) st->prof[iter % 10][6] = clock64();
        }
        if (!do_solve) break;
        grid.sync();
        if (prof) st->prof[iter % 10][7] = clock64();
        if (__ldcg(&st->converged)) break;                   // MO:1344-1345, uniform over the grid
    }
}

__global__ void s2m_solve_kernel(S2mParams prm, int iter, S2mState *st, const double *acc)
{
    if (st->converged || st->skipped) return;
    lm_solve(st, acc, iter, prm, true);
}

// matP on demand (see lm_solve): recompute the iteration-0 analysis from the saved AtA
__global__ void s2m_matp_kernel(S2mParams prm, S2mState *st)
{
    if (st->matP_valid) return;
    const int deg = st->is_degenerate;
    degeneracy_full(st, st->AtA0, prm.degeneracy_thresh);
    st->is_degenerate = deg;                                 // the decision was already made, keep it
}

}  // namespace

void S2mSolver::init(const S2mParams &p)
{
    prm_ = p;
    state_.ensure(1);
    int dev = 0, sms = 0, per_sm = 0;
    LLB_CUDA(cudaGetDevice(&dev));
    LLB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    LLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, s2m_loop_kernel, S2M_THREADS, 0));
    if (per_sm < 1) throw std::runtime_error("s2m_loop_kernel cannot be made resident");
    max_blocks_ = sms * per_sm;
    partials_.ensure((size_t)max_blocks_ * (S2M_ACC + 4));     // + 4 profile doubles per CTA
    acc_.ensure(32);
    s2m_state_init_kernel<<<1, 1>>>(state_.p);
    LLB_CUDA(cudaGetLastError());
    LLB_CUDA(cudaMemset(acc_.p, 0, 32 * sizeof(double)));
}

void S2mSolver::release()
{
    state_.release(); partials_.release(); acc_.release();
}

int S2mSolver::prepare(const float *T_host, const float *T_dev, const GridDesc *cd, const GridDesc *sd, cudaStream_t s)
{
    PoseArg pa{};
    if (T_host) for (int i = 0; i < 6; i++) pa.T[i] = T_host[i];
    s2m_prepare_kernel<<<1, 1, 0, s>>>(state_.p, pa, T_dev, cd, sd, prm_);
    LLB_CUDA(cudaGetLastError());
    return 1;
}

int S2mSolver::run(int it_begin, int it_end, const S2mQueries &q, const MapIndexView &cmap, const MapIndexView &smap,
                   const S2mDebug &dbg, int rank, int world, bool do_solve, cudaStream_t s)
{
    const int nq = std::max(1, div_up(q.nc_upper + q.ns_upper, world));
    int grid = std::max(1, std::min(max_blocks_, div_up(nq, S2M_QPB)));
    if (prm_.max_ctas > 0) grid = std::min(grid, prm_.max_ctas);
    S2mParams prm = prm_;
    S2mQueries qq = q; MapIndexView cm = cmap, sm = smap; S2mDebug dg = dbg;
    S2mState *st = state_.p; double *part = partials_.p, *acc = acc_.p;
    int ds = do_solve ? 1 : 0;
    int prof_off = max_blocks_ * S2M_ACC;
    last_grid_ = grid;
    void *args[] = { &prm, &it_begin, &it_end, &qq, &cm, &sm, &st, &part, &acc, &dg, &rank, &world, &ds, &prof_off };
    LLB_CUDA(cudaLaunchCooperativeKernel((const void *)s2m_loop_kernel, dim3(grid), dim3(S2M_THREADS), args, 0, s));
    return 1;
}

int S2mSolver::ensure_matp(cudaStream_t s)
{
    s2m_matp_kernel<<<1, 1, 0, s>>>(prm_, state_.p);
    LLB_CUDA(cudaGetLastError());
    return 1;
}

int S2mSolver::solve(int iter, cudaStream_t s)
{
    s2m_solve_kernel<<<1, 1, 0, s>>>(prm_, iter, state_.p, acc_.p);
    LLB_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace llb
