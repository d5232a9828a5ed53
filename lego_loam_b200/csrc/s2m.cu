// s2m.cu — K3 (fused transform -> radius-bounded exact 5-NN -> line / plane fit ->
// residual + Jacobian row) and K4 (J^T J / J^T r reduction, LM step, degeneracy
// projection, convergence test) in ONE kernel per LM iteration.
//
// Work decomposition: one warp per query.  The 9 contiguous cell runs around the query
// are read cooperatively (lane = candidate, coalesced float4 loads, all 9 loads in flight
// before the first use), each lane keeps a private sorted top-5 in registers and five
// rounds of warp-wide lexicographic (distance, original index) arg-min produce the exact
// 5-NN with the oracle's tie rule.  The fit is evaluated redundantly by every lane (no
// divergence, no shuffles) so that lane k can accumulate the k-th of the 27 products of
// the normal equations in fp64 without any data exchange.  Blocks write deterministic
// partial sums; the last block to finish (ticket) adds them in a fixed order, rounds to
// fp32 exactly like cv::gemm does, and one thread performs the 6x6 LM step so the new
// pose, its sin/cos and the convergence flag are on the device before the next launch.
//
// Numerics: IEEE fp32 without contraction (-fmad=false) in the reference's association
// order; the double-promoted sub-expressions of SURVEY.md Appendix B are evaluated in
// fp64; pose sin/cos are (float)sin((double)x) (correctly rounded float).
#include "s2m.cuh"
#include "linalg.cuh"

namespace llb {

namespace {

constexpr int S2M_THREADS = 256;
constexpr int S2M_NW = S2M_THREADS / 32;
constexpr int S2M_MAX_BLOCKS = 148 * 4;

__device__ __forceinline__ bool lex_less(float d1, int i1, float d2, int i2)
{
    return d1 < d2 || (d1 == d2 && i1 < i2);
}

struct Top5 {
    float D[5]; int I[5]; int S[5];
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int k = 0; k < 5; k++) { D[k] = __int_as_float(0x7f800000); I[k] = INT_MAX; S[k] = -1; }
    }
    __device__ __forceinline__ void insert(float d, int oi, int pos)
    {
        if (lex_less(d, oi, D[4], I[4])) {
            D[4] = d; I[4] = oi; S[4] = pos;
#pragma unroll
            for (int k = 4; k > 0; k--) {
                if (lex_less(D[k], I[k], D[k - 1], I[k - 1])) {
                    float td = D[k]; D[k] = D[k - 1]; D[k - 1] = td;
                    int ti = I[k]; I[k] = I[k - 1]; I[k - 1] = ti;
                    int ts = S[k]; S[k] = S[k - 1]; S[k - 1] = ts;
                }
            }
        }
    }
    __device__ __forceinline__ void pop()
    {
#pragma unroll
        for (int k = 0; k < 4; k++) { D[k] = D[k + 1]; I[k] = I[k + 1]; S[k] = S[k + 1]; }
        D[4] = __int_as_float(0x7f800000); I[4] = INT_MAX; S[4] = -1;
    }
};

// flann::L2_Simple<float>: sequential float sum of squared differences
__device__ __forceinline__ float l2_simple(float qx, float qy, float qz, const float4 &p)
{
    float diff = qx - p.x;
    float d = diff * diff;
    diff = qy - p.y; d += diff * diff;
    diff = qz - p.z; d += diff * diff;
    return d;
}

// Exact 5-NN of (qx,qy,qz) among the map points inside the 3x3x3 cell neighbourhood.
// Every lane returns the same result.  nn[k].w carries the original map index bits.
__device__ __forceinline__ void knn5_warp(const MapIndexView &m, float qx, float qy, float qz, int lane,
                                          float4 (&nn)[5], float (&nd)[5], int (&ni)[5])
{
    const GridDesc *g = m.desc;
    const int dimx = g->dim[0], dimy = g->dim[1], dimz = g->dim[2];
    const float inv = g->inv_cell;
    const int cx = grid_coord(qx, g->org[0], inv), cy = grid_coord(qy, g->org[1], inv), cz = grid_coord(qz, g->org[2], inv);

    int rb = 0, re = 0;
    if (lane < 9) {
        const int y = cy + (lane % 3) - 1, z = cz + (lane / 3) - 1;
        if (y >= 0 && y < dimy && z >= 0 && z < dimz) {
            const int x0 = max(cx - 1, 0), x1 = min(cx + 1, dimx - 1);
            if (x0 <= x1) {
                const int row = (z * dimy + y) * dimx;
                rb = __ldg(&m.cell_begin[row + x0]);
                re = __ldg(&m.cell_begin[row + x1 + 1]);
            }
        }
    }
    Top5 t;
    t.init();
    float4 c[9];
    int pos[9];
    bool longrun = false;
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const int b = __shfl_sync(FULL, rb, r), e = __shfl_sync(FULL, re, r);
        pos[r] = (b + lane < e) ? b + lane : -1;
        longrun |= (e - b) > 32;
        if (pos[r] >= 0) c[r] = __ldg(&m.sorted[pos[r]]);
    }
#pragma unroll
    for (int r = 0; r < 9; r++)
        if (pos[r] >= 0) t.insert(l2_simple(qx, qy, qz, c[r]), __float_as_int(c[r].w), pos[r]);
    if (longrun) {                                           // warp-uniform
        for (int r = 0; r < 9; r++) {
            const int b = __shfl_sync(FULL, rb, r), e = __shfl_sync(FULL, re, r);
            for (int i = b + 32 + lane; i < e; i += 32) {
                float4 p = __ldg(&m.sorted[i]);
                t.insert(l2_simple(qx, qy, qz, p), __float_as_int(p.w), i);
            }
        }
    }
    // five rounds of warp-wide lexicographic arg-min over the list heads
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const unsigned db = __float_as_uint(t.D[0]);          // d >= 0: uint order == float order
        const unsigned mind = __reduce_min_sync(FULL, db);
        const unsigned ci = (db == mind) ? (unsigned)t.I[0] : 0xffffffffu;
        const unsigned mini = __reduce_min_sync(FULL, ci);
        const bool win = (db == mind) && ((unsigned)t.I[0] == mini);
        const unsigned ball = __ballot_sync(FULL, win);
        const int src = __ffs(ball) - 1;
        const int p = __shfl_sync(FULL, t.S[0], src);
        nd[r] = __uint_as_float(mind);
        ni[r] = (p >= 0) ? (int)mini : -1;
        if (p >= 0) nn[r] = __ldg(&m.sorted[p]);
        else nn[r] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        if (lane == src) t.pop();
    }
}

// cornerOptimization body for one query (MO:1102-1170).  Returns true when the row is accepted.
__device__ __forceinline__ bool corner_fit(const float4 (&nn)[5], float x0, float y0, float z0, float4 &coeff)
{
    float cx = 0, cy = 0, cz = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) { cx += nn[j].x; cy += nn[j].y; cz += nn[j].z; }
    cx /= 5; cy /= 5; cz /= 5;

    float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        float ax = nn[j].x - cx, ay = nn[j].y - cy, az = nn[j].z - cz;
        a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
        a22 += ay * ay; a23 += ay * az;
        a33 += az * az;
    }
    a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;

    float A1[9] = { a11, a12, a13, a12, a22, a23, a13, a23, a33 }, D1[3], V1[9];
    cv_eigen<3>(A1, D1, V1);
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
__device__ __forceinline__ bool surf_fit(const float4 (&nn)[5], float x0, float y0, float z0, float4 &coeff)
{
    float A0[15], B0[5] = { -1.f, -1.f, -1.f, -1.f, -1.f }, X0[3];
#pragma unroll
    for (int j = 0; j < 5; j++) { A0[3 * j] = nn[j].x; A0[3 * j + 1] = nn[j].y; A0[3 * j + 2] = nn[j].z; }
    cv_solve_qr<5, 3>(A0, B0, X0);

    float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1;
    float ps = sqrtf(pa * pa + pb * pb + pc * pc);
    pa /= ps; pb /= ps; pc /= ps; pd /= ps;

    bool planeValid = true;
#pragma unroll
    for (int j = 0; j < 5; j++)
        if ((double)fabsf(pa * nn[j].x + pb * nn[j].y + pc * nn[j].z + pd) > 0.2) planeValid = false;
    if (!planeValid) return false;

    float pd2 = pa * x0 + pb * y0 + pc * z0 + pd;
    float s = (float)(1.0 - 0.9 * (double)fabsf(pd2) / (double)sqrtf(sqrtf(x0 * x0 + y0 * y0 + z0 * z0)));
    coeff = make_float4(s * pa, s * pb, s * pc, s * pd2);
    return (double)s > 0.1;
}

__device__ __forceinline__ float sel7(const float (&v)[7], int i)
{
    float r = v[0];
#pragma unroll
    for (int k = 1; k < 7; k++) r = (i == k) ? v[k] : r;
    return r;
}

__device__ void update_sincos(S2mState *st)
{
    st->cs[0] = (float)cos((double)st->T[0]); st->cs[1] = (float)sin((double)st->T[0]);
    st->cs[2] = (float)cos((double)st->T[1]); st->cs[3] = (float)sin((double)st->T[1]);
    st->cs[4] = (float)cos((double)st->T[2]); st->cs[5] = (float)sin((double)st->T[2]);
}

// LMOptimization tail MO:1273-1326 on the 28 reduced sums (one thread)
__device__ void lm_solve(S2mState *st, const double *sum, int iter, const S2mParams &prm)
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
        float E[6], V[36], V2[36], Vinv[36];
        for (int i = 0; i < 36; i++) A[i] = AtA[i];
        cv_eigen<6>(A, E, V);
        for (int i = 0; i < 36; i++) V2[i] = V[i];
        int deg = 0;
        for (int i = 5; i >= 0; i--) {
            if (E[i] < prm.degeneracy_thresh) {
                for (int j = 0; j < 6; j++) V2[i * 6 + j] = 0.f;
                deg = 1;
            } else break;
        }
        st->is_degenerate = deg;
        cv_inv_lu<6>(V, Vinv);
        cv_gemm<6, 6, 6>(Vinv, V2, st->matP);
    }
    if (st->is_degenerate) {
        float X2[6];
        for (int i = 0; i < 6; i++) X2[i] = X[i];
        cv_gemm<6, 6, 1>(st->matP, X2, X);
    }
    for (int i = 0; i < 6; i++) { st->T[i] += X[i]; st->X[i] = X[i]; }
    update_sincos(st);

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
    for (int i = 0; i < 36; i++) { st->matP[i] = 0.f; st->AtA[i] = 0.f; }
    st->converged = 0; st->iters = 0; st->n_corr = 0; st->is_degenerate = 0; st->skipped = 0; st->ticket = 0;
}

__global__ void __launch_bounds__(S2M_THREADS)
s2m_iter_kernel(S2mParams prm, int iter, S2mQueries q, MapIndexView cmap, MapIndexView smap,
                S2mState *__restrict__ st, double *__restrict__ partials, double *__restrict__ acc_out,
                S2mDebug dbg, int rank, int world, int do_solve)
{
    if (st->converged || st->skipped) return;                 // uniform over the grid

    __shared__ double s_acc[S2M_NW][32];
    __shared__ double s_tot[32];
    __shared__ int s_last;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nc = q.nc_dev ? *q.nc_dev : q.nc_upper;
    const int ns = q.ns_dev ? *q.ns_dev : q.ns_upper;
    const int nq = nc + ns;

    const float crx = st->cs[0], srx = st->cs[1], cry = st->cs[2], sry = st->cs[3], crz = st->cs[4], srz = st->cs[5];
    const float tX = st->T[3], tY = st->T[4], tZ = st->T[5];

    // lane k accumulates product (ia, ib) of v = {arx, ary, arz, cx, cy, cz, b}
    int ia = 0, ib = 0;
    {
        int k = 0;
        for (int i = 0; i < 6; i++)
            for (int j = i; j < 6; j++, k++)
                if (k == lane) { ia = i; ib = j; }
        if (lane >= 21 && lane < 27) { ia = lane - 21; ib = 6; }
    }
    double acc = 0.0;

    for (int qi = rank + world * (blockIdx.x * S2M_NW + w); qi < nq; qi += world * gridDim.x * S2M_NW) {
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

        float4 nn[5]; float nd[5]; int ni[5];
        knn5_warp(is_corner ? cmap : smap, sx, sy, sz, lane, nn, nd, ni);

        if (dbg.knn_idx && lane < 5) {
            // every lane holds all five; lane j writes entry j
            int vi = ni[0]; float vd = nd[0];
#pragma unroll
            for (int k = 1; k < 5; k++) { vi = (lane == k) ? ni[k] : vi; vd = (lane == k) ? nd[k] : vd; }
            dbg.knn_idx[qi * 5 + lane] = vi;
            dbg.knn_d2[qi * 5 + lane] = vd;
        }

        bool ok = (ni[4] >= 0) && ((double)nd[4] < (double)prm.knn_max_sqdist);   // MO:1101 / MO:1183
        float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) ok = is_corner ? corner_fit(nn, sx, sy, sz, coeff) : surf_fit(nn, sx, sy, sz, coeff);

        if (dbg.coeff && lane == 0) { dbg.coeff[qi] = coeff; dbg.valid[qi] = ok ? 1 : 0; }

        if (ok) {
            // Jacobian row MO:1252-1271
            float v[7];
            v[0] = (crx * sry * srz * po.x + crx * crz * sry * po.y - srx * sry * po.z) * coeff.x
                 + (-srx * srz * po.x - crz * srx * po.y - crx * po.z) * coeff.y
                 + (crx * cry * srz * po.x + crx * cry * crz * po.y - cry * srx * po.z) * coeff.z;
            v[1] = ((cry * srx * srz - crz * sry) * po.x + (sry * srz + cry * crz * srx) * po.y + crx * cry * po.z) * coeff.x
                 + ((-cry * crz - srx * sry * srz) * po.x + (cry * srz - crz * srx * sry) * po.y - crx * sry * po.z) * coeff.z;
            v[2] = ((crz * srx * sry - cry * srz) * po.x + (-cry * crz - srx * sry * srz) * po.y) * coeff.x
                 + (crx * crz * po.x - crx * srz * po.y) * coeff.y
                 + ((sry * srz + cry * crz * srx) * po.x + (crz * sry - cry * srx * srz) * po.y) * coeff.z;
            v[3] = coeff.x; v[4] = coeff.y; v[5] = coeff.z;
            v[6] = -coeff.w;
            if (lane < 27) acc += (double)sel7(v, ia) * (double)sel7(v, ib);
            else if (lane == 27) acc += 1.0;
        }
    }

    // ---- block partial (fixed order over warps)
    s_acc[w][lane] = acc;
    __syncthreads();
    if (tid < S2M_ACC) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < S2M_NW; k++) s += s_acc[k][tid];
        partials[(size_t)blockIdx.x * S2M_ACC + tid] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned t = atomicAdd(&st->ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    // ---- last block: deterministic grid reduction (8 interleaved slices, fixed order)
    {
        double s = 0.0;
        if (lane < S2M_ACC)
            for (int b = w; b < (int)gridDim.x; b += S2M_NW) s += __ldcg(&partials[(size_t)b * S2M_ACC + lane]);
        __syncthreads();
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
    }
    if (tid == 0) {
        st->ticket = 0;
        if (do_solve) lm_solve(st, s_tot, iter, prm);
    }
}

__global__ void s2m_solve_kernel(S2mParams prm, int iter, S2mState *st, const double *acc)
{
    if (st->converged || st->skipped) return;
    lm_solve(st, acc, iter, prm);
}

}  // namespace

void S2mSolver::init(const S2mParams &p)
{
    prm_ = p;
    state_.ensure(1);
    max_blocks_ = S2M_MAX_BLOCKS;
    partials_.ensure((size_t)max_blocks_ * S2M_ACC);
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

int S2mSolver::iterate(int iter, const S2mQueries &q, const MapIndexView &cmap, const MapIndexView &smap,
                       const S2mDebug &dbg, int rank, int world, bool do_solve, cudaStream_t s)
{
    const int nq = std::max(1, div_up(q.nc_upper + q.ns_upper, world));
    const int grid = std::min(max_blocks_, div_up(nq, S2M_NW));
    s2m_iter_kernel<<<grid, S2M_THREADS, 0, s>>>(prm_, iter, q, cmap, smap, state_.p, partials_.p, acc_.p, dbg,
                                                 rank, world, do_solve ? 1 : 0);
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
