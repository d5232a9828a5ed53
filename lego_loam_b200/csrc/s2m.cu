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
#include "s2m_dev.cuh"
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
constexpr int S2M_QPB = S2M_NW;
constexpr int S2M_LIST_CAP = 1024;       // candidates listed per pass in the sharded-map mode          // target queries per CTA when sizing the grid (one per warp)


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
    const int n_corner = prm.global_corner >= 0 ? prm.global_corner : cd->n, n_surf = prm.global_surf >= 0 ? prm.global_surf : sd->n;
    st->skipped = !(n_corner > prm.corner_map_min && n_surf > prm.surf_map_min);   // MO:1331
}

__global__ void s2m_state_init_kernel(S2mState *st)
{
    for (int i = 0; i < 6; i++) { st->T[i] = 0.f; st->cs[i] = (i & 1) ? 0.f : 1.f; st->AtB[i] = 0.f; st->X[i] = 0.f; }
    for (int i = 0; i < 36; i++) { st->matP[i] = 0.f; st->AtA[i] = 0.f; st->AtA0[i] = 0.f; }
    st->matP_valid = 1;                                      // the reference starts with matP = 0 (MO:361)
    st->converged = 0; st->iters = 0; st->n_corr = 0; st->is_degenerate = 0; st->skipped = 0; st->ticket = 0;
    st->queue = 0; st->peer_timeout = 0; st->xchg = 0ull;
}


__global__ void __launch_bounds__(S2M_THREADS, S2M_CTAS_PER_SM)
s2m_loop_kernel(S2mParams prm, int it_begin, int it_end, S2mQueries q, MapIndexView cmap, MapIndexView smap,
                S2mState *st, double *partials, double *acc_out, S2mDebug dbg, int rank, int world, int do_solve,
                int prof_off, S2mPeers peers)
{
    cg::grid_group grid = cg::this_grid();

    __shared__ float s_nn[15][S2M_TILE];                      // 5 neighbours x (x,y,z), SoA
    __shared__ float s_q[7][S2M_TILE];                        // po.xyz, sel.xyz, 5th squared distance (or -1)
    __shared__ float s_row[8][S2M_TILE + 1];                  // Jacobian row, -residual, valid (+1: bank spread in phase C)
    __shared__ double s_acc[S2M_NW][32];
    __shared__ double s_tot[32];
    __shared__ unsigned long long s_wkey[S2M_NW][KNN_CAP];    // per-warp in-gate candidate lists (phase A)
    __shared__ int s_wpos[S2M_NW][KNN_CAP];
    __shared__ int s_rb[9][S2M_TILE], s_re[9][S2M_TILE];      // cell-run ranges per query (phase A1 -> A2)
    __shared__ int s_list[S2M_LIST_CAP];                      // sharded map: this CTA's queries inside the slab
    __shared__ int s_scan[33];

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
    // sharded map (own_axis >= 0): a rank handles the queries inside its slab, which are neighbours in the sweep; CTA b
    // therefore takes the INTERLEAVED queries b, b + G, b + 2G, ... so that every CTA sees the same share of owned ones
    const bool slab = prm.own_axis >= 0;
    const int j0 = slab ? 0 : boundary(bx), j1 = slab ? (nq > bx ? (nq - bx + G - 1) / G : 0) : boundary(bx + 1);
    const int cnt = j1 - j0;
    const int q_base = slab ? bx : rank, q_stride = slab ? G : world;
    auto query_of = [&](int j) -> int { return q_base + q_stride * j; };

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

        // sharded map: the CTA first lists the queries of its share that lie inside this rank's slab at the current pose
        // (order kept), S2M_LIST_CAP candidates at a time, and the tiles below run over that list only - a CTA of a
        // 1/N slab stages 1/N of the tiles instead of mostly-empty ones
        for (int c0 = 0; c0 < cnt; c0 += (slab ? S2M_LIST_CAP : cnt)) {
        int cnt_eff = cnt;
        if (slab) {
            const int c1 = min(cnt, c0 + S2M_LIST_CAP);
            int nl = 0;
            for (int base = c0; base < c1; base += S2M_THREADS) {
                const int j = base + tid;
                int own = 0, qi = 0;
                if (j < c1) {
                    qi = query_of(j);
                    const float4 po = qi < nc ? __ldg(&q.corner[qi]) : __ldg(&q.surf[qi - nc]);
                    float sx, sy, sz;
                    associate_to_map(crx, srx, cry, sry, crz, srz, tX, tY, tZ, po, sx, sy, sz);
                    const float oc = prm.own_axis == 0 ? sx : (prm.own_axis == 1 ? sy : sz);
                    own = (oc >= prm.own_lo && oc < prm.own_hi) ? 1 : 0;
                }
                int total;
                const int pos = nl + block_excl_scan(own, s_scan, total);
                if (own) s_list[pos] = qi;
                nl += total;
            }
            __syncthreads();
            cnt_eff = nl;
        }
        for (int t0 = 0; t0 < cnt_eff; t0 += S2M_TILE) {
            const int tn = min(S2M_TILE, cnt_eff - t0);
            // slots [0, ncs) of this tile are corner queries, [ncs, tn) surf queries
            const int ncs = slab ? 0 : max(0, min(tn, ncl - (j0 + t0)));
            // ---------------- phase A1: thread per query: transform + the 9 cell-run ranges (18 independent loads)
            if (tid < tn) {
                const int s = tid;
                const int qi = slab ? s_list[t0 + s] : query_of(j0 + t0 + s);
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
                // sharded map: a query outside this rank's slab is another rank's (its runs stay empty: no neighbours,
                // no row); every rank evaluates the same comparison on the same bits, so exactly one rank takes it
                const float oc = prm.own_axis == 0 ? sx : (prm.own_axis == 1 ? sy : sz);
                if (prm.own_axis < 0 || (oc >= prm.own_lo && oc < prm.own_hi)) {
                    knn_ranges(is_corner ? cmap : smap, sx, sy, sz, rb, re);
                } else {
#pragma unroll
                    for (int r = 0; r < 9; r++) { rb[r] = 0; re[r] = 0; }
                }
#pragma unroll
                for (int r = 0; r < 9; r++) { s_rb[r][s] = rb[r]; s_re[r][s] = re[r]; }
            }
            __syncthreads();
            // ---------------- phase A2: warp per query: candidates -> in-gate list -> 5 smallest
            // (a thread-per-query walk, the throughput form of batch.cu, was measured here too: with one or two warps
            // per SM nothing hides its dependent loads - 70k cycles per tile instead of 14k)
            for (int s = w; s < tn; s += S2M_NW) {
                const int qi = slab ? s_list[t0 + s] : query_of(j0 + t0 + s);
                const bool is_corner = qi < nc;
                const float sx = s_q[3][s], sy = s_q[4][s], sz = s_q[5][s];
                int rbv[9], rev[9];
                bool any = false;
#pragma unroll
                for (int r = 0; r < 9; r++) { rbv[r] = s_rb[r][s]; rev[r] = s_re[r][s]; any |= rev[r] > rbv[r]; }
                if (!any && !dbg.knn_idx) {                  // another rank's query, or empty space: no row (warp-uniform)
                    if (lane < 15) s_nn[lane][s] = 0.f;
                    if (lane == 15) s_q[6][s] = -1.f;
                    continue;
                }

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
                const int qi = slab ? s_list[t0 + s] : query_of(j0 + t0 + s);
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
                    jacobian_row(crx, srx, cry, sry, crz, srz, px, py, pz, coeff, v);
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
        }   // candidates c0 ..
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
            if (lane < S2M_ACC) {
                // the (up to 10) partials of this warp's slice are fetched together - one L2 round trip instead of ten
                // dependent ones - and added in the same fixed order as before
                constexpr int RMAX = 10;                     // ceil(148 CTAs / 16 warps)
                for (int b0 = w; b0 < (int)gridDim.x; b0 += S2M_NW * RMAX) {
                    double v[RMAX];
#pragma unroll
                    for (int k = 0; k < RMAX; k++) {
                        const int b = b0 + k * S2M_NW;
                        v[k] = b < (int)gridDim.x ? __ldcg(&partials[(size_t)b * S2M_ACC + lane]) : 0.0;
                    }
#pragma unroll
                    for (int k = 0; k < RMAX; k++) if (b0 + k * S2M_NW < (int)gridDim.x) s += v[k];
                }
            }
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
            if (peers.world > 1) {
                // ---- fused exchange of the 28 sums over NVLink (BASELINE config 4): one-shot all-to-all of P2P
                // stores into every peer's mailbox + flag, then every rank adds the contributions in RANK ORDER, so
                // all ranks hold bit-identical normal equations and take the identical LM step - no NCCL call, no host.
                // Mailboxes are double-buffered by the parity of the EXCHANGE number: a rank can be at most one exchange
                // ahead of a peer (it needs the peer's sums of exchange e to leave exchange e), and the number runs on
                // over the registrations of the context - a registration that ended on an even iteration is not followed
                // by one that starts in the same mailbox while a slow peer still reads it.
                const unsigned long long xn = st->xchg;
                const int par = (int)(xn & 1ull);
                const unsigned long long seq = xn + 1ull;
                if (tid < S2M_ACC)
                    for (int r = 0; r < peers.world; r++)
                        reinterpret_cast<volatile double *>(peers.box[r])[(par * S2M_MAX_PEERS + peers.rank) * 32 + tid] = s_tot[tid];
                __threadfence_system();
                __syncthreads();
                if (tid < peers.world) {
                    volatile unsigned long long *f = peers.flag[tid] + par * S2M_MAX_PEERS + peers.rank;
                    *f = seq;                                // after the fence: the data is visible before the flag
                    __threadfence_system();
                    volatile unsigned long long *mine = peers.flag[peers.rank] + par * S2M_MAX_PEERS + tid;
                    const long long t_wait = clock64();
                    while (*mine < seq) {                    // a peer that never arrives must not hang the GPU: ~2 s
                        if (clock64() - t_wait > 4000000000ll) { st->peer_timeout = 1; break; }
                    }
                    __threadfence_system();
                }
                __syncthreads();
                if (tid < S2M_ACC) {
                    const volatile double *box = reinterpret_cast<const volatile double *>(peers.box[peers.rank]);
                    double tsum = 0.0;
                    for (int r = 0; r < peers.world; r++) tsum += box[(par * S2M_MAX_PEERS + r) * 32 + tid];
                    s_tot[tid] = tsum;
                }
                __syncthreads();
                if (tid == 0) st->xchg = xn + 1ull;
            }
            if (prof) st->prof[iter % 10][5] = clock64();
            if (tid == 0 && do_solve) lm_solve(st, s_tot, iter, prm, false);
            __syncthreads();
            // the six sin/cos of the new pose, one per thread (glibc's sinf / cosf restated)
            if (tid < 6 && do_solve) {
                st->cs[tid] = pose_trig(st->T, tid);
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
                   const S2mDebug &dbg, int rank, int world, bool do_solve, cudaStream_t s, const S2mPeers *peers_in)
{
    S2mPeers peers{};
    if (peers_in) peers = *peers_in;
    const int nq = std::max(1, div_up(q.nc_upper + q.ns_upper, world));
    int grid = std::max(1, std::min(max_blocks_, div_up(nq, S2M_QPB)));
    if (prm_.max_ctas > 0) grid = std::min(grid, prm_.max_ctas);
    S2mParams prm = prm_;
    S2mQueries qq = q; MapIndexView cm = cmap, sm = smap; S2mDebug dg = dbg;
    S2mState *st = state_.p; double *part = partials_.p, *acc = acc_.p;
    int ds = do_solve ? 1 : 0;
    int prof_off = max_blocks_ * S2M_ACC;
    last_prof_off_ = prof_off;
    last_grid_ = grid;
    void *args[] = { &prm, &it_begin, &it_end, &qq, &cm, &sm, &st, &part, &acc, &dg, &rank, &world, &ds, &prof_off, &peers };
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
