// s2m.cuh — K3 + K4: the scan-to-map iteration of mapOptimization on the device.
//   cornerOptimization MO:1093-1174, surfOptimization MO:1176-1227, LMOptimization
//   MO:1229-1327, loop scan2MapOptimization MO:1329-1350.
#pragma once
#include "common.cuh"
#include "grid_index.cuh"

namespace llb {

constexpr int S2M_ACC = 28;          // 21 upper-tri AtA + 6 AtB + row count

struct S2mParams {
    float knn_max_sqdist;            // 1.0
    int   min_corr;                  // 50
    float degeneracy_thresh;         // 100
    float converge_deg, converge_cm; // 0.05, 0.05
    int   corner_map_min, surf_map_min; // 10, 100
    int   max_ctas;                  // 0 = one per SM
    // sharded map (shard.cuh): this rank handles the queries whose MAPPED coordinate on own_axis lies in
    // [own_lo, own_hi); own_axis < 0: all queries.  global_corner / global_surf: sizes of the unsharded DS maps for the
    // guard MO:1331 (< 0: the sizes of the maps this context holds)
    int   own_axis = -1;
    float own_lo = -FLT_MAX, own_hi = FLT_MAX;
    int   global_corner = -1, global_surf = -1;
};

struct S2mState {                    // device-resident, persists across registrations
    float T[6];                      // transformTobeMapped
    float cs[6];                     // cRoll sRoll cPitch sPitch cYaw sYaw (MO:498-506)
    int   converged;
    int   iters;
    int   n_corr;
    int   is_degenerate;             // MO:202 (persists, C6)
    int   skipped;                   // guard MO:1331 failed
    float matP[36];                  // MO:203 (valid when matP_valid; else derived on demand from AtA0)
    float AtA0[36];                  // AtA of the last executed iteration 0
    int   matP_valid;
    float AtA[36], AtB[6], X[6];     // last LM step (diagnostics)
    unsigned ticket;                 // last-block election
    int peer_timeout;                // fused multi-GPU exchange: a peer's sums did not arrive within ~2 s
    unsigned long long xchg;         // exchanges performed so far: the same on every rank (identical LM steps), so it
                                     // numbers the exchanges and alternates the mailboxes ACROSS registrations too
    unsigned queue;                  // chunk queue of the persistent kernel (zero between iterations)
    long long prof[10][8];           // clock64 stamps of CTA 0 per iteration: start, A, B, C, sync1, reduce, solve, sync2
};

struct S2mDebug {                    // optional per-query outputs (nullptr = off)
    float4 *coeff;                   // [nq] coefficient row (valid or not)
    int *valid;                      // [nq] 1 when the row was accepted (s > 0.1)
    int *knn_idx;                    // [nq*5] original map indices, -1 if fewer than 5 in range
    float *knn_d2;                   // [nq*5]
};

constexpr int S2M_MAX_PEERS = 8;     // one NVSwitch domain
// peer mailboxes of the fused multi-GPU exchange (sharded registration, BASELINE config 4): box[r] / flag[r] are rank
// r's mailbox mapped into THIS process (cudaIpc); layout double[2][S2M_MAX_PEERS][32] / unsigned long long[2][S2M_MAX_PEERS]
struct S2mPeers {
    double *box[S2M_MAX_PEERS];
    unsigned long long *flag[S2M_MAX_PEERS];
    int world, rank;
};

struct S2mQueries {
    const float4 *corner; const int *nc_dev; int nc_upper;
    const float4 *surf;   const int *ns_dev; int ns_upper;
};

class S2mSolver {
public:
    void init(const S2mParams &p);
    void release();
    S2mState *state_dev() { return state_.p; }
    void set_shard(int axis, float lo, float hi, int global_corner, int global_surf)
    { prm_.own_axis = axis; prm_.own_lo = lo; prm_.own_hi = hi; prm_.global_corner = global_corner; prm_.global_surf = global_surf; }
    double *acc_dev() { return acc_.p; }
    // upload T, compute its sin/cos on the device, evaluate the map-size guard, reset flags
    int prepare(const float *T_host, const float *T_dev, const GridDesc *corner_desc, const GridDesc *surf_desc,
                cudaStream_t s);
    // iterations [it_begin, it_end) in ONE cooperative launch, stopping early on convergence.
    // rank/world shard the queries (world = 1: everything); do_solve = false runs exactly one
    // iteration and stops after the 28 sums are in acc_dev() (for an external all-reduce)
    int run(int it_begin, int it_end, const S2mQueries &q, const MapIndexView &cmap, const MapIndexView &smap,
            const S2mDebug &dbg, int rank, int world, bool do_solve, cudaStream_t s, const S2mPeers *peers = nullptr);
    int solve(int iter, cudaStream_t s);
    // makes S2mState::matP valid (it is computed lazily when the registration was not degenerate)
    int ensure_matp(cudaStream_t s);

private:
    S2mParams prm_{};
    DevBuf<S2mState> state_;
    DevBuf<double> partials_;
    DevBuf<double> acc_;
    int max_blocks_ = 0;
    int last_grid_ = 0;
    int last_prof_off_ = 0;
public:
    // per-CTA phase cycles {A, B, C, wait at the grid barrier} of the last iteration of the last run
    const double *cta_profile_dev() const { return partials_.p + (size_t)last_prof_off_; }
    int last_grid() const { return last_grid_; }
};

}  // namespace llb
