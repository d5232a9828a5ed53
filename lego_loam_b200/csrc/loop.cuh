// loop.cuh — SURVEY 8(f)-4: the data-parallel part of the loop closure of mapOptimization on the device.
//   detectLoopClosure MO:838-861 (cloud part): latestSurfKeyFrameCloud / nearHistorySurfKeyFrameCloud(DS) come from
//   the device key-frame store (keyframes.cuh) + K1; performLoopClosure MO:892-904: pcl::IterativeClosestPoint
//   <PointXYZI, PointXYZI>::align + getFitnessScore as configured there, ONE persistent cooperative kernel for all
//   iterations: exact 1-NN of every source point (the 100 m correspondence gate is no bound in practice, so the
//   search is a tiled exhaustive scan, target tiles staged in shared memory, (distance, index) minima combined with
//   64-bit atomicMin: smaller index wins ties), the 17 fp64 sums of pcl::umeyama, a 3x3 one-sided Jacobi SVD, the
//   float 4x4 products and in-place cloud transform of PCL's icp.hpp and DefaultConvergenceCriteria on one thread.
// PCL / Eigen are absent offline: the algorithm is PCL 1.8's as published (oracle/llo_loop.c states the same, "parity
// unpinned"); means and covariance are accumulated in fp64 where Eigen sums floats in its packet order.
#pragma once
#include "common.cuh"

namespace llb {

struct IcpParams {
    int max_iterations;            // 100   MO:894
    double max_corr_dist;          // 100   MO:893
    double transformation_epsilon; // 1e-6  MO:895
    double fitness_epsilon;        // 1e-6  MO:896 (relative MSE)
};

struct IcpState {                  // device-resident result
    float T[16];                   // final_transformation_ (row-major)
    float Tr[16];                  // transformation_ of the last iteration
    double prev_mse, fitness;
    double sums[20];               // n, sum p[3], sum q[3], sum q p^T[9], sum d^2 of the last iteration (diagnostics)
    int converged, iterations, state, n_corr, done;
};

class IcpSolver {
public:
    void init();
    void release();
    // src (n_src points, untouched) -> aligned against tgt; everything is enqueued on s, the result stays in state_dev()
    int run(const IcpParams &p, const float4 *src, int n_src, const float4 *tgt, int n_tgt, int max_iter_override, cudaStream_t s);
    IcpState *state_dev() { return state_.p; }
    const unsigned long long *nn_dev() const { return nn_.p; }   // (distance bits << 32 | target index) of the last search
private:
    DevBuf<IcpState> state_;
    DevBuf<float4> cur_;
    DevBuf<unsigned long long> nn_;
    DevBuf<double> partials_;
    int max_blocks_ = 0;
};

// stable compaction of the points with (int)intensity >= 0 (MO:845-849) by ONE CTA; *n_out_dev receives the count
int launch_loop_filter_intensity(const float4 *in, int n, float4 *out, int *n_out_dev, cudaStream_t s);

}  // namespace llb
