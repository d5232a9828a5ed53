// odom.cuh — K5: the featureAssociation scan-to-scan matcher on the device.
//   TransformToStart FA:860-883, findCorrespondingSurfFeatures FA:1155-1268,
//   findCorrespondingCornerFeatures FA:1044-1153, calculateTransformationSurf FA:1270-1377,
//   calculateTransformationCorner FA:1379-1478, updateTransformation FA:1666-1695.
// Up to 25 + 25 strictly sequential tiny iterations over <= a few hundred queries: the
// whole of updateTransformation is ONE persistent CTA (no launch or host round trip
// between iterations).
#pragma once
#include "common.cuh"
#include "grid_index.cuh"
#include <vector>

namespace llb {

struct OdomParams {
    float nearest_sqdist;        // 25   UT:125
    int   max_iter;              // 25
    int   min_corr;              // 10
    float degeneracy_thresh;     // 10
    float converge_deg, converge_cm;   // 0.1, 0.1
};

struct OdomState {               // device-resident, persists across sweeps
    float T[6];                  // transformCur
    int   is_degenerate;         // FA:179 (shared by both solvers, C6)
    float matP[9];               // FA:180
    int   iters[2];              // iterations executed: [0] surf loop, [1] corner loop
    int   converged[2];
    int   n_corr;                // laserCloudOri size of the last iteration run
    int   skipped;               // guard FA:1668
    int   more;                  // return value of the last calculateTransformation* (C8)
};

struct OdomBatchJob {            // one sweep pair of a batched launch (device-resident table entry)
    const float4 *sharp, *flat, *cornerLast, *surfLast;
    int nsharp, nflat, ncl, nsl;
    float *ind;                  // 5 x cap floats: cInd1, cInd2, sInd1, sInd2, sInd3 (kept between sweeps, C3 / C4)
    int cap;
    OdomState *st;               // T in / out, isDegenerate / matP persist per slot
};
void launch_odom_batch(const OdomParams &prm, const OdomBatchJob *jobs_dev, int count, cudaStream_t s);
void launch_odom_batch_set_pose(const OdomBatchJob *jobs_dev, const float *poses_dev, int count, cudaStream_t s);
void launch_odom_state_init(OdomState *st, int count, cudaStream_t s);
void launch_odom_fill(float *p, int n, float v, cudaStream_t s);

class OdomSolver {
public:
    void init(const OdomParams &p);
    void release();
    DevBuf<float4> &cornerLast() { return cornerLast_; }
    DevBuf<float4> &surfLast() { return surfLast_; }
    DevBuf<float4> &sharp() { return sharp_; }
    DevBuf<float4> &flat() { return flat_; }
    int set_last(int ncl, int nsl, cudaStream_t s);
    void set_features(int nsharp, int nflat);
    bool ready() const { return last_set_ && feat_set_; }
    OdomState *state_dev() { return state_.p; }
    int optimize(const float *T, cudaStream_t s);
    int iterate(int which, const float *T, int iter, cudaStream_t s);
    void download_correspondences(std::vector<float4> &ori, std::vector<float4> &coeff, cudaStream_t s);
    void download_search_ind(int which, std::vector<float> &i1, std::vector<float> &i2, std::vector<float> &i3,
                             cudaStream_t s);

private:
    void ensure_work();
    OdomParams prm_{};
    DevBuf<OdomState> state_;
    DevBuf<float4> cornerLast_, surfLast_, sharp_, flat_;
    DevBuf<float> ind_;           // cInd1, cInd2, sInd1, sInd2, sInd3 (float, C3)
    DevBuf<float4> dbg_coeff_; DevBuf<int> dbg_valid_;
    int ncl_ = 0, nsl_ = 0, nsharp_ = 0, nflat_ = 0, cap_ = 0;
    int dbg_which_ = -1;
    bool last_set_ = false, feat_set_ = false;
    GridIndex gridCorner_, gridSurf_;   // uniform grids over the previous sweep's clouds (cell = gate radius)
    bool grids_init_ = false, grids_built_ = false;
};

}  // namespace llb
