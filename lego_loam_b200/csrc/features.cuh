// features.cuh — feature extraction of one segmented sweep on the device: adjustDistortion (no-IMU
// branch, FA:491-619), calculateSmoothness (FA:621-641), markOccludedPoints (FA:643-678), extractFeatures (FA:680-784).
//
// Launches per sweep: 1 H2D (one pinned block: cloud, range, column, ground flag, ring bounds) ->
//   fe_point_kernel  per point: azimuth, first "half passed" point (atomicMin), curvature, state reset, sort records
//   fe_mark_kernel   per point: relative time -> intensity, axis swap; occlusion / parallel-beam marks
//   fe_ring_kernel   one CTA per ring: its 6 sectors sorted by 6 warps (libstdc++ std::sort move for move, std_sort.cuh),
//                    then the greedy picks of the sectors in order by one warp (32 candidates evaluated per step),
//                    then the ordered gather of the ring's less-flat points
//   voxel_cta_kernel one CTA per ring: VoxelGrid(0.2) of surfPointsLessFlatScan (FA:778-782), K1 as it is
//   fe_concat_kernel ring-major concatenation into the four output clouds
// -> 1 D2H (counts + four clouds).  Every kernel reads its sweep from a device table (blockIdx.y = sweep): a single
// context has one entry, FeatureBatch one per slot - the same five launches serve 64 sequences.  Per-point work is
// HBM-trivial (30k points); the step is bound by the dependent chains of a ring's CTA: ~22 warp-wide partitions per
// sector sort (std_sort.cuh, closed-form Hoare partition) and the <= 24 picks per sector, each of which can block the
// candidates next to it.  transform_to_end(): TransformToEnd (FA:885-953) of the less-sharp / less-flat clouds.
#pragma once
#include "common.cuh"
#include "voxel_dev.cuh"
#include <vector>

namespace llb {

constexpr int FE_MAX_RINGS = 128;
constexpr int FE_SHARP_PER_RING = 12, FE_LSHARP_PER_RING = 120, FE_FLAT_PER_RING = 24;   // 6 sectors x (2, 20, 4)

struct FeParams { float edge_threshold, surf_threshold, scan_period, leaf; };   // UT:116-117, UT:107, FA:214

struct FeHeader { int first_half; int counts[4]; int release_seq; int pad[2]; int prof[7]; int stale_ind; };

constexpr int FE_IMU_QUEUE = 200;                // imuQueLength UT:109
constexpr size_t FE_HEAD = 512;                   // bytes reserved for the FeView at the head of a staged sweep

struct FeImu {                   // the IMU ring buffers of FeatureAssociation (FA:82-135) for one sweep, device copy
    double time[FE_IMU_QUEUE];
    float roll[FE_IMU_QUEUE], pitch[FE_IMU_QUEUE], yaw[FE_IMU_QUEUE];
    float velo[3][FE_IMU_QUEUE], shift[3][FE_IMU_QUEUE], angular[3][FE_IMU_QUEUE];
    double time_scan_cur;
    int pointer_last, pointer_last_iteration;
};

struct FeImuOut {                // what the IMU branch of adjustDistortion leaves behind (FA:556-611)
    float start[9];              // imuRoll/Pitch/YawStart, imuVeloX/Y/ZStart, imuShiftX/Y/ZStart
    float angular_cur[3];        // imuAngularRotationX/Y/ZCur (at the first point)
    float cur[3];                // imuRoll/Pitch/YawCur of the last point
    float velo_from_start_cur[3];// imuVeloFromStartX/Y/ZCur of the last point (points >= 1 only)
    int valid, has_velo;
};

struct FeEndImu {                // IMU terms of TransformToEnd (FA:927-950); all zero angles / shifts without IMU messages
    float cs_start[6];           // cos / sin of imuRollStart, imuPitchStart, imuYawStart
    float shift_from_start[3];
    float last[3];               // imuRollLast, imuPitchLast, imuYawLast
};

struct FeView {
    const float4 *cloud_in; float4 *cloud_adj;
    int n, n_scan, horizon, cap;
    const int *start_ring, *end_ring;
    const unsigned char *ground; const unsigned *col; const float *range;
    float start_ori, end_ori, ori_diff;
    float *ori; float *curv; int *picked; int *label; unsigned long long *smooth;
    FeHeader *hdr;
    float4 *r_sharp, *r_lsharp, *r_flat, *r_lf_scan, *r_lf_ds;
    int *r_cnt;          // [n_scan][4]: sharp, less sharp, flat, less-flat-scan
    int *r_lf_ds_cnt;    // [n_scan]
    float4 *out[4]; FeHeader *out_hdr;
    FeParams prm;
    int seq;
    const FeImu *imu; FeImuOut *imu_out;     // nullptr: no IMU data for this sweep
};

class FeatureExtractor {
public:
    // out_dev_ext: result block of this sweep inside a batch's block (own allocation when null)
    void init(int n_scan, int horizon, cudaStream_t s, unsigned char *out_dev_ext = nullptr);
    static size_t out_block_bytes(int n_scan, int horizon);     // [FeHeader][sharp][less sharp][flat][less flat, n_scan * horizon]
    static size_t input_bytes(int n, int n_scan);
    size_t out_bytes_used(int n) const { return out_off_[3] + sizeof(float4) * (size_t)n; }
    void set_host_out(unsigned char *p) { out_pin_ = p; }      // where the batch's D2H put this sweep's result block
    bool ready() const { return n_scan_ > 0; }
    int n_scan() const { return n_scan_; }
    int horizon() const { return horizon_; }
    void release();
    // the pieces of extract(), for the batch: fill the pinned input block (returns the sweep's table entry), send it,
    // the five launches over a device table of sweeps, fetch the result block
    FeView stage(const float *cloud32, int n, const int *start_ring, const int *end_ring, float start_ori, float end_ori,
                 float ori_diff, const unsigned char *ground, const unsigned *col, const float *range,
                 unsigned char *hp_ext = nullptr, unsigned char *dp_ext = nullptr);     // ext: the sweep's place in a batch block
    void copy_in(cudaStream_t s);
    void copy_out(cudaStream_t s);
    static int launch(const FeView *table_dev, int count, int n_max, int n_scan, int horizon, const SmallJob *jobs_dev,
                      cudaStream_t s, bool imu = false);
    const std::vector<SmallJob> &jobs_host() const { return jobs_host_; }
    // stages the sweep (host pointers, cloud with a 32 B stride), enqueues everything; returns kernel launches
    int extract(const float *cloud32, int n, const int *start_ring, const int *end_ring, float start_ori, float end_ori,
                float ori_diff, const unsigned char *ground, const unsigned *col, const float *range, cudaStream_t s);
    // the same for a sweep that is already on the device (output of projection.cu): only the table entry (FE_HEAD bytes) goes up
    int extract_dev(const float4 *cloud_dev, int n, const int *start_ring_dev, const int *end_ring_dev, float start_ori, float end_ori,
                    float ori_diff, const unsigned char *ground_dev, const unsigned *col_dev, const float *range_dev, cudaStream_t s);
    // after the stream has been synchronised
    const int *counts() const { return reinterpret_cast<const FeHeader *>(out_pin_)->counts; }
    const int *phase_cycles() const { return reinterpret_cast<const FeHeader *>(out_pin_)->pad; }   // sort, picks (slowest ring), then prof[8]
    const float4 *host_cloud(int which) const;
    const float4 *dev_cloud(int which) const
    {
        return which == 4 ? cloud_adj_.p : reinterpret_cast<const float4 *>(out_dev_ + out_off_[which]);
    }
    int n_points() const { return n_; }
    // TransformToEnd (FA:885-953) of cornerPointsLessSharp / surfPointsLessFlat into the given device clouds
    int transform_to_end(const float T[6], float4 *corner_out, float4 *surf_out, cudaStream_t s, const FeEndImu *imu = nullptr);
    // IMU ring buffers for the NEXT extract / extract_dev (nullptr or pointer_last < 0: none); imu_out(): after the stream
    // has been synchronised
    void set_imu(const FeImu *imu_host);
    const FeImuOut &imu_out() const { return *pin_imu_out_.p; }
    bool has_imu_out() const { return pin_imu_out_.p != nullptr; }
    void get_state(float *curv, int *picked, int *label, int n, cudaStream_t s);
    FeParams prm{ 0.1f, 0.1f, 0.1f, 0.2f };
private:
    int n_scan_ = 0, horizon_ = 0, cap_ = 0, n_ = 0, seq_ = 0, staged_ = 0;
    size_t staged_bytes_ = 0;
    std::vector<SmallJob> jobs_host_;
    PinnedBuf<unsigned char> pin_in_[2]; cudaEvent_t in_ev_[2] = { nullptr, nullptr }; bool in_busy_[2] = { false, false };
    int ring_ = 0;
    DevBuf<unsigned char> in_dev_;
    DevBuf<float4> cloud_adj_, r_sharp_, r_lsharp_, r_flat_, r_lf_scan_, r_lf_ds_;
    DevBuf<unsigned char> out_block_;   // [FeHeader, 64 B][sharp][less sharp][flat][less flat]
    DevBuf<float> ori_, curv_; DevBuf<int> picked_, label_, r_cnt_, r_lf_ds_cnt_;
    DevBuf<unsigned long long> smooth_;
    DevBuf<FeHeader> hdr_;
    DevBuf<SmallJob> jobs_;
    PinnedBuf<unsigned char> pin_out_;
    unsigned char *out_dev_ = nullptr, *out_pin_ = nullptr;     // own buffers or the batch's
    size_t out_off_[4] = { 0, 0, 0, 0 };
    DevBuf<FeImu> imu_dev_; PinnedBuf<FeImu> pin_imu_; DevBuf<FeImuOut> imu_out_dev_; PinnedBuf<FeImuOut> pin_imu_out_;
    bool imu_set_ = false;
};

struct FeSweepHost {     // one sweep as the caller holds it (llb_segmented_cloud)
    const float *cloud32; int n; const int *start_ring, *end_ring; float start_ori, end_ori, ori_diff;
    const unsigned char *ground; const unsigned *col; const float *range;
};

// S independent sequences: one FeatureExtractor state per slot, ONE set of five launches per step (blockIdx.y = slot)
class FeatureBatch {
public:
    void init(int slots, int n_scan, int horizon, cudaStream_t s);
    void release();
    bool ready() const { return n_scan_ > 0; }
    int slots() const { return (int)ext_.size(); }
    int extract(const FeSweepHost *sweeps, cudaStream_t s);      // returns kernel launches
    FeatureExtractor &slot(int i) { return ext_[i]; }
private:
    std::vector<FeatureExtractor> ext_;
    DevBuf<SmallJob> jobs_;
    // one input block per step: [table of FeView][slot 0 input][slot 1 input]..., one H2D; one result block
    // [slots][stride], fetched with one 2-D D2H of the used part of every slot
    PinnedBuf<unsigned char> pin_in_[2]; cudaEvent_t in_ev_[2] = { nullptr, nullptr }; bool in_busy_[2] = { false, false };
    int ring_ = 0;
    DevBuf<unsigned char> in_dev_, out_dev_;
    PinnedBuf<unsigned char> pin_out_;
    size_t out_stride_ = 0;
    int n_scan_ = 0, horizon_ = 0;
};

}  // namespace llb
