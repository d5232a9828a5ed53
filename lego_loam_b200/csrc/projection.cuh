// projection.cuh — K8: imageProjection on the device: the reference's imageProjection.cpp (= IP)
//   findStartEndAngle IP:199-211, projectPointCloud IP:213-257 (useCloudRing), groundRemoval IP:259-310,
//   cloudSegmentation IP:312-368 with labelComponents IP:370-448.
// The reference is a chain of sequential loops; the data-parallel form used here was checked on the CPU against the
// compiled reference first (tests/test_projection_parallel_form.py):
//   * projection: "the LAST point of the cloud that falls into a pixel wins" = atomicMax of the point index per pixel;
//   * ground: the column sweep over the rows 0 .. groundScanInd (a later, invalid row pair overwrites an earlier mark)
//     is kept as it is, one thread per column;
//   * labelComponents' BFS evaluates a SYMMETRIC edge predicate on the (column-wrapped) 4-neighbourhood, so its segments
//     are the connected components of that edge set: lock-free union-find with the smaller index as the root (the root
//     is then the raster-first pixel, the BFS seed).  A segment is kept when it has >= 30 points, or >= 5 points on >= 3
//     rows where the seed's own row only counts through another point of the segment (the reference sets lineCountFlag
//     for pushed neighbours only).  Label numbers = 1 + the number of kept segments with an earlier seed;
//   * cloudSegmentation: keep flags per pixel, two exclusive scans in raster order, scatter.
// Discrete decisions hang on atan2f (column index IP:236-238, ground angle IP:284-286, segment angle IP:414-416):
// glibc's atan2f is restated operation for operation (glibc_atan2f.cuh); sinf / cosf of the two constant segment angles
// are taken on the host.  Launches per sweep: 6 (+ one H2D of the raw sweep).
#pragma once
#include "common.cuh"

namespace llb {

struct IpParams {
    int n_scan, horizon, ground_scan_ind;
    float ang_res_x, ang_res_y;
    float sensor_min_range, sensor_mount_angle, segment_theta;       // UT:111-113
    int valid_point_num, valid_line_num;                             // UT:114-115
    float sin_ax, cos_ax, sin_ay, cos_ay;                            // sinf / cosf of segmentAlphaX / segmentAlphaY (UT:116-117)
};

struct IpHeader {                  // device-resident results that are scalars
    int n_seg, n_outlier, n_labels, pad;
    float start_ori, end_ori, ori_diff, pad2;
};

struct IpView {                    // everything the kernels of one sweep need
    IpParams prm;
    const float *cloud32;          // raw sweep, pcl::PointXYZI(R) stride of 8 floats (x y z _ intensity ...)
    const unsigned short *ring;
    int n;
    int *winner;                   // [N*H] index of the last point of the cloud in the pixel, -1: none (kept at -1 between sweeps)
    float4 *full;                  // fullCloud: x y z, intensity = row + col / 10000; intensity -1: no return
    float *range_mat; signed char *ground_mat; int *label_mat;
    int *parent, *root, *cnt; unsigned *rowmask;   // union-find parents, final roots; per root: segment size, rows of its
                                   // non-seed points [N*H][4]
    int *number;                   // label number of a kept seed
    float4 *seg, *outlier; unsigned char *ground_flag; unsigned *col_ind; float *seg_range;
    int *start_ring, *end_ring;
    IpHeader *hdr;
};

class ImageProjector {
public:
    void init(int n_scan, int horizon, float ang_res_x, float ang_res_y, int ground_scan_ind, cudaStream_t s);
    void release();
    bool ready() const { return prm_.n_scan > 0; }
    const IpParams &params() const { return prm_; }
    // cloudHandler IP:181-197 without the publishing: uploads the raw sweep (host, 32 B stride) and its ring channel,
    // enqueues the six kernels and the read-back of the header; returns kernel launches
    int process(const float *cloud32_host, const unsigned short *ring_host, int n, cudaStream_t s);
    const IpHeader &header() const { return *pin_hdr_.p; }      // after the stream has been synchronised
    const int *start_ring_host() const { return pin_rings_.p; }
    const int *end_ring_host() const { return pin_rings_.p + prm_.n_scan; }
    // device-resident results
    const float4 *seg_dev() const { return seg_.p; }
    const float4 *outlier_dev() const { return outlier_.p; }
    const unsigned char *ground_flag_dev() const { return ground_flag_.p; }
    const unsigned *col_ind_dev() const { return col_ind_.p; }
    const float *seg_range_dev() const { return seg_range_.p; }
    const int *start_ring_dev() const { return start_ring_.p; }
    const int *end_ring_dev() const { return end_ring_.p; }
    const float *range_mat_dev() const { return range_mat_.p; }
    const signed char *ground_mat_dev() const { return ground_mat_.p; }
    const int *label_mat_dev() const { return label_mat_.p; }

private:
    IpParams prm_{};
    int cap_ = 0;
    PinnedBuf<unsigned char> pin_in_[2]; cudaEvent_t in_ev_[2] = { nullptr, nullptr }; bool in_busy_[2] = { false, false };
    int ring_pos_ = 0;
    DevBuf<unsigned char> in_dev_;
    DevBuf<int> winner_, label_mat_, parent_, root_, cnt_, number_, start_ring_, end_ring_;
    DevBuf<unsigned> rowmask_, col_ind_;
    DevBuf<float4> full_, seg_, outlier_;
    DevBuf<float> range_mat_, seg_range_;
    DevBuf<signed char> ground_mat_;
    DevBuf<unsigned char> ground_flag_;
    DevBuf<IpHeader> hdr_;
    PinnedBuf<IpHeader> pin_hdr_;
    PinnedBuf<int> pin_rings_;     // startRingIndex | endRingIndex of the last sweep
};

}  // namespace llb
