// ref_ip.cpp — ORACLE tier B (test infrastructure): compiles the UNMODIFIED reference
// /root/reference/LeGO-LOAM/src/imageProjection.cpp against shim/llref_shim.hpp (see ref_mo.cpp).  It turns a raw sweep
// into what featureAssociation receives - segmented cloud, cloud_info, outlier cloud - so that the feature-extraction
// tests and golden vectors can start from raw lidar points the way the node does (SURVEY 8(f)-3 is the next row).
#include "shim/llref_shim.hpp"
#define private public
#define main ref_ip_node_main
#include "imageProjection.cpp"
#undef main
#undef private

namespace {
int dump(const pcl::PointCloud<PointType>::Ptr &c, llo_point *out, int cap)
{
    int n = (int)c->points.size();
    for (int i = 0; i < n && i < cap; i++) { out[i].x = c->points[i].x; out[i].y = c->points[i].y; out[i].z = c->points[i].z; out[i].intensity = c->points[i].intensity; }
    return n;
}
}  // namespace

extern "C" {
void *ref_ip_create() { return new ImageProjection(); }
void ref_ip_destroy(void *h) { delete (ImageProjection *)h; }
int ref_ip_n_scan() { return N_SCAN; }
int ref_ip_horizon_scan() { return Horizon_SCAN; }
// cloudHandler IP:181-197 without its last two steps (publishCloud, resetParameters): the results stay in the members
void ref_ip_process(void *h, const llo_point *pts, const unsigned short *ring, int n)
{
    ImageProjection *ip = (ImageProjection *)h;
    ip->resetParameters();
    sensor_msgs::PointCloud2::Ptr msg(new sensor_msgs::PointCloud2());
    msg->xyzi.resize((size_t)n * 4); msg->ring.assign(ring, ring + n);
    for (int i = 0; i < n; i++) { msg->xyzi[4 * i] = pts[i].x; msg->xyzi[4 * i + 1] = pts[i].y; msg->xyzi[4 * i + 2] = pts[i].z; msg->xyzi[4 * i + 3] = pts[i].intensity; }
    ip->copyPointCloud(msg);
    ip->findStartEndAngle();
    ip->projectPointCloud();
    ip->groundRemoval();
    ip->cloudSegmentation();
}
int ref_ip_get_cloud(void *h, int which, llo_point *out, int cap)
{   // 0 segmentedCloud, 1 outlierCloud, 2 groundCloud, 3 fullCloud
    ImageProjection *ip = (ImageProjection *)h;
    return dump(which == 0 ? ip->segmentedCloud : which == 1 ? ip->outlierCloud : which == 2 ? ip->groundCloud : ip->fullCloud, out, cap);
}
void ref_ip_get_info(void *h, int *start_ring, int *end_ring, float ori[3], unsigned char *ground, unsigned *col, float *range, int n)
{
    ImageProjection *ip = (ImageProjection *)h;
    for (int i = 0; i < N_SCAN; i++) { start_ring[i] = ip->segMsg.startRingIndex[i]; end_ring[i] = ip->segMsg.endRingIndex[i]; }
    ori[0] = ip->segMsg.startOrientation; ori[1] = ip->segMsg.endOrientation; ori[2] = ip->segMsg.orientationDiff;
    for (int i = 0; i < n; i++) { ground[i] = ip->segMsg.segmentedCloudGroundFlag[i]; col[i] = ip->segMsg.segmentedCloudColInd[i]; range[i] = ip->segMsg.segmentedCloudRange[i]; }
}
void ref_ip_get_images(void *h, float *range_mat, signed char *ground_mat, int *label_mat)
{
    ImageProjection *ip = (ImageProjection *)h;
    for (int i = 0; i < N_SCAN; i++)
        for (int j = 0; j < Horizon_SCAN; j++) {
            range_mat[i * Horizon_SCAN + j] = ip->rangeMat.at<float>(i, j);
            ground_mat[i * Horizon_SCAN + j] = ip->groundMat.at<int8_t>(i, j);
            label_mat[i * Horizon_SCAN + j] = ip->labelMat.at<int>(i, j);
        }
}
}  // extern "C"
