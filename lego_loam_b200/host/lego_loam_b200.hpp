// lego_loam_b200.hpp — C++ host adapters: the reference's own class and member-function names
// on top of the C ABI (include/llb200.h).
//
// The reference has no plugin interface; its boundary is the set of member functions of two
// monolithic classes that communicate through members:
//   class mapOptimization   (LeGO-LOAM/src/mapOptmization.cpp:49)
//   class FeatureAssociation (LeGO-LOAM/src/featureAssociation.cpp:37)
// The classes below keep those method names / signatures and the pcl::PointCloud<PointType>::Ptr
// members the hot path reads and writes, so a maintainer swaps the bodies of the hot methods for a
// call into the matching adapter method (INTEGRATION.md shows the exact patch).
//
// Error behaviour mirrors the reference: no exceptions, failures are silent skips that leave
// the pose untouched (guards MO:1331, MO:1238, FA:1668); the last C-ABI status is kept in
// `last_status` for callers that want to look.  There is no CPU fallback: if the CUDA library
// cannot create a context the constructor throws std::runtime_error.
#pragma once
#include "../../include/llb200.h"
#include "compat/pcl/llb_pcl_compat.h"

#include <cmath>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace lego_loam_b200 {

typedef pcl::PointXYZI PointType;                       // UT:51
typedef pcl::PointCloud<PointType> Cloud;

static_assert(sizeof(PointType) == sizeof(llb_point), "llb_point must alias pcl::PointXYZI");

inline const llb_point *as_llb(const Cloud &c) { return reinterpret_cast<const llb_point *>(c.points.data()); }
inline llb_point *as_llb(Cloud &c) { return reinterpret_cast<llb_point *>(c.points.data()); }

// ---------------------------------------------------------------------------------------------
// mapOptimization: downsampleCurrentScan / cornerOptimization / surfOptimization /
// LMOptimization / scan2MapOptimization (MO:1067-1350) + the map voxel tail MO:1057-1064
// ---------------------------------------------------------------------------------------------
class mapOptimization {
public:
    // members with the reference's names (MO:109-126, MO:173-178, MO:202-210)
    Cloud::Ptr laserCloudCornerLast, laserCloudSurfLast, laserCloudOutlierLast;
    Cloud::Ptr laserCloudCornerLastDS, laserCloudSurfLastDS, laserCloudOutlierLastDS;
    Cloud::Ptr laserCloudSurfTotalLast, laserCloudSurfTotalLastDS;
    Cloud::Ptr laserCloudOri, coeffSel;
    Cloud::Ptr laserCloudCornerFromMap, laserCloudSurfFromMap, laserCloudCornerFromMapDS, laserCloudSurfFromMapDS;
    float transformTobeMapped[6], transformSum[6], transformBefMapped[6], transformAftMapped[6];
    bool isDegenerate;
    float matP[36];
    int laserCloudCornerFromMapDSNum, laserCloudSurfFromMapDSNum;
    int laserCloudCornerLastDSNum, laserCloudSurfLastDSNum, laserCloudOutlierLastDSNum, laserCloudSurfTotalLastDSNum;
    // adapter knobs
    bool fetch_downsampled_clouds = true;   // copy the *DS clouds back to the host members (key-frame store needs them)
    int last_status = LLB_OK;
    llb_stats last_stats;

    explicit mapOptimization(int device = 0, const llb_params *params = nullptr)
    {
        for (Cloud::Ptr *p : { &laserCloudCornerLast, &laserCloudSurfLast, &laserCloudOutlierLast, &laserCloudCornerLastDS,
                               &laserCloudSurfLastDS, &laserCloudOutlierLastDS, &laserCloudSurfTotalLast,
                               &laserCloudSurfTotalLastDS, &laserCloudOri, &coeffSel, &laserCloudCornerFromMap,
                               &laserCloudSurfFromMap, &laserCloudCornerFromMapDS, &laserCloudSurfFromMapDS })
            p->reset(new Cloud());
        for (int i = 0; i < 6; i++) transformTobeMapped[i] = transformSum[i] = transformBefMapped[i] = transformAftMapped[i] = 0;
        isDegenerate = false;
        std::memset(matP, 0, sizeof matP);
        std::memset(&last_stats, 0, sizeof last_stats);
        laserCloudCornerFromMapDSNum = laserCloudSurfFromMapDSNum = 0;
        laserCloudCornerLastDSNum = laserCloudSurfLastDSNum = laserCloudOutlierLastDSNum = laserCloudSurfTotalLastDSNum = 0;
        int rc = llb_create(params, device, &ctx_);
        if (rc != LLB_OK) throw std::runtime_error("lego_loam_b200: llb_create failed (no CUDA device / no CPU fallback)");
    }
    ~mapOptimization() { llb_destroy(ctx_); }
    mapOptimization(const mapOptimization &) = delete;
    mapOptimization &operator=(const mapOptimization &) = delete;

    // tail of extractSurroundingKeyFrames, MO:1057-1064: two voxel filters on the raw local map.
    // The DS map stays on the device (index built); it is copied back only if asked.
    void downsampleSurroundingMap(bool fetch = false)
    {
        last_status = llb_map_set_raw(ctx_, as_llb(*laserCloudCornerFromMap), (int)laserCloudCornerFromMap->size(),
                                      as_llb(*laserCloudSurfFromMap), (int)laserCloudSurfFromMap->size());
        if (last_status != LLB_OK) return;
        llb_map_get_ds(ctx_, 0, nullptr, 0, &laserCloudCornerFromMapDSNum);
        llb_map_get_ds(ctx_, 1, nullptr, 0, &laserCloudSurfFromMapDSNum);
        map_on_device_ = true;
        if (fetch) {
            fetch_cloud(&llb_map_get_ds, 0, *laserCloudCornerFromMapDS);
            fetch_cloud(&llb_map_get_ds, 1, *laserCloudSurfFromMapDS);
        }
    }

    // ---- device-resident key-frame store ----
    // cloud part of saveKeyFramesAndFactor, MO:1443-1453: the DS clouds of the current sweep (already in HBM after
    // downsampleCurrentScan) become key-frame number `return value`; nothing is copied to the host
    int saveKeyFrameClouds()
    {
        int id = -1;
        last_status = llb_keyframe_add(ctx_, &id);
        return last_status == LLB_OK ? id : -1;
    }
    // cloud part of extractSurroundingKeyFrames, MO:1033-1064 (or the recent-frames branch MO:962-1001): ids in the
    // order of surroundingExistingKeyPosesID, poses6d[6*i ..] = cloudKeyPoses6D[ids[i]] {roll, pitch, yaw, x, y, z}.
    // transformPointCloud + concatenation + the two map voxel filters + the index build, all on the device.
    void assembleSurroundingMap(const std::vector<int> &ids, const std::vector<float> &poses6d)
    {
        last_status = llb_map_assemble(ctx_, ids.data(), poses6d.data(), (int)ids.size());
        if (last_status != LLB_OK) return;
        llb_map_get_ds(ctx_, 0, nullptr, 0, &laserCloudCornerFromMapDSNum);
        llb_map_get_ds(ctx_, 1, nullptr, 0, &laserCloudSurfFromMapDSNum);
        map_on_device_ = true;
    }

    // ---- loop closure + global map (loopClosureEnableFlag UT:104) ----
    // Members with the reference's names (MO:140-146, MO:151-154).  The host keeps the key-pose searches and the pose
    // graph (MO:822-837, MO:919-944); the device does the cloud work on its key-frame store.
    Cloud::Ptr latestSurfKeyFrameCloud{ new Cloud() }, nearHistorySurfKeyFrameCloudDS{ new Cloud() }, globalMapKeyFramesDS{ new Cloud() };
    int closestHistoryFrameID = -1, latestFrameIDLoopCloure = -1;
    llb_icp_result last_icp{};
    // cloud part of detectLoopClosure MO:838-861 once the caller has found closestHistoryFrameID (MO:822-837):
    // poses6d(i, out) must fill out[6] with cloudKeyPoses6D[i] as {roll, pitch, yaw, x, y, z}; historyKeyframeSearchNum = 25 (UT:133)
    template <typename PoseOf>
    bool detectLoopClosureClouds(int closest_id, int latest_id, PoseOf poses6d, int history_num = 25, bool fetch = false)
    {
        closestHistoryFrameID = closest_id; latestFrameIDLoopCloure = latest_id;
        std::vector<int> ids; std::vector<float> hp;
        for (int j = -history_num; j <= history_num; ++j) {                         // MO:853-855
            if (closest_id + j < 0 || closest_id + j > latest_id) continue;
            ids.push_back(closest_id + j);
            float p[6]; poses6d(closest_id + j, p); hp.insert(hp.end(), p, p + 6);
        }
        int counts[2] = { 0, 0 };
        float lp[6]; poses6d(latest_id, lp);
        last_status = llb_loop_set_clouds(ctx_, latest_id, lp, ids.data(), hp.data(), (int)ids.size(), 0.4f, counts);
        if (last_status != LLB_OK) return false;
        if (fetch) {
            fetch_cloud(&llb_loop_get_cloud, 0, *latestSurfKeyFrameCloud);
            fetch_cloud(&llb_loop_get_cloud, 2, *nearHistorySurfKeyFrameCloudDS);
        }
        return true;
    }
    // the ICP of performLoopClosure MO:892-904: icp.align + hasConverged + getFitnessScore; true when the reference would
    // go on to add the factor (MO:904); last_icp.T is icp.getFinalTransformation() (row-major)
    bool performLoopClosureICP(bool clouds_from_members = false, float historyKeyframeFitnessScore = 0.3f)
    {
        if (clouds_from_members) {
            last_status = llb_loop_set_clouds_host(ctx_, as_llb(*latestSurfKeyFrameCloud), (int)latestSurfKeyFrameCloud->size(),
                                                   as_llb(*nearHistorySurfKeyFrameCloudDS), (int)nearHistorySurfKeyFrameCloudDS->size());
            if (last_status != LLB_OK) return false;
        }
        last_status = llb_loop_icp(ctx_, nullptr, &last_icp);
        if (last_status != LLB_OK) return false;
        return last_icp.has_converged && !(last_icp.fitness_score > (double)historyKeyframeFitnessScore);
    }
    // cloud part of publishGlobalMap MO:780-788 for the key-frames the caller selected (MO:766-778)
    void publishGlobalMapClouds(const std::vector<int> &ids, const std::vector<float> &poses6d)
    {
        int n = 0;
        last_status = llb_global_map_assemble(ctx_, ids.data(), poses6d.data(), (int)ids.size(), 0.4f, &n);
        if (last_status == LLB_OK) fetch_cloud(&llb_loop_get_cloud, 3, *globalMapKeyFramesDS);
    }

    void downsampleCurrentScan()                                   // MO:1067
    {
        last_status = llb_scan_set(ctx_, as_llb(*laserCloudCornerLast), (int)laserCloudCornerLast->size(),
                                   as_llb(*laserCloudSurfLast), (int)laserCloudSurfLast->size(),
                                   as_llb(*laserCloudOutlierLast), (int)laserCloudOutlierLast->size());
        if (last_status != LLB_OK) return;
        int counts[4] = { 0, 0, 0, 0 };
        last_status = llb_downsample_current_scan(ctx_, counts);
        if (last_status != LLB_OK) return;
        laserCloudCornerLastDSNum = counts[0]; laserCloudSurfLastDSNum = counts[1];
        laserCloudOutlierLastDSNum = counts[2]; laserCloudSurfTotalLastDSNum = counts[3];
        if (fetch_downsampled_clouds) {
            fetch_cloud(&llb_scan_get_ds, 0, *laserCloudCornerLastDS);
            fetch_cloud(&llb_scan_get_ds, 1, *laserCloudSurfLastDS);
            fetch_cloud(&llb_scan_get_ds, 2, *laserCloudOutlierLastDS);
            fetch_cloud(&llb_scan_get_ds, 3, *laserCloudSurfTotalLastDS);
        }
    }

    // cornerOptimization + surfOptimization + LMOptimization are one fused device iteration; the three
    // reference entry points are kept: the first two stage the iteration, LMOptimization runs it.
    void cornerOptimization(int /*iterCount*/) { staged_ |= 1; }    // MO:1093
    void surfOptimization(int /*iterCount*/) { staged_ |= 2; }      // MO:1176
    bool LMOptimization(int iterCount)                              // MO:1229, returns true when converged
    {
        // the device iteration is the fusion of the three reference calls: it needs BOTH searches staged (the reference
        // order MO:1339-1343); laserCloudOri / coeffSel are filled here, not by the two stagers
        if (staged_ != 3) { staged_ = 0; last_status = LLB_ERR_STATE; return false; }
        staged_ = 0;
        if (!ensure_map()) return false;
        int conv = 0, n = 0;
        last_status = llb_s2m_iterate(ctx_, transformTobeMapped, iterCount, &conv, &n);
        if (last_status != LLB_OK) return false;
        pull_correspondences();
        pull_degeneracy();
        return conv != 0;
    }

    void scan2MapOptimization()                                     // MO:1329
    {
        if (laserCloudCornerFromMapDSNum > 10 && laserCloudSurfFromMapDSNum > 100) {
            if (!ensure_map()) return;                              // MO:1333-1334 (index build)
            last_status = llb_s2m_optimize(ctx_, transformTobeMapped, &last_stats);   // MO:1336-1346
            if (last_status != LLB_OK) return;
            pull_degeneracy();
            transformUpdate();                                      // MO:1348
        }
    }

    // imuHandler MO:643-652 after the quaternion -> roll / pitch conversion (tf stays with the caller)
    void imuHandler(double stamp, double roll, double pitch)
    {
        imuPointerLast = (imuPointerLast + 1) % imuQueLength;
        imuTime[imuPointerLast] = stamp;
        imuRoll[imuPointerLast] = (float)roll;
        imuPitch[imuPointerLast] = (float)pitch;
    }

    // MO:463-496: with IMU messages roll and pitch are pulled towards the IMU's (complementary blend 0.998 / 0.002,
    // IMU values interpolated at the end of the sweep); host arithmetic, O(1)
    void transformUpdate()
    {
        if (imuPointerLast >= 0) {
            float imuRollLast = 0, imuPitchLast = 0;
            while (imuPointerFront != imuPointerLast) {
                if (timeLaserOdometry + scanPeriod < imuTime[imuPointerFront]) break;
                imuPointerFront = (imuPointerFront + 1) % imuQueLength;
            }
            if (timeLaserOdometry + scanPeriod > imuTime[imuPointerFront]) {
                imuRollLast = imuRoll[imuPointerFront];
                imuPitchLast = imuPitch[imuPointerFront];
            } else {
                const int back = (imuPointerFront + imuQueLength - 1) % imuQueLength;
                const float ratioFront = (float)((timeLaserOdometry + scanPeriod - imuTime[back]) / (imuTime[imuPointerFront] - imuTime[back]));
                const float ratioBack = (float)((imuTime[imuPointerFront] - timeLaserOdometry - scanPeriod) / (imuTime[imuPointerFront] - imuTime[back]));
                imuRollLast = imuRoll[imuPointerFront] * ratioFront + imuRoll[back] * ratioBack;
                imuPitchLast = imuPitch[imuPointerFront] * ratioFront + imuPitch[back] * ratioBack;
            }
            transformTobeMapped[0] = (float)(0.998 * transformTobeMapped[0] + 0.002 * imuPitchLast);
            transformTobeMapped[2] = (float)(0.998 * transformTobeMapped[2] + 0.002 * imuRollLast);
        }
        for (int i = 0; i < 6; i++) { transformBefMapped[i] = transformSum[i]; transformAftMapped[i] = transformTobeMapped[i]; }
    }
    static constexpr int imuQueLength = 200;                       // UT:109
    static constexpr float scanPeriod = 0.1f;                      // UT:107 (a float: it is promoted to double in the time sums)
    double timeLaserOdometry = 0;                                  // MO:160
    double imuTime[imuQueLength] = {};                             // MO:184
    float imuRoll[imuQueLength] = {}, imuPitch[imuQueLength] = {}; // MO:185-186
    int imuPointerFront = 0, imuPointerLast = -1;                  // MO:181-182

    // call when laserCloud*FromMapDS were (re)filled on the host by the caller
    void mapChanged() { map_on_device_ = false; }
    llb_ctx *context() { return ctx_; }

private:
    typedef int (*get_fn)(llb_ctx *, int, llb_point *, int, int *);
    void fetch_cloud(get_fn fn, int which, Cloud &out)
    {
        int n = 0;
        if (fn(ctx_, which, nullptr, 0, &n) != LLB_OK) return;
        out.resize(n);
        if (n > 0) fn(ctx_, which, as_llb(out), n, &n);
    }
    bool ensure_map()
    {
        if (map_on_device_) return true;
        laserCloudCornerFromMapDSNum = (int)laserCloudCornerFromMapDS->size();
        laserCloudSurfFromMapDSNum = (int)laserCloudSurfFromMapDS->size();
        last_status = llb_map_set_ds(ctx_, as_llb(*laserCloudCornerFromMapDS), laserCloudCornerFromMapDSNum,
                                     as_llb(*laserCloudSurfFromMapDS), laserCloudSurfFromMapDSNum);
        map_on_device_ = (last_status == LLB_OK);
        return map_on_device_;
    }
    void pull_correspondences()
    {
        int n = 0;
        if (llb_get_correspondences(ctx_, nullptr, nullptr, 0, &n) != LLB_OK) { laserCloudOri->clear(); coeffSel->clear(); return; }
        laserCloudOri->resize(n); coeffSel->resize(n);
        if (n > 0) llb_get_correspondences(ctx_, as_llb(*laserCloudOri), as_llb(*coeffSel), n, &n);
    }
    void pull_degeneracy()
    {
        int d = 0;
        if (llb_get_degeneracy(ctx_, &d, matP) == LLB_OK) isDegenerate = d != 0;
    }
    llb_ctx *ctx_ = nullptr;
    bool map_on_device_ = false;
    int staged_ = 0;
};

// cloud_msgs::cloud_info (cloud_msgs/msg/cloud_info.msg) as imageProjection fills it (IP:312-368)
struct cloud_info {
    std::vector<int32_t> startRingIndex, endRingIndex;
    float startOrientation = 0, endOrientation = 0, orientationDiff = 0;
    std::vector<uint8_t> segmentedCloudGroundFlag;
    std::vector<uint32_t> segmentedCloudColInd;
    std::vector<float> segmentedCloudRange;
};

// ---------------------------------------------------------------------------------------------
// ImageProjection: cloudHandler's steps (IP:181-197): findStartEndAngle, projectPointCloud, groundRemoval,
// cloudSegmentation (imageProjection.cpp:199-368 with labelComponents :370-448).  The reference's members keep their
// names; the four steps run as one device pass when cloudSegmentation() is reached (they only communicate through the
// images, which stay on the device).  laserCloudInRing: the ring channel of the sweep (useCloudRing, UT:60).
// ---------------------------------------------------------------------------------------------
class ImageProjection {
public:
    Cloud::Ptr laserCloudIn;                                       // IP:48 (NaN-free, IP:170)
    std::vector<uint16_t> laserCloudInRing;                        // IP:49: points[i].ring
    Cloud::Ptr segmentedCloud, outlierCloud;                       // IP:55, IP:57
    cloud_info segMsg;                                             // IP:68
    int last_status = LLB_OK;

    // sensor block of utility.h (UT:62-84): N_SCAN, Horizon_SCAN, ang_res_x, ang_res_y, groundScanInd
    ImageProjection(int n_scan, int horizon_scan, float ang_res_x, float ang_res_y, int ground_scan_ind, int device = 0,
                    const llb_params *params = nullptr)
    {
        for (Cloud::Ptr *p : { &laserCloudIn, &segmentedCloud, &outlierCloud }) p->reset(new Cloud());
        if (llb_create(params, device, &ctx_) != LLB_OK)
            throw std::runtime_error("lego_loam_b200: llb_create failed (no CUDA device / no CPU fallback)");
        last_status = llb_projection_init(ctx_, n_scan, horizon_scan, ang_res_x, ang_res_y, ground_scan_ind);
        n_scan_ = n_scan;
    }
    ~ImageProjection() { llb_destroy(ctx_); }
    ImageProjection(const ImageProjection &) = delete;
    ImageProjection &operator=(const ImageProjection &) = delete;

    void findStartEndAngle() { staged_ |= 1; }                     // IP:199
    void projectPointCloud() { staged_ |= 2; }                     // IP:213
    void groundRemoval() { staged_ |= 4; }                         // IP:259
    // fetch = copy segmentedCloud / outlierCloud / segMsg to the members (not needed when the next stage is
    // llb_projection_to_features on the same context)
    void cloudSegmentation(bool fetch = true)                      // IP:312
    {
        if (staged_ != 7 || laserCloudInRing.size() < laserCloudIn->size()) { staged_ = 0; last_status = LLB_ERR_STATE; return; }
        staged_ = 0;
        int ns = 0, no = 0;
        last_status = llb_projection_process(ctx_, as_llb(*laserCloudIn), laserCloudInRing.data(), (int)laserCloudIn->size(),
                                             &ns, &no, nullptr);
        if (last_status != LLB_OK || !fetch) return;
        segmentedCloud->resize(ns); outlierCloud->resize(no);
        if (ns > 0) last_status = llb_projection_get_cloud(ctx_, 0, as_llb(*segmentedCloud), ns, &ns);
        if (no > 0 && last_status == LLB_OK) last_status = llb_projection_get_cloud(ctx_, 1, as_llb(*outlierCloud), no, &no);
        segMsg.startRingIndex.assign(n_scan_, 0); segMsg.endRingIndex.assign(n_scan_, 0);
        segMsg.segmentedCloudGroundFlag.assign(ns, 0); segMsg.segmentedCloudColInd.assign(ns, 0); segMsg.segmentedCloudRange.assign(ns, 0.f);
        float ori[3] = { 0, 0, 0 };
        if (last_status == LLB_OK)
            last_status = llb_projection_get_info(ctx_, segMsg.startRingIndex.data(), segMsg.endRingIndex.data(), ori,
                                                  segMsg.segmentedCloudGroundFlag.data(), segMsg.segmentedCloudColInd.data(),
                                                  segMsg.segmentedCloudRange.data(), ns > 0 ? ns : 1);
        segMsg.startOrientation = ori[0]; segMsg.endOrientation = ori[1]; segMsg.orientationDiff = ori[2];
    }
    llb_ctx *context() { return ctx_; }

private:
    llb_ctx *ctx_ = nullptr;
    int staged_ = 0, n_scan_ = 0;
};

// ---------------------------------------------------------------------------------------------
// FeatureAssociation: adjustDistortion / calculateSmoothness / markOccludedPoints / extractFeatures (FA:491-784),
// findCorresponding{Corner,Surf}Features / calculateTransformation{Surf,Corner}
// / updateTransformation (FA:1044-1478, FA:1666-1695)
// ---------------------------------------------------------------------------------------------
class FeatureAssociation {
public:
    Cloud::Ptr segmentedCloud;                                     // FA:52
    cloud_info segInfo;                                            // FA:75
    Cloud::Ptr cornerPointsSharp, surfPointsFlat;                  // FA:56-58
    Cloud::Ptr cornerPointsLessSharp, surfPointsLessFlat;          // FA:57-59
    Cloud::Ptr laserCloudCornerLast, laserCloudSurfLast;           // FA:161-162
    Cloud::Ptr laserCloudOri, coeffSel;                            // FA:163-164
    float transformCur[6];                                         // FA:154
    int laserCloudCornerLastNum, laserCloudSurfLastNum;            // FA:142-143
    bool isDegenerate;                                             // FA:179
    float matP[9];
    int last_status = LLB_OK;
    llb_stats stats_surf, stats_corner;

    explicit FeatureAssociation(int device = 0, const llb_params *params = nullptr)
    {
        for (Cloud::Ptr *p : { &cornerPointsSharp, &surfPointsFlat, &laserCloudCornerLast, &laserCloudSurfLast, &laserCloudOri, &coeffSel,
                               &segmentedCloud, &cornerPointsLessSharp, &surfPointsLessFlat })
            p->reset(new Cloud());
        for (int i = 0; i < 6; i++) transformCur[i] = 0;
        laserCloudCornerLastNum = laserCloudSurfLastNum = 0;
        isDegenerate = false;
        std::memset(matP, 0, sizeof matP);
        std::memset(&stats_surf, 0, sizeof stats_surf); std::memset(&stats_corner, 0, sizeof stats_corner);
        if (llb_create(params, device, &ctx_) != LLB_OK)
            throw std::runtime_error("lego_loam_b200: llb_create failed (no CUDA device / no CPU fallback)");
    }
    ~FeatureAssociation() { llb_destroy(ctx_); }
    FeatureAssociation(const FeatureAssociation &) = delete;
    FeatureAssociation &operator=(const FeatureAssociation &) = delete;

    // N_SCAN / Horizon_SCAN of the sensor (UT:63-84); allocates the per-point state FA:210-223 on the device
    void initFeatureExtraction(int n_scan, int horizon_scan)
    {
        last_status = llb_features_init(ctx_, n_scan, horizon_scan);
        n_scan_ = last_status == LLB_OK ? n_scan : 0;
    }
    // runFeatureAssociation FA:1827-1833: the four steps run as one device pass when extractFeatures() is reached
    void adjustDistortion() { fe_staged_ |= 1; }                   // FA:491 (IMU branch: when imuHandler has been fed)
    void calculateSmoothness() { fe_staged_ |= 2; }                // FA:621
    void markOccludedPoints() { fe_staged_ |= 4; }                 // FA:643
    void extractFeatures()                                         // FA:680
    {
        if (fe_staged_ != 7) { last_status = LLB_ERR_STATE; return; }
        fe_staged_ = 0;
        // the library reads n_scan ring bounds and one flag / column / range per point: refuse short vectors here
        const size_t np = segmentedCloud->size();
        if (n_scan_ <= 0 || (int)segInfo.startRingIndex.size() < n_scan_ || (int)segInfo.endRingIndex.size() < n_scan_ ||
            segInfo.segmentedCloudGroundFlag.size() < np || segInfo.segmentedCloudColInd.size() < np ||
            segInfo.segmentedCloudRange.size() < np) { last_status = LLB_ERR_INVALID; return; }
        llb_segmented_cloud seg;
        seg.cloud = as_llb(*segmentedCloud); seg.n = (int)segmentedCloud->size();
        seg.start_ring = segInfo.startRingIndex.data(); seg.end_ring = segInfo.endRingIndex.data();
        seg.start_orientation = segInfo.startOrientation; seg.end_orientation = segInfo.endOrientation;
        seg.orientation_diff = segInfo.orientationDiff;
        seg.ground_flag = segInfo.segmentedCloudGroundFlag.data(); seg.col_ind = segInfo.segmentedCloudColInd.data();
        seg.range = segInfo.segmentedCloudRange.data();
        int counts[4] = { 0, 0, 0, 0 };
        if (imuPointerLast >= 0) {                           // FA:525: the IMU branch of adjustDistortion runs on the device
            imu_.time_scan_cur = timeScanCur; imu_.pointer_last = imuPointerLast; imu_.pointer_last_iteration = imuPointerLastIteration;
            last_status = llb_features_set_imu(ctx_, &imu_);
            if (last_status != LLB_OK) return;
        }
        last_status = llb_features_extract(ctx_, &seg, counts, nullptr);
        if (last_status != LLB_OK) return;
        if (imuPointerLast >= 0) pull_imu_sweep();
        imuPointerLastIteration = imuPointerLast;            // FA:616
        Cloud *out[5] = { cornerPointsSharp.get(), cornerPointsLessSharp.get(), surfPointsFlat.get(), surfPointsLessFlat.get(),
                          segmentedCloud.get() };
        for (int k = 0; k < 5; k++) {
            int n = k < 4 ? counts[k] : seg.n;
            out[k]->resize(n);
            if (n > 0) last_status = llb_features_get(ctx_, k, as_llb(*out[k]), n, &n);
        }
        features_on_device_ = true;
        features_pushed_ = false;
    }

    // cloud part of publishCloudsLast (FA:1759-1788): TransformToEnd of the less-sharp / less-flat clouds with transformCur,
    // which become laserCloudCornerLast / laserCloudSurfLast (indexed on the device); fetch = copy them to the members too
    void publishCloudsLast(bool fetch = true)
    {
        llb_imu_end e;                                       // updateImuRollPitchYawStartSinCos FA:317-324 + the members FA:927-950 reads
        e.cs_start[0] = std::cos(imuRollStart); e.cs_start[1] = std::sin(imuRollStart);
        e.cs_start[2] = std::cos(imuPitchStart); e.cs_start[3] = std::sin(imuPitchStart);
        e.cs_start[4] = std::cos(imuYawStart); e.cs_start[5] = std::sin(imuYawStart);
        e.shift_from_start[0] = imuShiftFromStartX; e.shift_from_start[1] = imuShiftFromStartY; e.shift_from_start[2] = imuShiftFromStartZ;
        e.last[0] = imuRollLast; e.last[1] = imuPitchLast; e.last[2] = imuYawLast;
        last_status = llb_features_publish_last_imu(ctx_, transformCur, &e);
        if (last_status != LLB_OK) return;
        Cloud *out[2] = { laserCloudCornerLast.get(), laserCloudSurfLast.get() };
        int n[2] = { 0, 0 };
        for (int k = 0; k < 2; k++) {
            last_status = llb_features_get(ctx_, 5 + k, nullptr, 0, &n[k]);
            if (last_status != LLB_OK) return;
            if (fetch) { out[k]->resize(n[k]); if (n[k] > 0) last_status = llb_features_get(ctx_, 5 + k, as_llb(*out[k]), n[k], &n[k]); }
        }
        laserCloudCornerLastNum = n[0]; laserCloudSurfLastNum = n[1];
    }

    // ---- IMU.  imuHandler FA:417-448 after tf's quaternion -> roll / pitch / yaw (tf is the caller's):
    // gravity compensation, ring-buffer entry, AccumulateIMUShiftAndRotation FA:390-415.  One message at a time and a few
    // dozen flops each: host work; the per-point use of the buffers runs on the device (llb_features_set_imu).
    void imuHandler(double stamp, double roll, double pitch, double yaw, const double linear_acceleration[3],
                    const double angular_velocity[3])
    {
        const float accX = (float)(linear_acceleration[1] - std::sin(roll) * std::cos(pitch) * 9.81);
        const float accY = (float)(linear_acceleration[2] - std::cos(roll) * std::cos(pitch) * 9.81);
        const float accZ = (float)(linear_acceleration[0] + std::sin(pitch) * 9.81);
        imuPointerLast = (imuPointerLast + 1) % imuQueLength;
        const int l = imuPointerLast;
        imu_.time[l] = stamp;
        imu_.roll[l] = (float)roll; imu_.pitch[l] = (float)pitch; imu_.yaw[l] = (float)yaw;
        imuAcc[0][l] = accX; imuAcc[1][l] = accY; imuAcc[2][l] = accZ;
        for (int a = 0; a < 3; a++) imuAngularVelo[a][l] = (float)angular_velocity[a];
        AccumulateIMUShiftAndRotation();
    }
    void AccumulateIMUShiftAndRotation()                     // FA:390
    {
        const int l = imuPointerLast;
        const float roll = imu_.roll[l], pitch = imu_.pitch[l], yaw = imu_.yaw[l];
        float accX = imuAcc[0][l], accY = imuAcc[1][l], accZ = imuAcc[2][l];
        const float x1 = std::cos(roll) * accX - std::sin(roll) * accY;
        const float y1 = std::sin(roll) * accX + std::cos(roll) * accY;
        const float z1 = accZ;
        const float x2 = x1;
        const float y2 = std::cos(pitch) * y1 - std::sin(pitch) * z1;
        const float z2 = std::sin(pitch) * y1 + std::cos(pitch) * z1;
        accX = std::cos(yaw) * x2 + std::sin(yaw) * z2;
        accY = y2;
        accZ = -std::sin(yaw) * x2 + std::cos(yaw) * z2;
        const float acc[3] = { accX, accY, accZ };
        const int b = (l + imuQueLength - 1) % imuQueLength;
        const double timeDiff = imu_.time[l] - imu_.time[b];
        if (timeDiff < scanPeriod) {
            for (int a = 0; a < 3; a++) {
                imu_.shift[a][l] = (float)(imu_.shift[a][b] + imu_.velo[a][b] * timeDiff + acc[a] * timeDiff * timeDiff / 2);
                imu_.velo[a][l] = (float)(imu_.velo[a][b] + acc[a] * timeDiff);
                imu_.angular[a][l] = (float)(imu_.angular[a][b] + imuAngularVelo[a][b] * timeDiff);
            }
        }
    }
    void laserCloudHandlerStamp(double stamp) { timeScanCur = stamp; }         // FA:453
    void updateInitialGuess()                                // FA:1639
    {
        imuPitchLast = imuPitchCur; imuYawLast = imuYawCur; imuRollLast = imuRollCur;
        imuShiftFromStartX = imuShiftFromStartXCur; imuShiftFromStartY = imuShiftFromStartYCur; imuShiftFromStartZ = imuShiftFromStartZCur;
        imuVeloFromStartX = imuVeloFromStartXCur; imuVeloFromStartY = imuVeloFromStartYCur; imuVeloFromStartZ = imuVeloFromStartZCur;
        if (imuAngularFromStartX != 0 || imuAngularFromStartY != 0 || imuAngularFromStartZ != 0) {
            transformCur[0] = -imuAngularFromStartY; transformCur[1] = -imuAngularFromStartZ; transformCur[2] = -imuAngularFromStartX;
        }
        if (imuVeloFromStartX != 0 || imuVeloFromStartY != 0 || imuVeloFromStartZ != 0) {
            transformCur[3] -= imuVeloFromStartX * scanPeriod; transformCur[4] -= imuVeloFromStartY * scanPeriod;
            transformCur[5] -= imuVeloFromStartZ * scanPeriod;
        }
    }
    static constexpr int imuQueLength = LLB_IMU_QUEUE;       // UT:109
    static constexpr float scanPeriod = 0.1f;                // UT:107
    double timeScanCur = 0;                                  // FA:67
    int imuPointerLast = -1, imuPointerLastIteration = 0;    // FA:84-85 (FA:268-269)
    llb_imu_queue imu_ = {};                                 // imuTime, imuRoll/Pitch/Yaw, imuVelo*, imuShift*, imuAngularRotation* FA:105-127
    float imuAcc[3][LLB_IMU_QUEUE] = {}, imuAngularVelo[3][LLB_IMU_QUEUE] = {};   // FA:112-114, FA:124-126 (host only)
    float imuRollStart = 0, imuPitchStart = 0, imuYawStart = 0;                  // FA:87-90
    float imuRollCur = 0, imuPitchCur = 0, imuYawCur = 0;
    float imuVeloXStart = 0, imuVeloYStart = 0, imuVeloZStart = 0, imuShiftXStart = 0, imuShiftYStart = 0, imuShiftZStart = 0;
    float imuShiftFromStartXCur = 0, imuShiftFromStartYCur = 0, imuShiftFromStartZCur = 0;   // never written: ShiftToStartIMU has no caller
    float imuVeloFromStartXCur = 0, imuVeloFromStartYCur = 0, imuVeloFromStartZCur = 0;
    float imuAngularRotationXCur = 0, imuAngularRotationYCur = 0, imuAngularRotationZCur = 0;
    float imuAngularRotationXLast = 0, imuAngularRotationYLast = 0, imuAngularRotationZLast = 0;
    float imuAngularFromStartX = 0, imuAngularFromStartY = 0, imuAngularFromStartZ = 0;
    float imuRollLast = 0, imuPitchLast = 0, imuYawLast = 0;
    float imuShiftFromStartX = 0, imuShiftFromStartY = 0, imuShiftFromStartZ = 0;
    float imuVeloFromStartX = 0, imuVeloFromStartY = 0, imuVeloFromStartZ = 0;

    // replaces the two kdtree->setInputCloud calls of FA:1615-1616 / FA:1786-1787
    void setLastClouds()
    {
        laserCloudCornerLastNum = (int)laserCloudCornerLast->size();
        laserCloudSurfLastNum = (int)laserCloudSurfLast->size();
        last_status = llb_odom_set_last(ctx_, as_llb(*laserCloudCornerLast), laserCloudCornerLastNum,
                                        as_llb(*laserCloudSurfLast), laserCloudSurfLastNum);
    }

    // call when cornerPointsSharp / surfPointsFlat were (re)filled on the host by the caller
    void featuresChanged() { features_pushed_ = false; features_on_device_ = false; }
    void findCorrespondingSurfFeatures(int iterCount) { staged_which_ = 0; staged_iter_ = iterCount; push_features(); }     // FA:1155
    void findCorrespondingCornerFeatures(int iterCount) { staged_which_ = 1; staged_iter_ = iterCount; push_features(); }   // FA:1044
    // return value as in the reference: FALSE when converged (C8)
    bool calculateTransformationSurf(int iterCount) { return step(0, iterCount); }       // FA:1270
    bool calculateTransformationCorner(int iterCount) { return step(1, iterCount); }     // FA:1379

    void updateTransformation()                                                          // FA:1666
    {
        if (laserCloudCornerLastNum < 10 || laserCloudSurfLastNum < 100) return;
        features_pushed_ = false;                            // one call per sweep: the members are (re)sent
        push_features();
        last_status = llb_odom_optimize(ctx_, transformCur, &stats_surf, &stats_corner);
        if (last_status != LLB_OK) return;
        int d = 0;
        if (llb_odom_get_degeneracy(ctx_, &d, matP) == LLB_OK) isDegenerate = d != 0;
    }
    llb_ctx *context() { return ctx_; }

private:
    void pull_imu_sweep()                // the members adjustDistortion's IMU branch writes (FA:556-611), from the device pass
    {
        llb_imu_sweep w;
        last_status = llb_features_get_imu(ctx_, &w);
        if (last_status != LLB_OK || !w.valid) return;
        imuRollStart = w.start[0]; imuPitchStart = w.start[1]; imuYawStart = w.start[2];
        imuVeloXStart = w.start[3]; imuVeloYStart = w.start[4]; imuVeloZStart = w.start[5];
        imuShiftXStart = w.start[6]; imuShiftYStart = w.start[7]; imuShiftZStart = w.start[8];
        imuAngularRotationXCur = w.angular_cur[0]; imuAngularRotationYCur = w.angular_cur[1]; imuAngularRotationZCur = w.angular_cur[2];
        imuAngularFromStartX = imuAngularRotationXCur - imuAngularRotationXLast;           // FA:596-602
        imuAngularFromStartY = imuAngularRotationYCur - imuAngularRotationYLast;
        imuAngularFromStartZ = imuAngularRotationZCur - imuAngularRotationZLast;
        imuAngularRotationXLast = imuAngularRotationXCur; imuAngularRotationYLast = imuAngularRotationYCur;
        imuAngularRotationZLast = imuAngularRotationZCur;
        imuRollCur = w.cur[0]; imuPitchCur = w.cur[1]; imuYawCur = w.cur[2];
        if (w.has_velo) { imuVeloFromStartXCur = w.velo_from_start_cur[0]; imuVeloFromStartYCur = w.velo_from_start_cur[1];
                          imuVeloFromStartZCur = w.velo_from_start_cur[2]; }
    }
    void push_features()                 // once per sweep: the step-wise calls (<= 50 per sweep) do not upload again
    {
        if (features_pushed_) return;
        if (features_on_device_) { last_status = llb_features_to_odometry(ctx_); features_on_device_ = false; }
        else last_status = llb_odom_set_features(ctx_, as_llb(*cornerPointsSharp), (int)cornerPointsSharp->size(),
                                                 as_llb(*surfPointsFlat), (int)surfPointsFlat->size());
        features_pushed_ = last_status == LLB_OK;
    }
    bool step(int which, int iterCount)
    {
        int more = 1, n = 0;
        last_status = llb_odom_iterate(ctx_, which, transformCur, iterCount, &more, &n);
        if (last_status != LLB_OK) return false;             // a dead context must end the caller's loop, not spin it
        int m = 0;
        if (llb_odom_get_correspondences(ctx_, nullptr, nullptr, 0, &m) == LLB_OK) {
            laserCloudOri->resize(m); coeffSel->resize(m);
            if (m > 0) llb_odom_get_correspondences(ctx_, as_llb(*laserCloudOri), as_llb(*coeffSel), m, &m);
        }
        int d = 0;
        if (llb_odom_get_degeneracy(ctx_, &d, matP) == LLB_OK) isDegenerate = d != 0;
        return more != 0;
    }
    llb_ctx *ctx_ = nullptr;
    int staged_which_ = 0, staged_iter_ = 0;
    int fe_staged_ = 0, n_scan_ = 0;
    bool features_pushed_ = false;      // cornerPointsSharp / surfPointsFlat of this sweep are in the odometry's buffers
    bool features_on_device_ = false;   // cornerPointsSharp / surfPointsFlat of the last extractFeatures() are still on the device
};

}  // namespace lego_loam_b200
