// ref_mo.cpp — ORACLE tier B (test infrastructure): compiles the UNMODIFIED reference
// /root/reference/LeGO-LOAM/src/mapOptmization.cpp (included below from where it lies, never
// copied) against shim/llref_shim.hpp and exposes its own member functions through a C ABI
// shaped like oracle/llo.h's llo_mapopt_* so the same tests drive both.
// `private` -> `public` lets the harness reach the class members the ROS callbacks would fill;
// `main` is renamed so the node's main() does not clash.
#include "shim/llref_shim.hpp"          // every standard header first, before the keyword hack
#define private public
#define main ref_mo_node_main
#include "mapOptmization.cpp"
#undef main
#undef private

namespace {
void fill(pcl::PointCloud<PointType>::Ptr &c, const llo_point *p, int n)
{
    c->clear();
    c->points.resize(n);
    for (int i = 0; i < n; i++) {
        PointType q;
        q.x = p[i].x; q.y = p[i].y; q.z = p[i].z; q.intensity = p[i].intensity;
        c->points[i] = q;
    }
    c->width = n; c->height = 1;
}
sensor_msgs::PointCloud2ConstPtr to_msg(const llo_point *p, int n, double stamp)
{
    auto m = std::make_shared<sensor_msgs::PointCloud2>();
    m->header.stamp = ros::Time(stamp);
    m->xyzi.resize((size_t)n * 4);
    for (int i = 0; i < n; i++) { m->xyzi[4 * i] = p[i].x; m->xyzi[4 * i + 1] = p[i].y; m->xyzi[4 * i + 2] = p[i].z; m->xyzi[4 * i + 3] = p[i].intensity; }
    return m;
}
int dump(const pcl::PointCloud<PointType>::Ptr &c, llo_point *out, int cap)
{
    int n = (int)c->points.size();
    for (int i = 0; i < n && i < cap; i++) { out[i].x = c->points[i].x; out[i].y = c->points[i].y; out[i].z = c->points[i].z; out[i].intensity = c->points[i].intensity; }
    return n;
}
}  // namespace

extern "C" {
void *ref_mo_create() { return new mapOptimization(); }
void ref_mo_destroy(void *h) { delete (mapOptimization *)h; }

// hand-over of the DS local map (what MO:1057-1064 leaves in the members)
void ref_mo_set_map_ds(void *h, const llo_point *c, int mc, const llo_point *s, int ms)
{
    mapOptimization *m = (mapOptimization *)h;
    fill(m->laserCloudCornerFromMapDS, c, mc); fill(m->laserCloudSurfFromMapDS, s, ms);
    m->laserCloudCornerFromMapDSNum = mc; m->laserCloudSurfFromMapDSNum = ms;
}
// raw local map + the reference's own filter objects, statements of MO:1058-1064
void ref_mo_set_map_raw(void *h, const llo_point *c, int rc, const llo_point *s, int rs)
{
    mapOptimization *m = (mapOptimization *)h;
    fill(m->laserCloudCornerFromMap, c, rc); fill(m->laserCloudSurfFromMap, s, rs);
    m->downSizeFilterCorner.setInputCloud(m->laserCloudCornerFromMap);
    m->downSizeFilterCorner.filter(*m->laserCloudCornerFromMapDS);
    m->laserCloudCornerFromMapDSNum = m->laserCloudCornerFromMapDS->points.size();
    m->downSizeFilterSurf.setInputCloud(m->laserCloudSurfFromMap);
    m->downSizeFilterSurf.filter(*m->laserCloudSurfFromMapDS);
    m->laserCloudSurfFromMapDSNum = m->laserCloudSurfFromMapDS->points.size();
}
// through the reference's own topic handlers MO:608-627
void ref_mo_set_scan(void *h, const llo_point *c, int nc, const llo_point *s, int ns, const llo_point *o, int no)
{
    mapOptimization *m = (mapOptimization *)h;
    m->laserCloudCornerLastHandler(to_msg(c, nc, 0.0));
    m->laserCloudSurfLastHandler(to_msg(s, ns, 0.0));
    m->laserCloudOutlierLastHandler(to_msg(o, no, 0.0));
}
void ref_mo_set_pose(void *h, const float *t) { memcpy(((mapOptimization *)h)->transformTobeMapped, t, 24); }
void ref_mo_get_pose(void *h, float *t) { memcpy(t, ((mapOptimization *)h)->transformTobeMapped, 24); }
void ref_mo_set_transform_sum(void *h, const float *t) { memcpy(((mapOptimization *)h)->transformSum, t, 24); }
void ref_mo_get_bef_aft(void *h, float *b, float *a)
{
    memcpy(b, ((mapOptimization *)h)->transformBefMapped, 24); memcpy(a, ((mapOptimization *)h)->transformAftMapped, 24);
}
void ref_mo_get_degenerate(void *h, int *d, float *P)
{
    mapOptimization *m = (mapOptimization *)h;
    *d = m->isDegenerate ? 1 : 0;
    for (int i = 0; i < 36; i++) P[i] = i < (int)m->matP.d.size() ? m->matP.d[i] : 0.f;
}
void ref_mo_downsampleCurrentScan(void *h) { ((mapOptimization *)h)->downsampleCurrentScan(); }
void ref_mo_build_kdtrees(void *h)
{   // MO:1333-1334
    mapOptimization *m = (mapOptimization *)h;
    m->kdtreeCornerFromMap->setInputCloud(m->laserCloudCornerFromMapDS);
    m->kdtreeSurfFromMap->setInputCloud(m->laserCloudSurfFromMapDS);
}
void ref_mo_clear_correspondences(void *h)
{   // MO:1338-1339
    mapOptimization *m = (mapOptimization *)h;
    m->laserCloudOri->clear(); m->coeffSel->clear();
}
void ref_mo_cornerOptimization(void *h, int it) { ((mapOptimization *)h)->cornerOptimization(it); }
void ref_mo_surfOptimization(void *h, int it) { ((mapOptimization *)h)->surfOptimization(it); }
int ref_mo_LMOptimization(void *h, int it) { return ((mapOptimization *)h)->LMOptimization(it) ? 1 : 0; }
void ref_mo_scan2MapOptimization(void *h) { ((mapOptimization *)h)->scan2MapOptimization(); }
int ref_mo_get_scan_ds(void *h, int which, llo_point *out, int cap)
{
    mapOptimization *m = (mapOptimization *)h;
    return dump(which == 0 ? m->laserCloudCornerLastDS : which == 1 ? m->laserCloudSurfLastDS
              : which == 2 ? m->laserCloudOutlierLastDS : m->laserCloudSurfTotalLastDS, out, cap);
}
int ref_mo_get_map_ds(void *h, int which, llo_point *out, int cap)
{
    mapOptimization *m = (mapOptimization *)h;
    return dump(which == 0 ? m->laserCloudCornerFromMapDS : m->laserCloudSurfFromMapDS, out, cap);
}
int ref_mo_get_correspondences(void *h, llo_point *ori, llo_point *co, int cap)
{
    mapOptimization *m = (mapOptimization *)h;
    dump(m->laserCloudOri, ori, cap);
    return dump(m->coeffSel, co, cap);
}
// the rest of run() (MO:1503-1519) for sequence replays: key-frame store, local-map assembly
void ref_mo_set_odometry(void *h, const float *sum, double stamp)
{
    mapOptimization *m = (mapOptimization *)h;
    memcpy(m->transformSum, sum, 24); m->timeLaserOdometry = stamp;
}
void ref_mo_transformAssociateToMap(void *h) { ((mapOptimization *)h)->transformAssociateToMap(); }
void ref_mo_transformUpdate(void *h) { ((mapOptimization *)h)->transformUpdate(); }
// imuHandler MO:643-652 after the quaternion -> roll / pitch conversion (tf is not part of the path)
void ref_mo_push_imu(void *h, double stamp, double roll, double pitch)
{
    mapOptimization *m = (mapOptimization *)h;
    m->imuPointerLast = (m->imuPointerLast + 1) % imuQueLength;
    m->imuTime[m->imuPointerLast] = stamp;
    m->imuRoll[m->imuPointerLast] = roll;
    m->imuPitch[m->imuPointerLast] = pitch;
}
void ref_mo_get_aft_mapped(void *h, float *t) { memcpy(t, ((mapOptimization *)h)->transformAftMapped, 24); }
int ref_mo_map_ds_sizes(void *h, int *nc, int *ns)
{
    mapOptimization *m = (mapOptimization *)h;
    *nc = m->laserCloudCornerFromMapDSNum; *ns = m->laserCloudSurfFromMapDSNum;
    return 0;
}
void ref_mo_extractSurroundingKeyFrames(void *h) { ((mapOptimization *)h)->extractSurroundingKeyFrames(); }
void ref_mo_saveKeyFramesAndFactor(void *h) { ((mapOptimization *)h)->saveKeyFramesAndFactor(); }
void ref_mo_correctPoses(void *h) { ((mapOptimization *)h)->correctPoses(); }
void ref_mo_clearCloud(void *h) { ((mapOptimization *)h)->clearCloud(); }
int ref_mo_num_keyframes(void *h) { return (int)((mapOptimization *)h)->cloudKeyPoses3D->points.size(); }
// key-frame bookkeeping the reference keeps on the host (inputs of the device key-frame store's parity test)
int ref_mo_get_surrounding_ids(void *h, int *out, int cap)
{   // surroundingExistingKeyPosesID in order (MO:1033-1049; loopClosureEnableFlag is false, UT:104)
    mapOptimization *m = (mapOptimization *)h;
    int n = (int)m->surroundingExistingKeyPosesID.size();
    for (int i = 0; i < n && i < cap; i++) out[i] = m->surroundingExistingKeyPosesID[i];
    return n;
}
void ref_mo_get_keypose6d(void *h, int idx, float *out6)
{   // cloudKeyPoses6D[idx]: roll, pitch, yaw, x, y, z
    mapOptimization *m = (mapOptimization *)h;
    const PointTypePose &p = m->cloudKeyPoses6D->points[idx];
    out6[0] = p.roll; out6[1] = p.pitch; out6[2] = p.yaw; out6[3] = p.x; out6[4] = p.y; out6[5] = p.z;
}
int ref_mo_get_keyframe_cloud(void *h, int idx, int which, llo_point *out, int cap)
{
    mapOptimization *m = (mapOptimization *)h;
    return dump(which == 0 ? m->cornerCloudKeyFrames[idx] : which == 1 ? m->surfCloudKeyFrames[idx]
                                                                       : m->outlierCloudKeyFrames[idx], out, cap);
}
int ref_mo_get_map_raw(void *h, int which, llo_point *out, int cap)
{
    mapOptimization *m = (mapOptimization *)h;
    return dump(which == 0 ? m->laserCloudCornerFromMap : m->laserCloudSurfFromMap, out, cap);
}
// ---- loop closure + global map: the reference's own detectLoopClosure / performLoopClosure /
// publishGlobalMap; pcl::IterativeClosestPoint is the restatement in oracle/llo_loop.c behind the shim class
void ref_mo_set_robot_pos(void *h, float x, float y, float z)
{
    mapOptimization *m = (mapOptimization *)h;
    m->currentRobotPosPoint.x = x; m->currentRobotPosPoint.y = y; m->currentRobotPosPoint.z = z;
}
void ref_mo_set_time(void *h, double stamp) { ((mapOptimization *)h)->timeLaserOdometry = stamp; }
int ref_mo_detectLoopClosure(void *h) { return ((mapOptimization *)h)->detectLoopClosure() ? 1 : 0; }
void ref_mo_loop_ids(void *h, int *closest, int *latest)
{
    mapOptimization *m = (mapOptimization *)h;
    *closest = m->closestHistoryFrameID; *latest = m->latestFrameIDLoopCloure;
}
// which: 0 latestSurfKeyFrameCloud, 1 nearHistorySurfKeyFrameCloud, 2 nearHistorySurfKeyFrameCloudDS, 3 globalMapKeyFramesDS,
// 4 globalMapKeyPosesDS
int ref_mo_get_loop_cloud(void *h, int which, llo_point *out, int cap)
{
    mapOptimization *m = (mapOptimization *)h;
    return dump(which == 0 ? m->latestSurfKeyFrameCloud : which == 1 ? m->nearHistorySurfKeyFrameCloud
              : which == 2 ? m->nearHistorySurfKeyFrameCloudDS : which == 3 ? m->globalMapKeyFramesDS : m->globalMapKeyPosesDS, out, cap);
}
// performLoopClosure MO:875-945 (returns aLoopIsClosed); the ICP it ran is read back with ref_mo_icp_last
int ref_mo_performLoopClosure(void *h)
{
    mapOptimization *m = (mapOptimization *)h;
    m->aLoopIsClosed = false;
    m->performLoopClosure();
    return m->aLoopIsClosed ? 1 : 0;
}
int ref_mo_icp_last(float *T16, int *converged, int *iterations, int *state, double *fitness)
{
    const pcl::LlrefIcpRecord &r = pcl::llref_icp_last();
    memcpy(T16, r.T, 64); *converged = r.converged; *iterations = r.iterations; *state = r.state; *fitness = r.fitness;
    return r.calls;
}
// publishGlobalMap MO:758-800 with one subscriber; globalMapKeyPosesDS is cleared by the function, so its selection is
// returned through ids (thisKeyInd of MO:781, in order)
int ref_mo_publishGlobalMap(void *h, int *ids, int cap)
{
    mapOptimization *m = (mapOptimization *)h;
    // the key-frame selection of MO:766-778 repeated on copies (the function clears its own), same objects, same order
    int n = 0;
    if (!m->cloudKeyPoses3D->points.empty()) {
        std::vector<int> ind; std::vector<float> dis;
        m->kdtreeGlobalMap->setInputCloud(m->cloudKeyPoses3D);
        m->kdtreeGlobalMap->radiusSearch(m->currentRobotPosPoint, globalMapVisualizationSearchRadius, ind, dis, 0);
        pcl::PointCloud<PointType>::Ptr kp(new pcl::PointCloud<PointType>()), kpds(new pcl::PointCloud<PointType>());
        for (size_t i = 0; i < ind.size(); ++i) kp->points.push_back(m->cloudKeyPoses3D->points[ind[i]]);
        m->downSizeFilterGlobalMapKeyPoses.setInputCloud(kp);
        m->downSizeFilterGlobalMapKeyPoses.filter(*kpds);
        for (size_t i = 0; i < kpds->points.size(); ++i) { if (n < cap) ids[n] = (int)kpds->points[i].intensity; n++; }
    }
    ros::llref_num_subscribers() = 1;
    m->publishGlobalMap();
    ros::llref_num_subscribers() = 0;
    return n;
}
}  // extern "C"
