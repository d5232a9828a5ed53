// ref_fa.cpp — ORACLE tier B (test infrastructure): compiles the UNMODIFIED reference
// /root/reference/LeGO-LOAM/src/featureAssociation.cpp against shim/llref_shim.hpp (see ref_mo.cpp).
#include "shim/llref_shim.hpp"
#define private public
#define main ref_fa_node_main
#include "featureAssociation.cpp"
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
int dump(const pcl::PointCloud<PointType>::Ptr &c, llo_point *out, int cap)
{
    int n = (int)c->points.size();
    for (int i = 0; i < n && i < cap; i++) { out[i].x = c->points[i].x; out[i].y = c->points[i].y; out[i].z = c->points[i].z; out[i].intensity = c->points[i].intensity; }
    return n;
}
}  // namespace

extern "C" {
void *ref_fa_create()
{
    FeatureAssociation *f = new FeatureAssociation();
    // the reference reads cloudCurvature / cloudNeighborPicked / cloudLabel at indices its calculateSmoothness never
    // writes (index 0..4, FA:624 vs FA:688): `new T[]` leaves them uninitialised, the harness defines them as zero
    memset(f->cloudCurvature, 0, sizeof(float) * N_SCAN * Horizon_SCAN);
    // ... and it writes cloudNeighborPicked[-1..-5] when the stale smoothness record {0, 0} is picked (FA:766-773 with
    // ind = 0): give the array 16 entries of slack in front so that this stays inside the harness's own memory
    delete[] f->cloudNeighborPicked;
    f->cloudNeighborPicked = (new int[N_SCAN * Horizon_SCAN + 16]) + 16;
    memset(f->cloudNeighborPicked - 16, 0, sizeof(int) * (N_SCAN * Horizon_SCAN + 16));
    memset(f->cloudLabel, 0, sizeof(int) * N_SCAN * Horizon_SCAN);
    return f;
}
void ref_fa_destroy(void *h) { delete (FeatureAssociation *)h; }
// laserCloudCornerLast / SurfLast + the statements of FA:1615-1619 (force) or FA:1782-1788
void ref_fa_set_last(void *h, const llo_point *c, int nc, const llo_point *s, int ns, int force)
{
    FeatureAssociation *f = (FeatureAssociation *)h;
    fill(f->laserCloudCornerLast, c, nc); fill(f->laserCloudSurfLast, s, ns);
    f->laserCloudCornerLastNum = nc; f->laserCloudSurfLastNum = ns;
    if (force || (nc > 10 && ns > 100)) {
        f->kdtreeCornerLast->setInputCloud(f->laserCloudCornerLast);
        f->kdtreeSurfLast->setInputCloud(f->laserCloudSurfLast);
    }
}
void ref_fa_set_features(void *h, const llo_point *sharp, int nsharp, const llo_point *flat, int nflat)
{
    FeatureAssociation *f = (FeatureAssociation *)h;
    fill(f->cornerPointsSharp, sharp, nsharp); fill(f->surfPointsFlat, flat, nflat);
}
void ref_fa_set_transform(void *h, const float *t) { memcpy(((FeatureAssociation *)h)->transformCur, t, 24); }
void ref_fa_get_transform(void *h, float *t) { memcpy(t, ((FeatureAssociation *)h)->transformCur, 24); }
void ref_fa_get_degenerate(void *h, int *d, float *P)
{
    FeatureAssociation *f = (FeatureAssociation *)h;
    *d = f->isDegenerate ? 1 : 0;
    for (int i = 0; i < 9; i++) P[i] = i < (int)f->matP.d.size() ? f->matP.d[i] : 0.f;
}
void ref_fa_clear_correspondences(void *h)
{   // FA:1672-1673
    FeatureAssociation *f = (FeatureAssociation *)h;
    f->laserCloudOri->clear(); f->coeffSel->clear();
}
void ref_fa_findCorrespondingCornerFeatures(void *h, int it) { ((FeatureAssociation *)h)->findCorrespondingCornerFeatures(it); }
void ref_fa_findCorrespondingSurfFeatures(void *h, int it) { ((FeatureAssociation *)h)->findCorrespondingSurfFeatures(it); }
int ref_fa_calculateTransformationSurf(void *h, int it) { return ((FeatureAssociation *)h)->calculateTransformationSurf(it) ? 1 : 0; }
int ref_fa_calculateTransformationCorner(void *h, int it) { return ((FeatureAssociation *)h)->calculateTransformationCorner(it) ? 1 : 0; }
void ref_fa_updateTransformation(void *h) { ((FeatureAssociation *)h)->updateTransformation(); }
int ref_fa_get_correspondences(void *h, llo_point *ori, llo_point *co, int cap)
{
    FeatureAssociation *f = (FeatureAssociation *)h;
    dump(f->laserCloudOri, ori, cap);
    return dump(f->coeffSel, co, cap);
}
int ref_fa_get_search_ind(void *h, int which, float *i1, float *i2, float *i3, int cap)
{
    FeatureAssociation *f = (FeatureAssociation *)h;
    int n = which == 0 ? (int)f->cornerPointsSharp->points.size() : (int)f->surfPointsFlat->points.size();
    for (int i = 0; i < n && i < cap; i++) {
        if (which == 0) { i1[i] = f->pointSearchCornerInd1[i]; i2[i] = f->pointSearchCornerInd2[i]; }
        else { i1[i] = f->pointSearchSurfInd1[i]; i2[i] = f->pointSearchSurfInd2[i]; if (i3) i3[i] = f->pointSearchSurfInd3[i]; }
    }
    return n;
}
// ---- feature extraction: segmentedCloud + segInfo in (as laserCloudHandler / laserCloudInfoHandler
// leave them, FA:461-489), then the statements of runFeatureAssociation FA:1827-1833
int ref_fa_n_scan() { return N_SCAN; }
int ref_fa_horizon_scan() { return Horizon_SCAN; }
void ref_fa_set_segmented(void *h, const llo_point *pts, int n, const int *start_ring, const int *end_ring,
                          float start_ori, float end_ori, float ori_diff,
                          const unsigned char *ground, const unsigned *col, const float *range)
{
    FeatureAssociation *f = (FeatureAssociation *)h;
    fill(f->segmentedCloud, pts, n);
    cloud_msgs::cloud_info &s = f->segInfo;
    s.startRingIndex.assign(start_ring, start_ring + N_SCAN);
    s.endRingIndex.assign(end_ring, end_ring + N_SCAN);
    s.startOrientation = start_ori; s.endOrientation = end_ori; s.orientationDiff = ori_diff;
    s.segmentedCloudGroundFlag.assign(N_SCAN * Horizon_SCAN, 0);     // sized as IP:141-143 sizes them
    s.segmentedCloudColInd.assign(N_SCAN * Horizon_SCAN, 0);
    s.segmentedCloudRange.assign(N_SCAN * Horizon_SCAN, 0);
    for (int i = 0; i < n; i++) { s.segmentedCloudGroundFlag[i] = ground[i]; s.segmentedCloudColInd[i] = col[i]; s.segmentedCloudRange[i] = range[i]; }
}
void ref_fa_extract_features(void *h)
{
    FeatureAssociation *f = (FeatureAssociation *)h;
    f->adjustDistortion(); f->calculateSmoothness(); f->markOccludedPoints(); f->extractFeatures();
}
int ref_fa_get_cloud(void *h, int which, llo_point *out, int cap)
{
    FeatureAssociation *f = (FeatureAssociation *)h;
    switch (which) {
    case 0: return dump(f->cornerPointsSharp, out, cap);
    case 1: return dump(f->cornerPointsLessSharp, out, cap);
    case 2: return dump(f->surfPointsFlat, out, cap);
    case 3: return dump(f->surfPointsLessFlat, out, cap);
    case 5: return dump(f->laserCloudCornerLast, out, cap);
    case 6: return dump(f->laserCloudSurfLast, out, cap);
    default: return dump(f->segmentedCloud, out, cap);
    }
}
// publishCloudsLast FA:1759-1815: TransformToEnd on cornerPointsLessSharp / surfPointsLessFlat, swap into
// laserCloudCornerLast / laserCloudSurfLast, kd-tree rebuild
void ref_fa_publishCloudsLast(void *h) { ((FeatureAssociation *)h)->publishCloudsLast(); }
void ref_fa_get_point_state(void *h, int n, float *curv, int *picked, int *label)
{
    FeatureAssociation *f = (FeatureAssociation *)h;
    for (int i = 0; i < n; i++) { curv[i] = f->cloudCurvature[i]; picked[i] = f->cloudNeighborPicked[i]; label[i] = f->cloudLabel[i]; }
}
// libstdc++'s own sort of (value, ind) records by value, as FA:699 calls it; depth_limit >= 0 drives the internal
// introsort loop with that limit so that tests reach its heap-sort branch
void ref_fa_std_sort(float *value, unsigned long long *ind, int n, int depth_limit)
{
    std::vector<smoothness_t> v(n);
    for (int i = 0; i < n; i++) { v[i].value = value[i]; v[i].ind = ind[i]; }
    if (depth_limit < 0) std::sort(v.begin(), v.end(), by_value());
    else if (n > 0) {
        std::__introsort_loop(v.begin(), v.end(), (long)depth_limit, __gnu_cxx::__ops::__iter_comp_iter(by_value()));
        std::__final_insertion_sort(v.begin(), v.end(), __gnu_cxx::__ops::__iter_comp_iter(by_value()));
    }
    for (int i = 0; i < n; i++) { value[i] = v[i].value; ind[i] = v[i].ind; }
}

// the per-sweep bookkeeping around the matcher, for sequence replays (FA:1639-1725, FA:1759-1788)
// ---- IMU (FA:417-448): imuHandler after the quaternion -> roll / pitch / yaw conversion (tf is not part of the path):
// gravity compensation, ring-buffer entry, AccumulateIMUShiftAndRotation (the reference's own member function)
void ref_fa_push_imu(void *h, double stamp, double roll, double pitch, double yaw, const double lin_acc[3], const double ang_vel[3])
{
    FeatureAssociation *f = (FeatureAssociation *)h;
    float accX = lin_acc[1] - sin(roll) * cos(pitch) * 9.81;
    float accY = lin_acc[2] - cos(roll) * cos(pitch) * 9.81;
    float accZ = lin_acc[0] + sin(pitch) * 9.81;
    f->imuPointerLast = (f->imuPointerLast + 1) % imuQueLength;
    f->imuTime[f->imuPointerLast] = stamp;
    f->imuRoll[f->imuPointerLast] = roll;
    f->imuPitch[f->imuPointerLast] = pitch;
    f->imuYaw[f->imuPointerLast] = yaw;
    f->imuAccX[f->imuPointerLast] = accX;
    f->imuAccY[f->imuPointerLast] = accY;
    f->imuAccZ[f->imuPointerLast] = accZ;
    f->imuAngularVeloX[f->imuPointerLast] = ang_vel[0];
    f->imuAngularVeloY[f->imuPointerLast] = ang_vel[1];
    f->imuAngularVeloZ[f->imuPointerLast] = ang_vel[2];
    f->AccumulateIMUShiftAndRotation();
}
void ref_fa_set_time_scan_cur(void *h, double t) { ((FeatureAssociation *)h)->timeScanCur = t; }
void ref_fa_updateInitialGuess(void *h) { ((FeatureAssociation *)h)->updateInitialGuess(); }
// the IMU state a sweep leaves behind: {imuRoll/Pitch/YawStart, imuVeloX/Y/ZStart, imuShiftX/Y/ZStart, imuRoll/Pitch/YawCur,
// imuVeloFromStartX/Y/ZCur, imuAngularFromStartX/Y/Z, imuAngularRotationX/Y/ZCur, imuRoll/Pitch/YawLast} (24 floats) and the
// ring-buffer entry `idx`: {time (as double in out_d), roll, pitch, yaw, velo xyz, shift xyz, angular rotation xyz}
void ref_fa_get_imu_state(void *h, float *o)
{
    FeatureAssociation *f = (FeatureAssociation *)h;
    const float v[24] = { f->imuRollStart, f->imuPitchStart, f->imuYawStart, f->imuVeloXStart, f->imuVeloYStart, f->imuVeloZStart,
                          f->imuShiftXStart, f->imuShiftYStart, f->imuShiftZStart, f->imuRollCur, f->imuPitchCur, f->imuYawCur,
                          f->imuVeloFromStartXCur, f->imuVeloFromStartYCur, f->imuVeloFromStartZCur,
                          f->imuAngularFromStartX, f->imuAngularFromStartY, f->imuAngularFromStartZ,
                          f->imuAngularRotationXCur, f->imuAngularRotationYCur, f->imuAngularRotationZCur,
                          f->imuRollLast, f->imuPitchLast, f->imuYawLast };
    memcpy(o, v, sizeof v);
}
int ref_fa_get_imu_entry(void *h, int idx, double *t, float *o)
{
    FeatureAssociation *f = (FeatureAssociation *)h;
    if (idx < 0) idx = f->imuPointerLast;
    if (idx < 0) return -1;
    *t = f->imuTime[idx];
    const float v[12] = { f->imuRoll[idx], f->imuPitch[idx], f->imuYaw[idx], f->imuVeloX[idx], f->imuVeloY[idx], f->imuVeloZ[idx],
                          f->imuShiftX[idx], f->imuShiftY[idx], f->imuShiftZ[idx],
                          f->imuAngularRotationX[idx], f->imuAngularRotationY[idx], f->imuAngularRotationZ[idx] };
    memcpy(o, v, sizeof v);
    return idx;
}
void ref_fa_integrateTransformation(void *h) { ((FeatureAssociation *)h)->integrateTransformation(); }
void ref_fa_get_transform_sum(void *h, float *t) { memcpy(t, ((FeatureAssociation *)h)->transformSum, 24); }
}  // extern "C"
