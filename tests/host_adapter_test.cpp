// Drives the C++ adapter classes (reference member-function names) end to end on a GPU:
// reads clouds + expected results written by tests/test_gpu_adapter.py, runs
// downsampleCurrentScan() + scan2MapOptimization() and updateTransformation(), prints poses.
#include "../lego_loam_b200/host/lego_loam_b200.hpp"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

using namespace lego_loam_b200;

static bool read_cloud(std::ifstream &f, Cloud &c)
{
    int n = 0;
    f.read((char *)&n, 4);
    std::vector<float> buf((size_t)n * 4);
    f.read((char *)buf.data(), sizeof(float) * buf.size());
    c.clear();
    for (int i = 0; i < n; i++) {
        PointType p;
        p.x = buf[4 * i]; p.y = buf[4 * i + 1]; p.z = buf[4 * i + 2]; p.intensity = buf[4 * i + 3];
        c.push_back(p);
    }
    return (bool)f;
}

static void read_sweep(std::ifstream &f, FeatureAssociation &FE, int n_scan)
{
    read_cloud(f, *FE.segmentedCloud);
    const size_t n = FE.segmentedCloud->size();
    FE.segInfo.startRingIndex.resize(n_scan); FE.segInfo.endRingIndex.resize(n_scan);
    f.read((char *)FE.segInfo.startRingIndex.data(), 4 * n_scan); f.read((char *)FE.segInfo.endRingIndex.data(), 4 * n_scan);
    f.read((char *)&FE.segInfo.startOrientation, 4); f.read((char *)&FE.segInfo.endOrientation, 4); f.read((char *)&FE.segInfo.orientationDiff, 4);
    FE.segInfo.segmentedCloudGroundFlag.resize(n); FE.segInfo.segmentedCloudColInd.resize(n); FE.segInfo.segmentedCloudRange.resize(n);
    f.read((char *)FE.segInfo.segmentedCloudGroundFlag.data(), n); f.read((char *)FE.segInfo.segmentedCloudColInd.data(), 4 * n);
    f.read((char *)FE.segInfo.segmentedCloudRange.data(), 4 * n);
}

static unsigned fnv_cloud(const Cloud &c, bool with_intensity)
{
    unsigned h = 2166136261u;                          // FNV-1a over the x, y, z (, intensity) words
    for (const PointType &p : c.points) {
        const float v[4] = { p.x, p.y, p.z, p.intensity };
        for (int a = 0; a < (with_intensity ? 4 : 3); a++) { unsigned w; memcpy(&w, &v[a], 4); h = (h ^ w) * 16777619u; }
    }
    return h;
}

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    std::ifstream f(argv[1], std::ios::binary);
    mapOptimization MO;
    read_cloud(f, *MO.laserCloudCornerFromMap); read_cloud(f, *MO.laserCloudSurfFromMap);
    read_cloud(f, *MO.laserCloudCornerLast); read_cloud(f, *MO.laserCloudSurfLast); read_cloud(f, *MO.laserCloudOutlierLast);
    f.read((char *)MO.transformTobeMapped, 24);
    MO.downsampleSurroundingMap();              // MO:1057-1064
    MO.downsampleCurrentScan();                 // MO:1507
    MO.scan2MapOptimization();                  // MO:1509
    printf("MO %d %d %d %d %d %d", MO.laserCloudCornerLastDSNum, MO.laserCloudSurfLastDSNum, MO.laserCloudOutlierLastDSNum,
           MO.laserCloudSurfTotalLastDSNum, MO.last_stats.iterations, (int)MO.isDegenerate);
    for (int i = 0; i < 6; i++) printf(" %.9g", MO.transformTobeMapped[i]);
    printf(" %zu\n", MO.laserCloudSurfTotalLastDS->size());
    // transformUpdate with IMU messages (MO:463-496): K samples {stamp, roll, pitch}, then the odometry stamp and pose
    int n_imu = 0;
    f.read((char *)&n_imu, 4);
    for (int k = 0; k < n_imu; k++) {
        double v[3];
        f.read((char *)v, 24);
        MO.imuHandler(v[0], v[1], v[2]);
    }
    f.read((char *)&MO.timeLaserOdometry, 8);
    f.read((char *)MO.transformSum, 24);
    MO.transformUpdate();
    printf("TU");
    for (int i = 0; i < 6; i++) printf(" %.9g", MO.transformBefMapped[i]);
    for (int i = 0; i < 6; i++) printf(" %.9g", MO.transformAftMapped[i]);
    printf("\n");

    FeatureAssociation FA;
    read_cloud(f, *FA.laserCloudCornerLast); read_cloud(f, *FA.laserCloudSurfLast);
    read_cloud(f, *FA.cornerPointsSharp); read_cloud(f, *FA.surfPointsFlat);
    FA.setLastClouds();
    FA.updateTransformation();                  // FA:1853
    printf("FA %d %d", FA.stats_surf.iterations, FA.stats_corner.iterations);
    for (int i = 0; i < 6; i++) printf(" %.9g", FA.transformCur[i]);
    printf("\n");
    // feature extraction: segmentedCloud + segInfo -> the four feature clouds (FA:1827-1833)
    int n_scan = 0, horizon = 0;
    f.read((char *)&n_scan, 4); f.read((char *)&horizon, 4);
    if (!f) return 0;
    FeatureAssociation FE;
    FE.initFeatureExtraction(n_scan, horizon);
    read_sweep(f, FE, n_scan);
    FE.adjustDistortion(); FE.calculateSmoothness(); FE.markOccludedPoints(); FE.extractFeatures();
    const Cloud *out[4] = { FE.cornerPointsSharp.get(), FE.cornerPointsLessSharp.get(), FE.surfPointsFlat.get(), FE.surfPointsLessFlat.get() };
    printf("FE %d", FE.last_status);
    for (int k = 0; k < 4; k++) {
        unsigned h = 2166136261u;                      // FNV-1a over the x, y, z words
        for (const PointType &p : out[k]->points) {
            const float v[3] = { p.x, p.y, p.z };
            for (int a = 0; a < 3; a++) { unsigned w; memcpy(&w, &v[a], 4); h = (h ^ w) * 16777619u; }
        }
        printf(" %zu %u", out[k]->size(), h);
    }
    printf("\n");
    // IMU branches (FA:417-448 imuHandler + AccumulateIMUShiftAndRotation on the host, adjustDistortion's IMU branch and
    // TransformToEnd's IMU terms on the device, updateInitialGuess): two sweeps, each preceded by a block of messages
    int n_sweeps = 0;
    f.read((char *)&n_sweeps, 4);
    if (!f) return 0;
    FeatureAssociation FI;
    FI.initFeatureExtraction(n_scan, horizon);
    for (int sweep = 0; sweep < n_sweeps; sweep++) {
        int n_msg = 0;
        f.read((char *)&n_msg, 4);
        for (int k = 0; k < n_msg; k++) {
            double v[10];
            f.read((char *)v, 80);
            FI.imuHandler(v[0], v[1], v[2], v[3], v + 4, v + 7);
        }
        double stamp = 0;
        f.read((char *)&stamp, 8);
        FI.laserCloudHandlerStamp(stamp);
        read_sweep(f, FI, n_scan);
        FI.adjustDistortion(); FI.calculateSmoothness(); FI.markOccludedPoints(); FI.extractFeatures();
        f.read((char *)FI.transformCur, 24);
        FI.updateInitialGuess();
        const float st[24] = { FI.imuRollStart, FI.imuPitchStart, FI.imuYawStart, FI.imuVeloXStart, FI.imuVeloYStart, FI.imuVeloZStart,
                               FI.imuShiftXStart, FI.imuShiftYStart, FI.imuShiftZStart, FI.imuRollCur, FI.imuPitchCur, FI.imuYawCur,
                               FI.imuVeloFromStartXCur, FI.imuVeloFromStartYCur, FI.imuVeloFromStartZCur,
                               FI.imuAngularFromStartX, FI.imuAngularFromStartY, FI.imuAngularFromStartZ,
                               FI.imuAngularRotationXCur, FI.imuAngularRotationYCur, FI.imuAngularRotationZCur,
                               FI.imuRollLast, FI.imuPitchLast, FI.imuYawLast };
        printf("IM %d", FI.last_status);
        for (int k = 0; k < 24; k++) { unsigned w; memcpy(&w, &st[k], 4); printf(" %u", w); }
        for (int k = 0; k < 6; k++) { unsigned w; memcpy(&w, &FI.transformCur[k], 4); printf(" %u", w); }
        printf(" %zu %u", FI.segmentedCloud->size(), fnv_cloud(*FI.segmentedCloud, true));
        FI.publishCloudsLast();
        printf(" %zu %u %zu %u\n", FI.laserCloudCornerLast->size(), fnv_cloud(*FI.laserCloudCornerLast, true),
               FI.laserCloudSurfLast->size(), fnv_cloud(*FI.laserCloudSurfLast, true));
    }
    // loop closure: the ICP of performLoopClosure MO:892-904 on the adapter's members
    read_cloud(f, *MO.latestSurfKeyFrameCloud); read_cloud(f, *MO.nearHistorySurfKeyFrameCloudDS);
    const bool closed = MO.performLoopClosureICP(true);
    printf("LC %d %d %d %d %.17g", MO.last_status, (int)closed, MO.last_icp.iterations, MO.last_icp.convergence_state, MO.last_icp.fitness_score);
    for (int k = 0; k < 16; k++) printf(" %.9g", MO.last_icp.T[k]);
    printf("\n");
    // the device key-frame store behind the adapter: one key-frame, loop clouds from it (compiles the member template)
    MO.fetch_downsampled_clouds = false;
    MO.downsampleCurrentScan();
    const int kf = MO.saveKeyFrameClouds();
    const bool lc_ok = MO.detectLoopClosureClouds(kf, kf, [](int, float out[6]) { for (int k = 0; k < 6; k++) out[k] = 0.f; }, 25, true);
    printf("LK %d %d %zu %zu\n", kf, (int)lc_ok, MO.latestSurfKeyFrameCloud->size(), MO.nearHistorySurfKeyFrameCloudDS->size());
    return 0;
}
