// Drives the C++ adapter classes (reference member-function names) end to end on a GPU:
// reads clouds + expected results written by tests/test_gpu_adapter.py, runs
// downsampleCurrentScan() + scan2MapOptimization() and updateTransformation(), prints poses.
#include "../lego_loam_b200/host/lego_loam_b200.hpp"
#include <cstdio>
#include <cstdlib>
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

    FeatureAssociation FA;
    read_cloud(f, *FA.laserCloudCornerLast); read_cloud(f, *FA.laserCloudSurfLast);
    read_cloud(f, *FA.cornerPointsSharp); read_cloud(f, *FA.surfPointsFlat);
    FA.setLastClouds();
    FA.updateTransformation();                  // FA:1853
    printf("FA %d %d", FA.stats_surf.iterations, FA.stats_corner.iterations);
    for (int i = 0; i < 6; i++) printf(" %.9g", FA.transformCur[i]);
    printf("\n");
    return 0;
}
