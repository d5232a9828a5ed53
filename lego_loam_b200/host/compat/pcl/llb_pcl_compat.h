// llb_pcl_compat.h — layout-compatible stand-ins for pcl::PointXYZI and pcl::PointCloud<T>, used
// by the adapter classes ONLY when real PCL headers are not available (define LLB_USE_REAL_PCL and
// include <pcl/point_cloud.h>, <pcl/point_types.h> before the adapters to build against real PCL).
// Layout per SURVEY.md A.5: PointXYZI is 32 bytes, 16-byte aligned: {x,y,z,1.0f}{intensity,pad x3}.
#pragma once
#ifndef LLB_USE_REAL_PCL
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace pcl {

struct PCLHeader { std::uint32_t seq = 0; std::uint64_t stamp = 0; std::string frame_id; };

struct alignas(16) PointXYZI {
    union { float data[4]; struct { float x, y, z; }; };
    union { struct { float intensity; }; float data_c[4]; };
    PointXYZI() { x = y = z = 0.f; data[3] = 1.f; intensity = 0.f; data_c[1] = data_c[2] = data_c[3] = 0.f; }
};
static_assert(sizeof(PointXYZI) == 32 && alignof(PointXYZI) == 16, "pcl::PointXYZI layout");

template <typename PointT>
class PointCloud {
public:
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
    PCLHeader header;
    std::vector<PointT> points;
    std::uint32_t width = 0, height = 0;
    bool is_dense = true;

    void push_back(const PointT &p) { points.push_back(p); width = (std::uint32_t)points.size(); height = 1; }
    void clear() { points.clear(); width = 0; height = 0; }
    void resize(std::size_t n) { points.resize(n); width = (std::uint32_t)n; height = 1; }
    std::size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    PointCloud &operator+=(const PointCloud &o)
    {
        points.insert(points.end(), o.points.begin(), o.points.end());
        width = (std::uint32_t)points.size(); height = 1;
        is_dense = is_dense && o.is_dense;
        return *this;
    }
};

}  // namespace pcl
#endif  // LLB_USE_REAL_PCL
