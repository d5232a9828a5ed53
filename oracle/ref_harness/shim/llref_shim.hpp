// llref_shim.hpp — ORACLE tier B (test infrastructure): the minimum of ROS / PCL / OpenCV / tf /
// GTSAM API that /root/reference/LeGO-LOAM/src/{mapOptmization,featureAssociation}.cpp use, so
// that those files compile UNMODIFIED, from where they lie, into oracle/_ref/*.so.  None of the
// real libraries exists offline.  The numerical primitives forward to the
// plain-C restatements of oracle/llo.h (cv::solve/eigen/inv/gemm are pinned bit-for-bit to the
// OpenCV 4.13 wheel; pcl::VoxelGrid / pcl::KdTreeFLANN follow SURVEY Appendix A); everything
// else (control flow, thresholds, Jacobians, pose bookkeeping) then comes from the reference
// verbatim.  This header is never included by the product.
#pragma once

#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <deque>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

extern "C" {
#include "llo.h"
}

// ============================================================ Eigen (only what the sources name)
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_ALIGN16 __attribute__((aligned(16)))
namespace Eigen {
template <typename T> using aligned_allocator = std::allocator<T>;
struct Matrix4f {
    float m[16];
    Matrix4f() { for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.f : 0.f; }
    float &operator()(int r, int c) { return m[r * 4 + c]; }
    float operator()(int r, int c) const { return m[r * 4 + c]; }
    static Matrix4f Identity() { return Matrix4f(); }
};
struct Affine3f {
    Matrix4f mat;
    Affine3f() {}
    Affine3f(const Matrix4f &a) : mat(a) {}
    Affine3f &operator=(const Matrix4f &a) { mat = a; return *this; }
    float operator()(int r, int c) const { return mat(r, c); }
    float &operator()(int r, int c) { return mat(r, c); }
    const Matrix4f &matrix() const { return mat; }
    Affine3f operator*(const Affine3f &o) const
    {
        Affine3f r;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) {
                float s = 0;
                for (int k = 0; k < 4; k++) s += mat(i, k) * o.mat(k, j);
                r.mat(i, j) = s;
            }
        return r;
    }
};
}  // namespace Eigen

// ============================================================ ROS core + messages
namespace ros {
struct Time {
    double t = 0;
    Time() {}
    explicit Time(double s) : t(s) {}
    Time &fromSec(double s) { t = s; return *this; }
    double toSec() const { return t; }
    static Time now() { return Time(); }
};
struct Duration { double d; explicit Duration(double s = 0) : d(s) {} void sleep() const {} };
struct Rate { explicit Rate(double) {} void sleep() {} };
inline int &llref_num_subscribers() { static int n = 0; return n; }   // the harness raises it to let publishGlobalMap run
struct Publisher {
    template <typename M> void publish(const M &) const {}
    int getNumSubscribers() const { return llref_num_subscribers(); }
};
struct Subscriber {};
struct NodeHandle {
    NodeHandle() {}
    explicit NodeHandle(const std::string &) {}
    template <typename M> Publisher advertise(const std::string &, int) { return Publisher(); }
    template <typename M, typename C, typename A> Subscriber subscribe(const std::string &, int, void (C::*)(A), C *) { return Subscriber(); }
};
inline void init(int &, char **, const std::string &) {}
inline bool ok() { return false; }
inline void spinOnce() {}
inline void spin() {}
inline void shutdown() {}
}  // namespace ros
#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)

namespace std_msgs {
struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; };
}
namespace geometry_msgs {
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
struct Pose { Point position; Quaternion orientation; };
struct PoseWithCovariance { Pose pose; };
struct Twist { Vector3 linear, angular; };
struct TwistWithCovariance { Twist twist; };
}  // namespace geometry_msgs
namespace sensor_msgs {
struct PointCloud2 {
    std_msgs::Header header;
    std::vector<float> xyzi;                 // shim payload: x y z intensity per point
    std::vector<uint16_t> ring;              // ... and the "ring" channel when the sender has one
    typedef std::shared_ptr<PointCloud2> Ptr;
    typedef std::shared_ptr<const PointCloud2> ConstPtr;
};
typedef std::shared_ptr<const PointCloud2> PointCloud2ConstPtr;
struct Imu {
    std_msgs::Header header;
    geometry_msgs::Quaternion orientation;
    geometry_msgs::Vector3 angular_velocity, linear_acceleration;
    typedef std::shared_ptr<const Imu> ConstPtr;
};
}  // namespace sensor_msgs
namespace nav_msgs {
struct Odometry {
    std_msgs::Header header;
    std::string child_frame_id;
    geometry_msgs::PoseWithCovariance pose;
    geometry_msgs::TwistWithCovariance twist;
    typedef std::shared_ptr<const Odometry> ConstPtr;
};
}  // namespace nav_msgs
namespace cloud_msgs {
struct cloud_info {                          // cloud_msgs/msg/cloud_info.msg:1-11
    std_msgs::Header header;
    std::vector<int32_t> startRingIndex, endRingIndex;
    float startOrientation = 0, endOrientation = 0, orientationDiff = 0;
    std::vector<uint8_t> segmentedCloudGroundFlag;
    std::vector<uint32_t> segmentedCloudColInd;
    std::vector<float> segmentedCloudRange;
};
typedef std::shared_ptr<const cloud_info> cloud_infoConstPtr;
}  // namespace cloud_msgs

// ============================================================ tf (double precision, as tf does)
namespace tf {
struct Vector3 { double x_, y_, z_; Vector3(double x = 0, double y = 0, double z = 0) : x_(x), y_(y), z_(z) {} };
struct Quaternion {
    double x_, y_, z_, w_;
    Quaternion() : x_(0), y_(0), z_(0), w_(1) {}
    Quaternion(double x, double y, double z, double w) : x_(x), y_(y), z_(z), w_(w) {}
    double x() const { return x_; } double y() const { return y_; } double z() const { return z_; } double w() const { return w_; }
};
struct Matrix3x3 {
    double m[3][3];
    explicit Matrix3x3(const Quaternion &q)
    {   // tf::Matrix3x3::setRotation
        double d = q.x_ * q.x_ + q.y_ * q.y_ + q.z_ * q.z_ + q.w_ * q.w_;
        double s = 2.0 / d;
        double xs = q.x_ * s, ys = q.y_ * s, zs = q.z_ * s;
        double wx = q.w_ * xs, wy = q.w_ * ys, wz = q.w_ * zs;
        double xx = q.x_ * xs, xy = q.x_ * ys, xz = q.x_ * zs;
        double yy = q.y_ * ys, yz = q.y_ * zs, zz = q.z_ * zs;
        m[0][0] = 1.0 - (yy + zz); m[0][1] = xy - wz; m[0][2] = xz + wy;
        m[1][0] = xy + wz; m[1][1] = 1.0 - (xx + zz); m[1][2] = yz - wx;
        m[2][0] = xz - wy; m[2][1] = yz + wx; m[2][2] = 1.0 - (xx + yy);
    }
    void getRPY(double &roll, double &pitch, double &yaw) const
    {   // tf::Matrix3x3::getEulerYPR, solution 1
        if (std::fabs(m[2][0]) >= 1) {
            yaw = 0;
            double delta = std::atan2(m[2][1], m[2][2]);
            if (m[2][0] < 0) { pitch = M_PI / 2.0; roll = delta; }
            else { pitch = -M_PI / 2.0; roll = delta; }
        } else {
            pitch = -std::asin(m[2][0]);
            roll = std::atan2(m[2][1] / std::cos(pitch), m[2][2] / std::cos(pitch));
            yaw = std::atan2(m[1][0] / std::cos(pitch), m[0][0] / std::cos(pitch));
        }
    }
};
inline geometry_msgs::Quaternion createQuaternionMsgFromRollPitchYaw(double roll, double pitch, double yaw)
{   // tf::Quaternion::setRPY
    double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
    double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr), sr = std::sin(hr);
    geometry_msgs::Quaternion q;
    q.x = sr * cp * cy - cr * sp * sy;
    q.y = cr * sp * cy + sr * cp * sy;
    q.z = cr * cp * sy - sr * sp * cy;
    q.w = cr * cp * cy + sr * sp * sy;
    return q;
}
inline void quaternionMsgToTF(const geometry_msgs::Quaternion &m, Quaternion &q) { q = Quaternion(m.x, m.y, m.z, m.w); }
struct StampedTransform {
    ros::Time stamp_; std::string frame_id_, child_frame_id_;
    Quaternion rot; Vector3 org;
    void setRotation(const Quaternion &q) { rot = q; }
    void setOrigin(const Vector3 &v) { org = v; }
};
struct TransformBroadcaster { void sendTransform(const StampedTransform &) {} };
}  // namespace tf

// ============================================================ OpenCV (CV_32F small dense only)
#define CV_32F 5
#define CV_8S 1
#define CV_32S 4
namespace cv {
struct Scalar { double v; static Scalar all(double x) { Scalar s; s.v = x; return s; } };
enum { DECOMP_LU = 0, DECOMP_SVD = 1, DECOMP_EIG = 2, DECOMP_CHOLESKY = 3, DECOMP_QR = 4, DECOMP_NORMAL = 16 };
struct Mat {
    int rows = 0, cols = 0, type = CV_32F;
    std::vector<float> d;                    // CV_32F
    std::vector<int32_t> di;                 // CV_32S
    std::vector<int8_t> db;                  // CV_8S (imageProjection's label / ground images)
    Mat() {}
    Mat(int r, int c, int t, const Scalar &s = Scalar::all(0)) : rows(r), cols(c), type(t)
    {
        if (t == CV_32S) di.assign((size_t)r * c, (int32_t)s.v);
        else if (t == CV_8S) db.assign((size_t)r * c, (int8_t)s.v);
        else d.assign((size_t)r * c, (float)s.v);
    }
    float &el(float *, size_t k) { return d[k]; }
    int32_t &el(int32_t *, size_t k) { return di[k]; }
    int8_t &el(int8_t *, size_t k) { return db[k]; }
    const float &el(float *, size_t k) const { return d[k]; }
    template <typename T> T &at(int i, int j) { return el((T *)nullptr, (size_t)i * cols + j); }
    template <typename T> const T &at(int i, int j) const { return el((T *)nullptr, (size_t)i * cols + j); }
    void copyTo(Mat &o) const { o = *this; }
    Mat inv(int = DECOMP_LU) const
    {
        Mat r(rows, cols, CV_32F);
        llo_cv_inv_f32(rows, d.data(), r.d.data());
        return r;
    }
};
inline Mat operator*(const Mat &a, const Mat &b)
{
    Mat r(a.rows, b.cols, CV_32F);
    llo_cv_gemm_f32(a.rows, a.cols, b.cols, a.d.data(), b.d.data(), r.d.data());
    return r;
}
inline void transpose(const Mat &a, Mat &b)
{
    Mat r(a.cols, a.rows, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) r.d[(size_t)j * a.rows + i] = a.d[(size_t)i * a.cols + j];
    b = r;
}
inline bool solve(const Mat &A, const Mat &B, Mat &X, int /*flags = DECOMP_QR*/)
{
    Mat r(A.cols, 1, CV_32F);
    int ok = llo_cv_solve_qr_f32(A.rows, A.cols, A.d.data(), B.d.data(), r.d.data());
    X = r;
    return ok != 0;
}
inline bool eigen(const Mat &A, Mat &E, Mat &V)
{
    Mat tmp = A, e(1, A.rows, CV_32F), v(A.rows, A.rows, CV_32F);
    llo_cv_eigen_f32(A.rows, tmp.d.data(), e.d.data(), v.d.data());
    E = e; V = v;
    return true;
}
}  // namespace cv

// ============================================================ PCL
#define PCL_ADD_POINT4D union { float data[4]; struct { float x; float y; float z; }; };
#define PCL_ADD_INTENSITY union { struct { float intensity; }; float data_c[4]; }
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fseq)

namespace pcl {
struct PCLHeader { uint32_t seq = 0; uint64_t stamp = 0; std::string frame_id; };

struct EIGEN_ALIGN16 PointXYZI {
    PCL_ADD_POINT4D
    PCL_ADD_INTENSITY;
    PointXYZI() { x = y = z = 0.f; data[3] = 1.f; intensity = 0.f; data_c[1] = data_c[2] = data_c[3] = 0.f; }
};
static_assert(sizeof(PointXYZI) == 32, "pcl::PointXYZI is 32 bytes (SURVEY A.5)");

template <typename T>
struct PointCloud {
    typedef std::shared_ptr<PointCloud<T>> Ptr;
    typedef std::shared_ptr<const PointCloud<T>> ConstPtr;
    PCLHeader header;
    std::vector<T> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    void push_back(const T &p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
    void clear() { points.clear(); width = 0; height = 0; }
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void resize(size_t n) { points.resize(n); width = (uint32_t)n; height = 1; }
    PointCloud &operator+=(const PointCloud &o)
    {
        points.insert(points.end(), o.points.begin(), o.points.end());
        width = (uint32_t)points.size(); height = 1;
        is_dense = is_dense && o.is_dense;
        return *this;
    }
    Ptr makeShared() const { return Ptr(new PointCloud<T>(*this)); }
    typename std::vector<T>::iterator begin() { return points.begin(); }
    typename std::vector<T>::iterator end() { return points.end(); }
};

template <typename A, typename B>
inline void copyPointCloud(const PointCloud<A> &in, PointCloud<B> &out)
{
    out.header = in.header; out.width = in.width; out.height = in.height; out.is_dense = in.is_dense;
    out.points.resize(in.points.size());
    for (size_t i = 0; i < in.points.size(); i++) {
        out.points[i].x = in.points[i].x; out.points[i].y = in.points[i].y; out.points[i].z = in.points[i].z;
        out.points[i].intensity = in.points[i].intensity;
    }
}

template <typename T> inline auto shim_set_ring(T &p, uint16_t r, int) -> decltype(p.ring, void()) { p.ring = r; }
template <typename T> inline void shim_set_ring(T &, uint16_t, long) {}
template <typename T>
inline void fromROSMsg(const sensor_msgs::PointCloud2 &msg, PointCloud<T> &cloud)
{
    size_t n = msg.xyzi.size() / 4;
    cloud.points.resize(n);
    for (size_t i = 0; i < n; i++) {
        T p;
        p.x = msg.xyzi[4 * i]; p.y = msg.xyzi[4 * i + 1]; p.z = msg.xyzi[4 * i + 2]; p.intensity = msg.xyzi[4 * i + 3];
        shim_set_ring(p, i < msg.ring.size() ? msg.ring[i] : (uint16_t)0, 0);
        cloud.points[i] = p;
    }
    cloud.width = (uint32_t)n; cloud.height = 1; cloud.is_dense = true;
}
// pcl::removeNaNFromPointCloud (pcl/filters/filter.h): keeps the points whose x, y, z are all finite, in order
template <typename T>
inline void removeNaNFromPointCloud(const PointCloud<T> &in, PointCloud<T> &out, std::vector<int> &index)
{
    std::vector<T> kept; kept.reserve(in.points.size());
    index.clear();
    for (size_t i = 0; i < in.points.size(); i++) {
        const T &p = in.points[i];
        if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
        kept.push_back(p); index.push_back((int)i);
    }
    out.header = in.header;
    out.points.swap(kept);
    out.width = (uint32_t)out.points.size(); out.height = 1; out.is_dense = true;
}
template <typename T>
inline void toROSMsg(const PointCloud<T> &cloud, sensor_msgs::PointCloud2 &msg)
{
    msg.xyzi.resize(cloud.points.size() * 4);
    for (size_t i = 0; i < cloud.points.size(); i++) {
        msg.xyzi[4 * i] = cloud.points[i].x; msg.xyzi[4 * i + 1] = cloud.points[i].y;
        msg.xyzi[4 * i + 2] = cloud.points[i].z; msg.xyzi[4 * i + 3] = cloud.points[i].intensity;
    }
}

inline float rad2deg(float alpha) { return (alpha * 57.29578f); }
inline double rad2deg(double alpha) { return (alpha * 57.29578); }
inline float deg2rad(float alpha) { return (alpha * 0.017453293f); }

// pcl::VoxelGrid<PointXYZI>, defaults + setLeafSize -> llo_voxel_grid
template <typename T>
class VoxelGrid {
public:
    void setLeafSize(float lx, float, float) { leaf_ = lx; }
    void setInputCloud(const typename PointCloud<T>::ConstPtr &c) { in_ = c; }
    void filter(PointCloud<T> &out)
    {
        // the caller's clouds are filtered where they lie (32-byte points), as PCL does: no conversion copies
        static_assert(sizeof(T) == 32, "pcl::PointXYZI layout");
        const size_t n = in_ ? in_->points.size() : 0;
        std::vector<T> tmp(n ? n : 1);                                         // default points: data[3] = 1, padding 0
        int ovf = 0;
        const int m = n ? llo_voxel_grid_pcl(reinterpret_cast<const float *>(in_->points.data()), (int)n, leaf_,
                                             reinterpret_cast<float *>(tmp.data()), &ovf) : 0;
        tmp.resize(m);
        out.points.swap(tmp);
        out.width = (uint32_t)m; out.height = 1; out.is_dense = true;
        if (in_) out.header = in_->header;
    }
private:
    float leaf_ = 0.f;
    typename PointCloud<T>::ConstPtr in_;
};

// pcl::KdTreeFLANN<PointXYZI> -> llo_kdtree (exact, ties by index)
template <typename T>
class KdTreeFLANN {
public:
    typedef std::shared_ptr<KdTreeFLANN<T>> Ptr;
    ~KdTreeFLANN() { llo_kdtree_free(tree_); }
    void setInputCloud(const typename PointCloud<T>::ConstPtr &c)
    {
        cloud_ = c;
        pts_.resize(c->points.size());
        for (size_t i = 0; i < pts_.size(); i++) {
            pts_[i].x = c->points[i].x; pts_[i].y = c->points[i].y; pts_[i].z = c->points[i].z; pts_[i].intensity = 0;
        }
        llo_kdtree_free(tree_);
        tree_ = llo_kdtree_build(pts_.data(), (int)pts_.size());
    }
    int nearestKSearch(const T &p, int k, std::vector<int> &idx, std::vector<float> &d2) const
    {
        if (k > (int)pts_.size()) k = (int)pts_.size();
        idx.resize(k); d2.resize(k);
        if (k == 0) return 0;
        float q[3] = { p.x, p.y, p.z };
        return llo_kdtree_knn(tree_, q, k, idx.data(), d2.data());
    }
    int radiusSearch(const T &p, double radius, std::vector<int> &idx, std::vector<float> &d2, unsigned int max_nn = 0) const
    {
        std::vector<std::pair<float, int>> hit;
        const float r2 = (float)(radius * radius);
        for (size_t i = 0; i < pts_.size(); i++) {
            float dx = p.x - pts_[i].x, dy = p.y - pts_[i].y, dz = p.z - pts_[i].z;
            float d = dx * dx; d += dy * dy; d += dz * dz;
            if (d <= r2) hit.push_back(std::make_pair(d, (int)i));
        }
        std::sort(hit.begin(), hit.end());
        if (max_nn && hit.size() > max_nn) hit.resize(max_nn);
        idx.resize(hit.size()); d2.resize(hit.size());
        for (size_t i = 0; i < hit.size(); i++) { idx[i] = hit[i].second; d2[i] = hit[i].first; }
        return (int)hit.size();
    }
private:
    typename PointCloud<T>::ConstPtr cloud_;
    std::vector<llo_point> pts_;
    llo_kdtree *tree_ = nullptr;
};

// pcl::IterativeClosestPoint as performLoopClosure uses it (MO:892-904): the restatement of PCL 1.8's algorithm in
// oracle/llo_loop.c ("parity unpinned": PCL itself is absent).  The object is a local of performLoopClosure, so the
// harness reads what the last align() saw and produced from llref_icp_last().
struct LlrefIcpRecord {
    std::vector<llo_point> source, target;
    float T[16]; int converged = 0, iterations = 0, state = 0, calls = 0; double fitness = 0;
};
inline LlrefIcpRecord &llref_icp_last() { static LlrefIcpRecord r; return r; }
template <typename A, typename B>
class IterativeClosestPoint {
public:
    void setMaxCorrespondenceDistance(double d) { max_dist_ = d; } void setMaximumIterations(int n) { max_iter_ = n; }
    void setTransformationEpsilon(double e) { teps_ = e; } void setEuclideanFitnessEpsilon(double e) { feps_ = e; }
    void setRANSACIterations(int) {}
    void setInputSource(const typename PointCloud<A>::ConstPtr &c) { src_ = c; }
    void setInputTarget(const typename PointCloud<B>::ConstPtr &c) { tgt_ = c; }
    void align(PointCloud<A> &)
    {
        LlrefIcpRecord &r = llref_icp_last();
        r.source.resize(src_->points.size()); r.target.resize(tgt_->points.size());
        for (size_t i = 0; i < r.source.size(); i++) r.source[i] = llo_point{ src_->points[i].x, src_->points[i].y, src_->points[i].z, src_->points[i].intensity };
        for (size_t i = 0; i < r.target.size(); i++) r.target[i] = llo_point{ tgt_->points[i].x, tgt_->points[i].y, tgt_->points[i].z, tgt_->points[i].intensity };
        llo_icp_align(r.source.data(), (int)r.source.size(), r.target.data(), (int)r.target.size(), max_iter_, max_dist_, teps_, feps_,
                      r.T, &r.converged, &r.iterations, &r.state, &r.fitness);
        r.calls++;
    }
    bool hasConverged() const { return llref_icp_last().converged != 0; }
    double getFitnessScore() const { return llref_icp_last().fitness; }
    Eigen::Matrix4f getFinalTransformation() const { Eigen::Matrix4f m; for (int i = 0; i < 16; i++) m.m[i] = llref_icp_last().T[i]; return m; }
private:
    typename PointCloud<A>::ConstPtr src_; typename PointCloud<B>::ConstPtr tgt_;
    double max_dist_ = 1e30, teps_ = 0, feps_ = 0; int max_iter_ = 10;
};
inline Eigen::Affine3f getTransformation(float, float, float, float, float, float) { return Eigen::Affine3f(); }
inline void getTranslationAndEulerAngles(const Eigen::Affine3f &, float &x, float &y, float &z, float &r, float &p, float &yw)
{ x = y = z = r = p = yw = 0; }
template <typename T, typename M>
inline void transformPointCloud(const PointCloud<T> &in, PointCloud<T> &out, const M &) { out = in; }
namespace io { template <typename T> inline int savePCDFileASCII(const std::string &, const PointCloud<T> &) { return 0; } }
}  // namespace pcl

// ============================================================ GTSAM (odometry chain only: loop closure is off, UT:104)
namespace gtsam {
typedef std::vector<double> VectorBase;
struct Vector : std::vector<double> {
    explicit Vector(int n = 0) : std::vector<double>(n, 0.0), pos_(0) {}
    Vector &operator<<(double v) { pos_ = 0; (*this)[pos_++] = v; return *this; }
    Vector &operator,(double v) { (*this)[pos_++] = v; return *this; }
    int pos_;
};
struct Point3 {
    double x_, y_, z_;
    Point3(double x = 0, double y = 0, double z = 0) : x_(x), y_(y), z_(z) {}
    double x() const { return x_; } double y() const { return y_; } double z() const { return z_; }
};
struct Rot3 {
    double R[3][3];
    Rot3() { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[i][j] = i == j; }
    // Rot3::RzRyRx(x, y, z): rotations about X, Y, Z composed as Rz * Ry * Rx (double)
    static Rot3 RzRyRx(double x, double y, double z)
    {
        double cx = std::cos(x), sx = std::sin(x), cy = std::cos(y), sy = std::sin(y), cz = std::cos(z), sz = std::sin(z);
        double ss_ = sx * sy, cs_ = cx * sy, sc_ = sx * cy, cc_ = cx * cy;
        double c_s = cx * sz, s_s = sx * sz, _cs = cy * sz, _cc = cy * cz;
        double s_c = sx * cz, c_c = cx * cz, ssc = ss_ * cz, csc = cs_ * cz, sss = ss_ * sz, css = cs_ * sz;
        Rot3 r;
        r.R[0][0] = _cc; r.R[0][1] = -c_s + ssc; r.R[0][2] = s_s + csc;
        r.R[1][0] = _cs; r.R[1][1] = c_c + sss; r.R[1][2] = -s_c + css;
        r.R[2][0] = -sy; r.R[2][1] = sc_; r.R[2][2] = cc_;
        return r;
    }
    // Rot3::rpy() via the RQ decomposition gtsam uses
    void rpy(double &x, double &y, double &z) const
    {
        x = -std::atan2(-R[2][1], R[2][2]);
        double cxm = std::cos(-x), sxm = std::sin(-x);       // B = A * Rx(-x)
        double B[3][3];
        for (int i = 0; i < 3; i++) { B[i][0] = R[i][0]; B[i][1] = R[i][1] * cxm + R[i][2] * sxm; B[i][2] = -R[i][1] * sxm + R[i][2] * cxm; }
        y = -std::atan2(B[2][0], B[2][2]);
        double cym = std::cos(-y), sym = std::sin(-y);       // C = B * Ry(-y)
        double C[3][3];
        for (int i = 0; i < 3; i++) { C[i][0] = B[i][0] * cym - B[i][2] * sym; C[i][1] = B[i][1]; C[i][2] = B[i][0] * sym + B[i][2] * cym; }
        z = -std::atan2(-C[1][0], C[1][1]);
    }
    double roll() const { double x, y, z; rpy(x, y, z); return x; }
    double pitch() const { double x, y, z; rpy(x, y, z); return y; }
    double yaw() const { double x, y, z; rpy(x, y, z); return z; }
};
struct Pose3 {
    Rot3 r; Point3 t;
    Pose3() {}
    Pose3(const Rot3 &R, const Point3 &T) : r(R), t(T) {}
    const Rot3 &rotation() const { return r; }
    const Point3 &translation() const { return t; }
    Pose3 between(const Pose3 &) const { return Pose3(); }   // only consumed by the (identity) iSAM2 stand-in
};
namespace noiseModel {
struct Diagonal {
    typedef std::shared_ptr<Diagonal> shared_ptr;
    static shared_ptr Variances(const Vector &) { return shared_ptr(new Diagonal()); }
};
}  // namespace noiseModel
template <typename T> struct PriorFactor { PriorFactor(int, const T &, const noiseModel::Diagonal::shared_ptr &) {} };
template <typename T> struct BetweenFactor { BetweenFactor(int, int, const T &, const noiseModel::Diagonal::shared_ptr &) {} };
struct NonlinearFactorGraph {
    template <typename F> void add(const F &) {}
    void resize(int) {}
};
struct Values {
    std::map<int, Pose3> v;
    void insert(int k, const Pose3 &p) { v[k] = p; }
    void clear() { v.clear(); }
    size_t size() const { return v.size(); }
    template <typename T> const T &at(int k) const { return v.at(k); }
};
struct ISAM2Params { double relinearizeThreshold = 0.1; int relinearizeSkip = 10; };
// With loop closure off the graph is a pure odometry chain whose optimum equals the inserted
// initial values: the stand-in just keeps them.
struct ISAM2 {
    Values all;
    explicit ISAM2(const ISAM2Params &) {}
    void update(const NonlinearFactorGraph &, const Values &init) { for (auto &kv : init.v) all.v[kv.first] = kv.second; }
    void update(const NonlinearFactorGraph &) {}
    void update() {}
    Values calculateEstimate() const { return all; }
};
}  // namespace gtsam
