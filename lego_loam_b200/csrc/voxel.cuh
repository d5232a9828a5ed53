// voxel.cuh — K1: pcl::VoxelGrid<PointXYZI>-exact voxel-grid down-sampling on the device.
// Replaces the filter() calls at MO:1058-1063 (local map), MO:1070-1089 (current scan)
// and FA:779-780.  Contract: bit-exact voxel membership, output
// order (ascending voxel index) and centroids (float sums in ascending input order, true
// division by (float)count, intensity averaged too), int32-overflow pass-through.
#pragma once
#include "common.cuh"

namespace llb {

// input = concatenation of up to two device segments whose lengths may live on the device
struct VoxelInput {
    const float4 *a = nullptr; const int *na_dev = nullptr; int na = 0;
    const float4 *b = nullptr; const int *nb_dev = nullptr; int nb = 0;
    int upper() const { return na + nb; }   // host-side upper bound of the length
};

struct VoxelDesc {           // device-resident, written by the setup step
    int mn[3], mx[3];        // ordered-int encoded min / max (atomics), reset after use
    float inv;
    int min_b[3], div_b[3], mul[3];
    int overflow;
    int n;                   // input length
    int nbits;               // significant key bits
    int n_out;
};

struct LargeVoxelJob;        // one filter of the multi-kernel path (voxel_dev.cuh)

class VoxelFilter {
public:
    static constexpr int SMALL_MAX = 16384;     // single-CTA path up to this many points
    static constexpr int MAX_BATCH = 4;         // independent small filters sharing one launch
    void init();
    void release();
    // batched form of the multi-kernel path: large_job() sizes this filter's scratch for the input's upper bound and
    // returns the job record; a device-resident table of such records is run by ONE set of 18 launches
    LargeVoxelJob large_job(const VoxelInput &in, float leaf, float4 *out, int *n_out_dev, const VoxelInput *bounds = nullptr);
    static int launch_large(const LargeVoxelJob *table_dev, int count, int n_upper, cudaStream_t s);
    // pre-sizes the scratch of the multi-kernel path for inputs of up to n points (no allocation at run time below n)
    void reserve(int n);
    // out must have room for in.upper() points; n_out_dev receives the output count.
    // All work is enqueued on `stream`; nothing synchronises. Returns kernels launched.
    int run(const VoxelInput &in, float leaf, float4 *out, int *n_out_dev, cudaStream_t stream);
    // the same through the multi-kernel path whatever the size, with the lattice bounds taken from ANOTHER cloud (`bounds`,
    // of which `in` is a part): a rank of a sharded map filters its share and numbers the voxels as the whole map does
    int run_with_bounds(const VoxelInput &in, const VoxelInput &bounds, float leaf, float4 *out, int *n_out_dev,
                        cudaStream_t stream) { return run_large(in, leaf, out, n_out_dev, stream, &bounds); }
    // `count` independent filters; when every one fits the single-CTA path they share ONE launch
    int run_batch(const VoxelInput *in, const float *leaf, float4 *const *out, int *const *n_out_dev,
                  int count, cudaStream_t stream);

private:
    int run_large(const VoxelInput &in, float leaf, float4 *out, int *n_out_dev, cudaStream_t s, const VoxelInput *bounds = nullptr);
    DevBuf<VoxelDesc> desc_;
    DevBuf<unsigned> keys_[2];
    DevBuf<int> vals_[2];
    DevBuf<int> hist_;          // radix histograms: 256 bins x nblocks
    DevBuf<int> blk_;           // per-block head counts / offsets
    DevBuf<unsigned char> job_raw_;
    PinnedBuf<unsigned char> job_pin_;
    cudaEvent_t job_ev_ = nullptr;
    bool small_attr_set_ = false;
};

}  // namespace llb
