// keyframes.cuh — device-resident key-frame store: the clouds the reference keeps in
// cornerCloudKeyFrames / surfCloudKeyFrames / outlierCloudKeyFrames (MO:128-130, filled by saveKeyFramesAndFactor
// MO:1443-1453) live in HBM from the moment downsampleCurrentScan produced them, and the cloud work of
// extractSurroundingKeyFrames (MO:962-1001 / MO:1033-1056: transformPointCloud of every selected key-frame,
// MO:545-575, + concatenation) is ONE launch that writes the raw local map straight into the buffers the map
// voxel filters (MO:1057-1064) read.  A registration then moves only the new sweep over PCIe.
#pragma once
#include "common.cuh"
#include <vector>

namespace llb {

struct AsmSeg {                       // one key-frame cloud -> its slice of an assembled raw map
    const float4 *src; float4 *dst; int n;
    float ctRoll, stRoll, ctPitch, stPitch, ctYaw, stYaw, tx, ty, tz;   // updateTransformPointCloudSinCos MO:529-543
};

struct KeyFrameRec { const float4 *cloud[3]; int n[3]; };               // corner, surf, outlier (DS clouds)

class KeyFrameStore {
public:
    void release();
    void clear();
    int size() const { return (int)recs_.size(); }
    const KeyFrameRec &rec(int i) const { return recs_[i]; }
    // reserves room for the three clouds of a new key-frame and returns where to copy them
    int add(const int n[3], float4 *dst[3]);
    // makes room for n_points more key-frame points without a further allocation
    void reserve(size_t n_points);
private:
    static constexpr size_t CHUNK = (size_t)4 << 20;                    // points per arena chunk (64 MB)
    float4 *alloc(size_t n);
    std::vector<float4 *> chunks_;
    size_t used_ = 0, cap_ = 0;
    std::vector<KeyFrameRec> recs_;
};

// segs: device-resident table; n_max: longest segment (sizes the grid)
void launch_kf_assemble(const AsmSeg *segs_dev, int count, int n_max, cudaStream_t s);

}  // namespace llb
