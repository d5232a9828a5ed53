// keyframes.cu — key-frame arena + the fused transformPointCloud / concatenation kernel (see keyframes.cuh).
#include "keyframes.cuh"
#include <cstdlib>

namespace llb {

namespace {

// transformPointCloud MO:545-575, float arithmetic in the reference's association order (this file is compiled
// with -fmad=false); the six sin/cos are computed by the caller exactly as the reference does (host libm)
__global__ void __launch_bounds__(256)
kf_assemble_kernel(const AsmSeg *__restrict__ segs)
{
    const AsmSeg sg = segs[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < sg.n; i += gridDim.x * blockDim.x) {
        const float4 p = __ldg(&sg.src[i]);
        const float x1 = sg.ctYaw * p.x - sg.stYaw * p.y;
        const float y1 = sg.stYaw * p.x + sg.ctYaw * p.y;
        const float z1 = p.z;
        const float x2 = x1;
        const float y2 = sg.ctRoll * y1 - sg.stRoll * z1;
        const float z2 = sg.stRoll * y1 + sg.ctRoll * z1;
        float4 o;
        o.x = sg.ctPitch * x2 + sg.stPitch * z2 + sg.tx;
        o.y = y2 + sg.ty;
        o.z = -sg.stPitch * x2 + sg.ctPitch * z2 + sg.tz;
        o.w = p.w;
        sg.dst[i] = o;
    }
}

}  // namespace

void launch_kf_assemble(const AsmSeg *segs_dev, int count, int n_max, cudaStream_t s)
{
    if (count <= 0) return;
    // points per CTA: a key-frame cloud has ~1-4k points and a table thousands of segments; fewer, fuller CTAs
    static const int per_cta = getenv("LLB_ASM_PTS_PER_CTA") ? std::max(256, atoi(getenv("LLB_ASM_PTS_PER_CTA"))) : 1024;
    const dim3 grid(std::max(1, std::min(div_up(std::max(n_max, 1), per_cta), 32)), count);
    kf_assemble_kernel<<<grid, 256, 0, s>>>(segs_dev);
    LLB_CUDA(cudaGetLastError());
}

float4 *KeyFrameStore::alloc(size_t n)
{
    if (chunks_.empty() || used_ + n > cap_) {
        const size_t want = std::max(n, CHUNK);
        float4 *p = nullptr;
        LLB_CUDA(cudaMalloc(&p, want * sizeof(float4)));
        chunks_.push_back(p);
        used_ = 0; cap_ = want;
    }
    float4 *r = chunks_.back() + used_;
    used_ += n;
    return r;
}

void KeyFrameStore::reserve(size_t n_points)
{
    if (!chunks_.empty() && cap_ - used_ >= n_points) return;
    float4 *p = nullptr;
    LLB_CUDA(cudaMalloc(&p, std::max(n_points, CHUNK) * sizeof(float4)));
    chunks_.push_back(p);
    used_ = 0; cap_ = std::max(n_points, CHUNK);
}

int KeyFrameStore::add(const int n[3], float4 *dst[3])
{
    KeyFrameRec r;
    const size_t tot = (size_t)std::max(n[0], 0) + std::max(n[1], 0) + std::max(n[2], 0);
    float4 *base = alloc(std::max<size_t>(tot, 1));
    size_t off = 0;
    for (int k = 0; k < 3; k++) {
        dst[k] = base + off; r.cloud[k] = dst[k]; r.n[k] = std::max(n[k], 0);
        off += (size_t)r.n[k];
    }
    recs_.push_back(r);
    return (int)recs_.size() - 1;
}

void KeyFrameStore::clear()
{
    recs_.clear();
    for (size_t i = 0; i + 1 < chunks_.size(); i++) cudaFree(chunks_[i]);     // keep the newest chunk for reuse
    if (!chunks_.empty()) { float4 *last = chunks_.back(); chunks_.clear(); chunks_.push_back(last); }
    used_ = 0;
}

void KeyFrameStore::release()
{
    for (float4 *p : chunks_) cudaFree(p);
    chunks_.clear(); recs_.clear(); used_ = cap_ = 0;
}

}  // namespace llb
