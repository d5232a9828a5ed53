// keyframes.cu — key-frame arena + the fused transformPointCloud / concatenation kernel (see keyframes.cuh).
#include "keyframes.cuh"

namespace llb {

namespace {

// transformPointCloud MO:545-575, float arithmetic in the reference's association order (this file is compiled
// with -fmad=false); the six sin/cos are computed by the caller exactly as the reference does (host libm)
__global__ void __launch_bounds__(256)
kf_assemble_kernel(const AsmSeg *__restrict__ segs)
{
    const AsmSeg sg = segs[blockIdx.y];
    float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
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
        mn[0] = fminf(mn[0], o.x); mx[0] = fmaxf(mx[0], o.x);
        mn[1] = fminf(mn[1], o.y); mx[1] = fmaxf(mx[1], o.y);
        mn[2] = fminf(mn[2], o.z); mx[2] = fmaxf(mx[2], o.z);
    }
    if (!sg.bounds || (int)(blockIdx.x * blockDim.x) >= sg.n) return;     // uniform over the CTA
    // getMinMax3D of the map's voxel filter (voxel_minmax_kernel), folded into this pass: CTA reduce, 6 atomics
    __shared__ float s_red[6][256 / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(FULL, mn[a], off));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL, mx[a], off));
        }
        if (lane == 0) { s_red[a][w] = mn[a]; s_red[3 + a][w] = mx[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float m0 = s_red[threadIdx.x][0], m1 = s_red[3 + threadIdx.x][0];
        for (int k = 1; k < 256 / 32; k++) { m0 = fminf(m0, s_red[threadIdx.x][k]); m1 = fmaxf(m1, s_red[3 + threadIdx.x][k]); }
        atomicMin(&sg.bounds[threadIdx.x], float_to_ordered(m0));
        atomicMax(&sg.bounds[3 + threadIdx.x], float_to_ordered(m1));
    }
}

}  // namespace

void launch_kf_assemble(const AsmSeg *segs_dev, int count, int n_max, cudaStream_t s)
{
    if (count <= 0) return;
    const dim3 grid(std::max(1, std::min(div_up(std::max(n_max, 1), 256), 32)), count);
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
