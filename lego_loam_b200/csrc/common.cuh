// common.cuh — shared host/device helpers for the llb200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <limits.h>
#include <float.h>
#include <algorithm>
#include <string>
#include <stdexcept>

namespace llb {

struct CudaError : std::runtime_error {
    cudaError_t code;
    CudaError(cudaError_t c, const char *what, const char *file, int line)
        : std::runtime_error(std::string(what) + ": " + cudaGetErrorString(c) + " (" + file + ":" + std::to_string(line) + ")"),
          code(c) {}
};

#define LLB_CUDA(expr)                                                         \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) throw ::llb::CudaError(_e, #expr, __FILE__, __LINE__); \
    } while (0)

// growable device buffer; growth synchronises the device (steady state never grows)
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    void ensure(size_t n)
    {
        if (n <= cap) return;
        size_t want = n + n / 4 + 64;
        if (p) { LLB_CUDA(cudaDeviceSynchronize()); LLB_CUDA(cudaFree(p)); p = nullptr; cap = 0; }
        LLB_CUDA(cudaMalloc(&p, want * sizeof(T)));
        cap = want;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
    }
};

// growable pinned host buffer
template <typename T>
struct PinnedBuf {
    T *p = nullptr;
    size_t cap = 0;
    void ensure(size_t n)
    {
        if (n <= cap) return;
        size_t want = n + n / 4 + 64;
        if (p) { LLB_CUDA(cudaDeviceSynchronize()); LLB_CUDA(cudaFreeHost(p)); p = nullptr; cap = 0; }
        LLB_CUDA(cudaMallocHost(&p, want * sizeof(T)));
        cap = want;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
    }
};

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__

constexpr unsigned FULL = 0xffffffffu;

// order-preserving float <-> int mapping for atomicMin/atomicMax on floats
__device__ __forceinline__ int float_to_ordered(float f)
{
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i)
{
    return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

// lanes of the warp that hold the same value in the low BITS bits of v (what __match_any_sync returns), from BITS
// ballots: on sm_100 the result latency of MATCH.ANY dominated the radix ranking loops (half of all stall samples,
// profiles/r02_voxel.md); the ballots of independent keys overlap
template <int BITS>
__device__ __forceinline__ unsigned match_low_bits(unsigned v)
{
    unsigned m = FULL;
#pragma unroll
    for (int b = 0; b < BITS; b++) {
        const bool bit = (v >> b) & 1u;
        const unsigned bal = __ballot_sync(FULL, bit);
        m &= bit ? bal : ~bal;
    }
    return m;
}

__device__ __forceinline__ int warp_incl_scan(int v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// exclusive scan of one int per thread over the whole block (blockDim.x multiple of 32,
// <= 1024).  `total` receives the block sum.  smem must hold 33 ints.
__device__ __forceinline__ int block_excl_scan(int v, int *smem, int &total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = warp_incl_scan(v);
    if (lane == 31) smem[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < nw ? smem[lane] : 0;
        int si = warp_incl_scan(s);
        smem[lane] = si - s;
        if (lane == 31) smem[32] = si;
    }
    __syncthreads();
    int r = smem[w] + inc - v;
    total = smem[32];
    __syncthreads();
    return r;
}

#endif  // __CUDACC__

}  // namespace llb
