// glibc_atan2f.cuh — atan2f as glibc computes it (sysdeps/ieee754/flt-32/e_atan2f.c + s_atanf.c, the fdlibm float code:
// argument reduction to [0, 7/16) with four break points, an 11-term odd/even polynomial, hi/lo constants), operation by
// operation in IEEE float arithmetic without contraction.  The reference takes the relative time of every point from
// `-atan2(y, x)` on float arguments (FA:504); glibc's result is within 1 ulp but not always the correctly rounded value,
// so only the same sequence of float operations reproduces the reference's intensities bit for bit
// (tests/test_host_atan2f.py: 0 differences against the host's libm over 2e7 arguments; the library is built with
// -fmad=false -prec-div=true).
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>

#ifndef __CUDACC__
#define LLB_AT_HD inline
#else
#define LLB_AT_HD __host__ __device__ __forceinline__
#endif

namespace llb {
namespace glibcm {

LLB_AT_HD uint32_t f2u(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; std::memcpy(&u, &f, 4); return u;
#endif
}
LLB_AT_HD float u2f(uint32_t u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; std::memcpy(&f, &u, 4); return f;
#endif
}

LLB_AT_HD float atanf_(float x)
{
    const float atanhi[4] = { 4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f };
    const float atanlo[4] = { 5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f };
    const float aT[11] = { 3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f, 9.0908870101e-02f,
                           -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f, -3.6531571299e-02f,
                           1.6285819933e-02f };
    const int32_t hx = (int32_t)f2u(x), ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c000000) {                       // |x| >= 2^25
        if (ix > 0x7f800000) return x + x;        // NaN
        return hx > 0 ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3];
    }
    if (ix < 0x3ee00000) {                        // |x| < 0.4375
        if (ix < 0x31000000) return x;            // |x| < 2^-29
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {                    // |x| < 1.1875
            if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
            else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
            else { id = 3; x = -1.0f / x; }
        }
    }
    const float z = x * x, w = z * z;
    const float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    const float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    const float r = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return hx < 0 ? -r : r;
}

LLB_AT_HD float atan2f_(float y, float x)
{
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
                pi_lo = -8.7422776573e-08f;
    const int32_t hx = (int32_t)f2u(x), ix = hx & 0x7fffffff, hy = (int32_t)f2u(y), iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return atanf_(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        if (m < 2) return y;
        return m == 2 ? pi + tiny : -pi - tiny;
    }
    if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) { case 0: return pi_o_4 + tiny; case 1: return -pi_o_4 - tiny; case 2: return 3.0f * pi_o_4 + tiny;
                         default: return -3.0f * pi_o_4 - tiny; }
        }
        switch (m) { case 0: return 0.0f; case 1: return -0.0f; case 2: return pi + tiny; default: return -pi - tiny; }
    }
    if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = atanf_(fabsf(y / x));
    switch (m) {
    case 0: return z;
    case 1: return u2f(f2u(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}

}  // namespace glibcm
}  // namespace llb
