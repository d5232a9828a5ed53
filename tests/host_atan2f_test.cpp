// Host build of lego_loam_b200/csrc/glibc_atan2f.cuh against the C library's atan2f (tests/test_host_atan2f.py).
#include "../lego_loam_b200/csrc/glibc_atan2f.cuh"
#include <cstdio>
#include <cstdlib>

extern "C" long host_atan2f_mismatches(long n, unsigned seed, float *first_bad /* y, x, libm, ours */)
{
    srand(seed);
    long bad = 0;
    const float special[] = { 0.f, -0.f, 1.f, -1.f, 1e-30f, -1e-30f, 1e30f, -1e30f, INFINITY, -INFINITY, 0.4375f, 0.6875f, 1.1875f, 2.4375f };
    const int ns = sizeof(special) / sizeof(special[0]);
    for (long i = 0; i < n; i++) {
        float y, x;
        if (i < ns * ns) { y = special[i / ns]; x = special[i % ns]; }
        else {
            y = (float)((rand() / (double)RAND_MAX - 0.5) * 200); x = (float)((rand() / (double)RAND_MAX - 0.5) * 200);
            if (i % 3 == 0) y *= 0.01f;
            if (i % 7 == 0) x *= 0.001f;
            if (i % 11 == 0) { y *= 1e-20f; }
        }
        const float a = atan2f(y, x), b = llb::glibcm::atan2f_(y, x);
        if (llb::glibcm::f2u(a) != llb::glibcm::f2u(b)) {
            if (bad == 0 && first_bad) { first_bad[0] = y; first_bad[1] = x; first_bad[2] = a; first_bad[3] = b; }
            bad++;
        }
    }
    return bad;
}
