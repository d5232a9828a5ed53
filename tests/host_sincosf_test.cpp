// Host build of lego_loam_b200/csrc/glibc_sincosf.cuh against the C library's sinf / cosf (tests/test_host_sincosf.py).
#include "../lego_loam_b200/csrc/glibc_sincosf.cuh"
#include <cstdlib>

extern "C" void host_sincosf_mismatches(long n, unsigned seed, long out[4])
{
    srand(seed);
    out[0] = out[1] = out[2] = out[3] = 0;
    for (long i = 0; i < n; i++) {
        const double span = i % 4 == 0 ? 200.0 : (i % 4 == 1 ? 8.0 : (i % 4 == 2 ? 0.5 : 0.02));
        const float y = (float)((rand() / (double)RAND_MAX - 0.5) * span);
        const float a = sinf(y), b = llb::glibcm::sinf_(y), c = cosf(y), d = llb::glibcm::cosf_(y);
        out[0] += llb::glibcm::sc_f2u(a) != llb::glibcm::sc_f2u(b);
        out[1] += llb::glibcm::sc_f2u(c) != llb::glibcm::sc_f2u(d);
        out[2] += llb::glibcm::sc_f2u(a) != llb::glibcm::sc_f2u((float)sin((double)y));     // libm vs correctly rounded
        out[3] += llb::glibcm::sc_f2u(c) != llb::glibcm::sc_f2u((float)cos((double)y));
    }
}
