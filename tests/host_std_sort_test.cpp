// Host build of the product's std::sort restatement (lego_loam_b200/csrc/std_sort.cuh), for tests/test_host_std_sort.py.
#include "../lego_loam_b200/csrc/std_sort.cuh"
#include <cstring>
#include <vector>

extern "C" void host_std_sort(float *value, unsigned *ind, int n, int depth_limit, int closed_form)
{
    std::vector<llb::stdsort::rec_t> r(n > 0 ? n : 1);
    for (int i = 0; i < n; i++) { unsigned b; std::memcpy(&b, &value[i], 4); r[i] = ((unsigned long long)b << 32) | ind[i]; }
    if (closed_form) {
        std::vector<unsigned short> L(n + 1), R(n + 1);
        llb::stdsort::sort_closed(r.data(), n, L.data(), R.data(), depth_limit);
    } else llb::stdsort::sort(r.data(), n, depth_limit);
    for (int i = 0; i < n; i++) { unsigned b = (unsigned)(r[i] >> 32); std::memcpy(&value[i], &b, 4); ind[i] = (unsigned)r[i]; }
}
