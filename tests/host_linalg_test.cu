// Host-side unit test of lego_loam_b200/csrc/linalg.cuh: the functions are __host__ __device__,
// so the very code the kernels run is executed here on the CPU (g++ back end, no FMA contraction)
// and compared BIT-FOR-BIT with the oracle's OpenCV restatement (oracle/llo_linalg.c).
// Built and run by tests/test_host_linalg.py:  nvcc -O2 -fmad=false ... && ./host_linalg_test
#include "../lego_loam_b200/csrc/linalg.cuh"
extern "C" {
#include "../oracle/llo.h"
}
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <random>

static int bits_equal(const float *a, const float *b, int n) { return memcmp(a, b, sizeof(float) * n) == 0; }

int main()
{
    std::mt19937 rng(20181001);
    std::normal_distribution<float> N01(0.f, 1.f);
    int bad = 0, total = 0;
    const float scales[3] = { 0.01f, 1.f, 30.f };
    // ---- symmetric eigen 3x3 (generic + register-resident) and 6x6
    for (int t = 0; t < 20000; t++) {
        float B[5][3], A[9];
        float sc = scales[t % 3];
        for (auto &r : B) for (auto &v : r) v = N01(rng) * sc;
        if (t % 5 == 0) for (int i = 0; i < 5; i++) { B[i][1] = B[i][0] * 0.5f + 1e-3f * N01(rng); B[i][2] = 1e-3f * N01(rng); }   // near-collinear
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { float s = 0; for (int k = 0; k < 5; k++) s += B[k][i] * B[k][j]; A[i * 3 + j] = s; }
        for (int i = 0; i < 3; i++) for (int j = 0; j < i; j++) A[i * 3 + j] = A[j * 3 + i];
        float A1[9], A2[9], W0[3], V0[9], W1[3], V1[9], W2[3], V2[9];
        memcpy(A1, A, sizeof A); memcpy(A2, A, sizeof A);
        llo_cv_eigen_f32(3, A1, W0, V0);
        llb::cv_eigen<3>(A2, W1, V1);
        llb::cv_eigen3(A[0], A[1], A[2], A[4], A[5], A[8], W2, V2);
        total++;
        if (!bits_equal(W0, W1, 3) || !bits_equal(V0, V1, 9) || !bits_equal(W0, W2, 3) || !bits_equal(V0, V2, 9)) bad++;
    }
    printf("eigen3: %d / %d mismatches\n", bad, total);
    int bad6 = 0;
    for (int t = 0; t < 3000; t++) {
        float B[9][6], A[36];
        float sc = scales[t % 3];
        for (auto &r : B) for (auto &v : r) v = N01(rng) * sc;
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { float s = 0; for (int k = 0; k < 9; k++) s += B[k][i] * B[k][j]; A[i * 6 + j] = s; }
        for (int i = 0; i < 6; i++) for (int j = 0; j < i; j++) A[i * 6 + j] = A[j * 6 + i];
        float A1[36], A2[36], W0[6], V0[36], W1[6], V1[36];
        memcpy(A1, A, sizeof A); memcpy(A2, A, sizeof A);
        llo_cv_eigen_f32(6, A1, W0, V0);
        llb::cv_eigen<6>(A2, W1, V1);
        if (!bits_equal(W0, W1, 6) || !bits_equal(V0, V1, 36)) bad6++;
        // inverse + product
        float I0[36], I1[36], P0[36], P1[36];
        llo_cv_inv_f32(6, V0, I0); llb::cv_inv_lu<6>(V0, I1);
        llo_cv_gemm_f32(6, 6, 6, I0, V0, P0); llb::cv_gemm<6, 6, 6>(I1, V0, P1);
        if (!bits_equal(I0, I1, 36) || !bits_equal(P0, P1, 36)) bad6++;
        // square solve
        float b[6], x0[6], x1[6], Ac[36], bc[6];
        for (auto &v : b) v = N01(rng);
        llo_cv_solve_qr_f32(6, 6, A, b, x0);
        memcpy(Ac, A, sizeof A); memcpy(bc, b, sizeof b);
        llb::cv_solve_qr<6, 6>(Ac, bc, x1);
        if (!bits_equal(x0, x1, 6)) bad6++;
    }
    printf("6x6 eigen/inv/gemm/solve: %d mismatches\n", bad6);
    int bad53 = 0;
    for (int t = 0; t < 20000; t++) {
        float A[15], b[5] = { -1, -1, -1, -1, -1 }, x0[3], x1[3], Ac[15], bc[5];
        float n[3] = { N01(rng), N01(rng), N01(rng) };
        float nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        float off = 1.f + 30.f * std::fabs(N01(rng));
        for (int i = 0; i < 5; i++) {
            float p[3] = { N01(rng), N01(rng), N01(rng) };
            float d = (p[0] * n[0] + p[1] * n[1] + p[2] * n[2]) / nn;
            for (int k = 0; k < 3; k++) A[3 * i + k] = p[k] - d * n[k] / nn + off * n[k] / nn + 0.01f * N01(rng);
        }
        llo_cv_solve_qr_f32(5, 3, A, b, x0);
        memcpy(Ac, A, sizeof A); memcpy(bc, b, sizeof b);
        llb::cv_solve_qr<5, 3>(Ac, bc, x1);
        if (!bits_equal(x0, x1, 3)) bad53++;
        float A3[9], b3[3] = { N01(rng), N01(rng), N01(rng) }, y0[3], y1[3], I0[9], I1[9];
        for (auto &v : A3) v = N01(rng);
        llo_cv_solve_qr_f32(3, 3, A3, b3, y0);
        memcpy(Ac, A3, sizeof A3); memcpy(bc, b3, sizeof b3);
        llb::cv_solve_qr<3, 3>(Ac, bc, y1);
        llo_cv_inv_f32(3, A3, I0); llb::cv_inv3(A3, I1);
        if (!bits_equal(y0, y1, 3) || !bits_equal(I0, I1, 9)) bad53++;
    }
    printf("5x3 / 3x3 solve, inv3: %d mismatches\n", bad53);
    return (bad || bad6 || bad53) ? 1 : 0;
}
