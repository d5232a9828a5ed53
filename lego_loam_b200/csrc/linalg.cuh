// linalg.cuh — per-thread small dense CV_32F primitives on the device.
//
// The reference solves its tiny systems through OpenCV (cv::eigen MO:1126/1283,
// FA:1334/1435; cv::solve(DECOMP_QR) MO:1189/1276, FA:1327/1428; Mat::inv MO:1298,
// FA:1349/1450; Mat*Mat MO:1298/1304).  To land on the reference's bits these
// device versions perform the same IEEE float32 operations in the same order as
// OpenCV's Jacobi / Householder-QR / LU (the translation unit is compiled with
// -fmad=false, IEEE div and sqrt).  Sizes are template parameters so the fixed
// trip-count loops unroll; the Jacobi pivot bookkeeping indexes dynamically and
// lives in local memory, which is fine for one 3x3 per query and one 6x6 per
// registration.
#pragma once
#include <cuda_runtime.h>
#include <float.h>

// host + device: tests/host_linalg_test.cu runs the very same code on the CPU against the oracle
#define LLB_HD __host__ __device__

namespace llb {

LLB_HD __forceinline__ float cv_hypot(float a, float b)
{
    a = fabsf(a);
    b = fabsf(b);
    if (a > b) {
        b /= a;
        return a * sqrtf(1 + b * b);
    }
    if (b > 0) {
        a /= b;
        return b * sqrtf(1 + a * a);
    }
    return 0.f;
}

// Symmetric eigen-decomposition, eigenvalues descending in W, eigenvectors as rows of V.
template <int N>
LLB_HD void cv_eigen(float *A, float *W, float *V)
{
    const float eps = FLT_EPSILON;
    int indR[N], indC[N];
    int i, j, k, m;
    float mv;

#pragma unroll
    for (i = 0; i < N; i++) {
#pragma unroll
        for (j = 0; j < N; j++) V[i * N + j] = (i == j) ? 1.f : 0.f;
    }

    for (k = 0; k < N; k++) {
        W[k] = A[(N + 1) * k];
        if (k < N - 1) {
            for (m = k + 1, mv = fabsf(A[N * k + m]), i = k + 2; i < N; i++) {
                float val = fabsf(A[N * k + i]);
                if (mv < val) mv = val, m = i;
            }
            indR[k] = m;
        }
        if (k > 0) {
            for (m = 0, mv = fabsf(A[k]), i = 1; i < k; i++) {
                float val = fabsf(A[N * i + k]);
                if (mv < val) mv = val, m = i;
            }
            indC[k] = m;
        }
    }

    const int maxIters = N * N * 30;
    for (int iters = 0; iters < maxIters; iters++) {
        for (k = 0, mv = fabsf(A[indR[0]]), i = 1; i < N - 1; i++) {
            float val = fabsf(A[N * i + indR[i]]);
            if (mv < val) mv = val, k = i;
        }
        int l = indR[k];
        for (i = 1; i < N; i++) {
            float val = fabsf(A[N * indC[i] + i]);
            if (mv < val) mv = val, k = indC[i], l = i;
        }

        float p = A[N * k + l];
        if (fabsf(p) <= eps) break;
        float y = (float)((W[l] - W[k]) * 0.5);
        float t = fabsf(y) + cv_hypot(p, y);
        float s = cv_hypot(p, t);
        float c = t / s;
        s = p / s;
        t = (p / t) * p;
        if (y < 0) s = -s, t = -t;
        A[N * k + l] = 0;

        W[k] -= t;
        W[l] += t;

        float a0, b0;
#define LLB_ROT(v0, v1) (a0 = (v0), b0 = (v1), (v0) = a0 * c - b0 * s, (v1) = a0 * s + b0 * c)
        for (i = 0; i < k; i++) LLB_ROT(A[N * i + k], A[N * i + l]);
        for (i = k + 1; i < l; i++) LLB_ROT(A[N * k + i], A[N * i + l]);
        for (i = l + 1; i < N; i++) LLB_ROT(A[N * k + i], A[N * l + i]);
        for (i = 0; i < N; i++) LLB_ROT(V[N * k + i], V[N * l + i]);
#undef LLB_ROT

        for (j = 0; j < 2; j++) {
            int idx = j == 0 ? k : l;
            if (idx < N - 1) {
                for (m = idx + 1, mv = fabsf(A[N * idx + m]), i = idx + 2; i < N; i++) {
                    float val = fabsf(A[N * idx + i]);
                    if (mv < val) mv = val, m = i;
                }
                indR[idx] = m;
            }
            if (idx > 0) {
                for (m = 0, mv = fabsf(A[idx]), i = 1; i < idx; i++) {
                    float val = fabsf(A[N * i + idx]);
                    if (mv < val) mv = val, m = i;
                }
                indC[idx] = m;
            }
        }
    }

    for (k = 0; k < N - 1; k++) {
        m = k;
        for (i = k + 1; i < N; i++)
            if (W[m] < W[i]) m = i;
        if (k != m) {
            float tmp = W[m]; W[m] = W[k]; W[k] = tmp;
            for (i = 0; i < N; i++) {
                tmp = V[N * m + i]; V[N * m + i] = V[N * k + i]; V[N * k + i] = tmp;
            }
        }
    }
}

// cv_eigen<3> with everything in registers.  Same arithmetic, same pivot choices (including
// OpenCV's stale indR/indC bookkeeping: for n = 3 only indR[0] and indC[2] are variable), but
// the (k,l) pivot is dispatched over its three possible values instead of indexing arrays, so
// one thread per query runs without local memory.  A = {a00,a01,a02,a11,a12,a22} upper triangle.
LLB_HD inline void cv_eigen3(float a00, float a01, float a02, float a11, float a12, float a22, float *W, float *V)
{
    const float eps = FLT_EPSILON;
    float w0 = a00, w1 = a11, w2 = a22;
    float v00 = 1.f, v01 = 0.f, v02 = 0.f, v10 = 0.f, v11 = 1.f, v12 = 0.f, v20 = 0.f, v21 = 0.f, v22 = 1.f;
    int indR0 = (fabsf(a01) < fabsf(a02)) ? 2 : 1;           // indR[1] == 2, indC[1] == 0 always
    int indC2 = (fabsf(a02) < fabsf(a12)) ? 1 : 0;

    for (int iters = 0; iters < 3 * 3 * 30; iters++) {
        // pivot search, in OpenCV's order: rows via indR, then columns via indC
        int k = 0, l;
        float mv = fabsf(indR0 == 1 ? a01 : a02);
        {
            float val = fabsf(a12);
            if (mv < val) mv = val, k = 1;
        }
        l = (k == 0) ? indR0 : 2;
        {
            float val = fabsf(a01);                          // i = 1: A[indC[1]][1] = A[0][1]
            if (mv < val) mv = val, k = 0, l = 1;
            val = fabsf(indC2 == 0 ? a02 : a12);             // i = 2: A[indC[2]][2]
            if (mv < val) mv = val, k = indC2, l = 2;
        }
        const int code = (k == 0) ? (l == 1 ? 0 : 1) : 2;    // (0,1) (0,2) (1,2)
        const float p = code == 0 ? a01 : (code == 1 ? a02 : a12);
        if (fabsf(p) <= eps) break;
        const float wk = (code == 2) ? w1 : w0, wl = (code == 0) ? w1 : w2;
        float y = (float)((wl - wk) * 0.5);
        float t = fabsf(y) + cv_hypot(p, y);
        float s = cv_hypot(p, t);
        float c = t / s;
        s = p / s;
        t = (p / t) * p;
        if (y < 0) s = -s, t = -t;
        float a0, b0;
#define LLB_ROT(x0, x1) (a0 = (x0), b0 = (x1), (x0) = a0 * c - b0 * s, (x1) = a0 * s + b0 * c)
        if (code == 0) {
            a01 = 0; w0 -= t; w1 += t;
            LLB_ROT(a02, a12);                               // i > l: (A[k][i], A[l][i])
            LLB_ROT(v00, v10); LLB_ROT(v01, v11); LLB_ROT(v02, v12);
        } else if (code == 1) {
            a02 = 0; w0 -= t; w2 += t;
            LLB_ROT(a01, a12);                               // k < i < l: (A[k][i], A[i][l])
            LLB_ROT(v00, v20); LLB_ROT(v01, v21); LLB_ROT(v02, v22);
        } else {
            a12 = 0; w1 -= t; w2 += t;
            LLB_ROT(a01, a02);                               // i < k: (A[i][k], A[i][l])
            LLB_ROT(v10, v20); LLB_ROT(v11, v21); LLB_ROT(v12, v22);
        }
#undef LLB_ROT
        if (k == 0) indR0 = (fabsf(a01) < fabsf(a02)) ? 2 : 1;
        if (l == 2) indC2 = (fabsf(a02) < fabsf(a12)) ? 1 : 0;
    }
    // selection sort, descending, rows of V follow
    W[0] = w0; W[1] = w1; W[2] = w2;
    V[0] = v00; V[1] = v01; V[2] = v02; V[3] = v10; V[4] = v11; V[5] = v12; V[6] = v20; V[7] = v21; V[8] = v22;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        int m = k;
#pragma unroll
        for (int i = k + 1; i < 3; i++)
            if (W[m] < W[i]) m = i;
        if (k != m) {
            float tmp = W[m]; W[m] = W[k]; W[k] = tmp;
#pragma unroll
            for (int i = 0; i < 3; i++) { tmp = V[3 * m + i]; V[3 * m + i] = V[3 * k + i]; V[3 * k + i] = tmp; }
        }
    }
}

// Least squares / square solve by Householder QR.  A (M x N, row-major) and b (M) are
// destroyed; x receives N values.  Returns false (x = 0) for a singular system.
template <int M, int N>
LLB_HD bool cv_solve_qr(float *A, float *b, float *x)
{
    const float eps = FLT_EPSILON * 10;
    float vl[M], hF[N];
#pragma unroll
    for (int l = 0; l < N; l++) {
        float vlNorm = 0.f;
#pragma unroll
        for (int i = 0; i < M - l; i++) {
            vl[i] = A[(l + i) * N + l];
            vlNorm += vl[i] * vl[i];
        }
        float tmpV = vl[0];
        vl[0] = vl[0] + (vl[0] >= 0 ? 1 : -1) * sqrtf(vlNorm);
        vlNorm = sqrtf(vlNorm + vl[0] * vl[0] - tmpV * tmpV);
#pragma unroll
        for (int i = 0; i < M - l; i++) vl[i] /= vlNorm;
#pragma unroll
        for (int j = l; j < N; j++) {
            float v_lA = 0.f;
#pragma unroll
            for (int i = l; i < M; i++) v_lA += vl[i - l] * A[i * N + j];
#pragma unroll
            for (int i = l; i < M; i++) A[i * N + j] -= 2 * vl[i - l] * v_lA;
        }
        hF[l] = vl[0] * vl[0];
#pragma unroll
        for (int i = 1; i < M - l; i++) A[(l + i) * N + l] = vl[i] / vl[0];
    }
#pragma unroll
    for (int l = 0; l < N; l++) {
        vl[0] = 1.f;
#pragma unroll
        for (int j = 1; j < M - l; j++) vl[j] = A[(j + l) * N + l];
        float v_lB = 0.f;
#pragma unroll
        for (int i = l; i < M; i++) v_lB += vl[i - l] * b[i];
#pragma unroll
        for (int i = l; i < M; i++) b[i] -= 2 * vl[i - l] * v_lB * hF[l];
    }
    bool ok = true;
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {
        if (ok) {
#pragma unroll
            for (int j = N - 1; j > i; j--) b[i] -= b[j] * A[i * N + j];
            if (fabsf(A[i * N + i]) < eps) ok = false;
            else b[i] /= A[i * N + i];
        }
    }
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = ok ? b[i] : 0.f;
    return ok;
}

// D (M x Nn) = A (M x K) * B (K x Nn), double accumulation, one rounding (cv::gemm CV_32F)
template <int M, int K, int Nn>
LLB_HD void cv_gemm(const float *A, const float *B, float *D)
{
    for (int i = 0; i < M; i++)
        for (int j = 0; j < Nn; j++) {
            double s = 0;
            for (int l = 0; l < K; l++) s += (double)A[i * K + l] * (double)B[l * Nn + j];
            D[i * Nn + j] = (float)s;
        }
}

// Mat::inv() DECOMP_LU, 3x3: closed form in double
LLB_HD inline bool cv_inv3(const float *A, float *D)
{
#define S(r, c) ((double)A[(r) * 3 + (c)])
    double d = S(0,0) * (S(1,1) * S(2,2) - S(1,2) * S(2,1)) -
               S(0,1) * (S(1,0) * S(2,2) - S(1,2) * S(2,0)) +
               S(0,2) * (S(1,0) * S(2,1) - S(1,1) * S(2,0));
    if (d != 0.) {
        d = 1. / d;
        D[0] = (float)((S(1,1) * S(2,2) - S(1,2) * S(2,1)) * d);
        D[1] = (float)((S(0,2) * S(2,1) - S(0,1) * S(2,2)) * d);
        D[2] = (float)((S(0,1) * S(1,2) - S(0,2) * S(1,1)) * d);
        D[3] = (float)((S(1,2) * S(2,0) - S(1,0) * S(2,2)) * d);
        D[4] = (float)((S(0,0) * S(2,2) - S(0,2) * S(2,0)) * d);
        D[5] = (float)((S(0,2) * S(1,0) - S(0,0) * S(1,2)) * d);
        D[6] = (float)((S(1,0) * S(2,1) - S(1,1) * S(2,0)) * d);
        D[7] = (float)((S(0,1) * S(2,0) - S(0,0) * S(2,1)) * d);
        D[8] = (float)((S(0,0) * S(1,1) - S(0,1) * S(1,0)) * d);
        return true;
    }
#undef S
    for (int i = 0; i < 9; i++) D[i] = 0.f;
    return false;
}

// Mat::inv() DECOMP_LU for N > 3: LU with partial pivoting against the identity, float
template <int N>
LLB_HD bool cv_inv_lu(const float *Ain, float *D)
{
    const float eps = FLT_EPSILON * 10;
    float a[N * N], b[N * N];
    int i, j, k;
    for (i = 0; i < N * N; i++) a[i] = Ain[i];
    for (i = 0; i < N; i++)
        for (j = 0; j < N; j++) b[i * N + j] = (i == j) ? 1.f : 0.f;
    for (i = 0; i < N; i++) {
        k = i;
        for (j = i + 1; j < N; j++)
            if (fabsf(a[j * N + i]) > fabsf(a[k * N + i])) k = j;
        if (fabsf(a[k * N + i]) < eps) {
            for (j = 0; j < N * N; j++) D[j] = 0.f;
            return false;
        }
        if (k != i) {
            for (j = i; j < N; j++) { float t = a[i * N + j]; a[i * N + j] = a[k * N + j]; a[k * N + j] = t; }
            for (j = 0; j < N; j++) { float t = b[i * N + j]; b[i * N + j] = b[k * N + j]; b[k * N + j] = t; }
        }
        float d = -1 / a[i * N + i];
        for (j = i + 1; j < N; j++) {
            float alpha = a[j * N + i] * d;
            for (k = i + 1; k < N; k++) a[j * N + k] += alpha * a[i * N + k];
            for (k = 0; k < N; k++) b[j * N + k] += alpha * b[i * N + k];
        }
    }
    for (i = N - 1; i >= 0; i--)
        for (j = 0; j < N; j++) {
            float s = b[i * N + j];
            for (k = i + 1; k < N; k++) s -= a[i * N + k] * b[k * N + j];
            b[i * N + j] = s / a[i * N + i];
        }
    for (i = 0; i < N * N; i++) D[i] = b[i];
    return true;
}

}  // namespace llb
