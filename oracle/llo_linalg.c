/*
 * llo_linalg.c — ORACLE (test infrastructure): restatement of the small dense
 * OpenCV CV_32F primitives the reference calls on the hot path.  OpenCV is an
 * un-vendored, un-pinned dependency of the reference (find_package(OpenCV),
 * /root/reference/LeGO-LOAM/CMakeLists.txt:25); the algorithms below are the
 * ones published in OpenCV's modules/core (lapack.cpp JacobiImpl_/hypot,
 * matrix_decomp.cpp LUImpl, hal QRImpl, matmul GEMMSingleMul) and are pinned
 * bit-for-bit against the OpenCV 4.13 wheel in tests/test_oracle_linalg.py.
 *
 * Call sites in the reference:
 *   cv::eigen   MO:1126 (3x3), MO:1283 (6x6), FA:1334, FA:1435 (3x3)
 *   cv::solve   MO:1189 (5x3 QR), MO:1276 (6x6 QR), FA:1327, FA:1428 (3x3 QR)
 *   Mat::inv    MO:1298 (6x6), FA:1349, FA:1450 (3x3)
 *   Mat * Mat   MO:1274-1275, MO:1298, MO:1304, FA:1325-1326, FA:1349, FA:1355 ...
 */
#include "llo.h"
#include <math.h>
#include <float.h>
#include <string.h>

#define LLO_MAXN 8

static int g_trig_mode = 0;
void llo_set_trig_mode(int mode) { g_trig_mode = mode; }
float llo_sinf(float x) { return g_trig_mode ? (float)sin((double)x) : sinf(x); }
float llo_cosf(float x) { return g_trig_mode ? (float)cos((double)x) : cosf(x); }

static float llo_hypotf(float a, float b)
{
    /* OpenCV's own hypot() template, not libm's */
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
    return 0;
}

int llo_cv_eigen_f32(int n, float *A, float *W, float *V)
{
    const float eps = FLT_EPSILON;
    int i, j, k, m;
    int indR[LLO_MAXN], indC[LLO_MAXN];
    float mv = 0;
    if (n < 1 || n > LLO_MAXN) return 0;

    for (i = 0; i < n; i++) {
        for (j = 0; j < n; j++) V[i * n + j] = 0;
        V[i * n + i] = 1;
    }

    for (k = 0; k < n; k++) {
        W[k] = A[(n + 1) * k];
        if (k < n - 1) {
            for (m = k + 1, mv = fabsf(A[n * k + m]), i = k + 2; i < n; i++) {
                float val = fabsf(A[n * k + i]);
                if (mv < val) mv = val, m = i;
            }
            indR[k] = m;
        }
        if (k > 0) {
            for (m = 0, mv = fabsf(A[k]), i = 1; i < k; i++) {
                float val = fabsf(A[n * i + k]);
                if (mv < val) mv = val, m = i;
            }
            indC[k] = m;
        }
    }

    if (n > 1) {
        int iters, maxIters = n * n * 30;
        for (iters = 0; iters < maxIters; iters++) {
            /* pivot = largest off-diagonal element */
            for (k = 0, mv = fabsf(A[indR[0]]), i = 1; i < n - 1; i++) {
                float val = fabsf(A[n * i + indR[i]]);
                if (mv < val) mv = val, k = i;
            }
            int l = indR[k];
            for (i = 1; i < n; i++) {
                float val = fabsf(A[n * indC[i] + i]);
                if (mv < val) mv = val, k = indC[i], l = i;
            }

            float p = A[n * k + l];
            if (fabsf(p) <= eps) break;
            float y = (float)((W[l] - W[k]) * 0.5);
            float t = fabsf(y) + llo_hypotf(p, y);
            float s = llo_hypotf(p, t);
            float c = t / s;
            s = p / s;
            t = (p / t) * p;
            if (y < 0) s = -s, t = -t;
            A[n * k + l] = 0;

            W[k] -= t;
            W[l] += t;

            float a0, b0;
#define LLO_ROT(v0, v1) (a0 = (v0), b0 = (v1), (v0) = a0 * c - b0 * s, (v1) = a0 * s + b0 * c)
            for (i = 0; i < k; i++) LLO_ROT(A[n * i + k], A[n * i + l]);
            for (i = k + 1; i < l; i++) LLO_ROT(A[n * k + i], A[n * i + l]);
            for (i = l + 1; i < n; i++) LLO_ROT(A[n * k + i], A[n * l + i]);
            for (i = 0; i < n; i++) LLO_ROT(V[n * k + i], V[n * l + i]);
#undef LLO_ROT

            for (j = 0; j < 2; j++) {
                int idx = j == 0 ? k : l;
                if (idx < n - 1) {
                    for (m = idx + 1, mv = fabsf(A[n * idx + m]), i = idx + 2; i < n; i++) {
                        float val = fabsf(A[n * idx + i]);
                        if (mv < val) mv = val, m = i;
                    }
                    indR[idx] = m;
                }
                if (idx > 0) {
                    for (m = 0, mv = fabsf(A[idx]), i = 1; i < idx; i++) {
                        float val = fabsf(A[n * i + idx]);
                        if (mv < val) mv = val, m = i;
                    }
                    indC[idx] = m;
                }
            }
        }
    }

    /* selection sort, descending */
    for (k = 0; k < n - 1; k++) {
        m = k;
        for (i = k + 1; i < n; i++)
            if (W[m] < W[i]) m = i;
        if (k != m) {
            float tmp = W[m]; W[m] = W[k]; W[k] = tmp;
            for (i = 0; i < n; i++) {
                tmp = V[n * m + i]; V[n * m + i] = V[n * k + i]; V[n * k + i] = tmp;
            }
        }
    }
    return 1;
}

int llo_cv_solve_qr_f32(int m, int n, const float *Ain, const float *bin, float *x)
{
    float A[LLO_MAXN * LLO_MAXN], b[LLO_MAXN], vl[LLO_MAXN], hF[LLO_MAXN];
    const float eps = FLT_EPSILON * 10;
    int i, j, l;
    if (m > LLO_MAXN || n > LLO_MAXN || m < n) return 0;
    memcpy(A, Ain, sizeof(float) * m * n);
    memcpy(b, bin, sizeof(float) * m);

    for (l = 0; l < n; l++) {
        int vlSize = m - l;
        float vlNorm = 0.f;
        for (i = 0; i < vlSize; i++) {
            vl[i] = A[(l + i) * n + l];
            vlNorm += vl[i] * vl[i];
        }
        float tmpV = vl[0];
        vl[0] = vl[0] + (vl[0] >= 0 ? 1 : -1) * sqrtf(vlNorm);
        vlNorm = sqrtf(vlNorm + vl[0] * vl[0] - tmpV * tmpV);
        for (i = 0; i < vlSize; i++) vl[i] /= vlNorm;

        for (j = l; j < n; j++) {
            float v_lA = 0.f;
            for (i = l; i < m; i++) v_lA += vl[i - l] * A[i * n + j];
            for (i = l; i < m; i++) A[i * n + j] -= 2 * vl[i - l] * v_lA;
        }

        hF[l] = vl[0] * vl[0];
        for (i = 1; i < vlSize; i++) A[(l + i) * n + l] = vl[i] / vl[0];
    }

    for (l = 0; l < n; l++) {
        vl[0] = 1.f;
        for (j = 1; j < m - l; j++) vl[j] = A[(j + l) * n + l];
        float v_lB = 0.f;
        for (i = l; i < m; i++) v_lB += vl[i - l] * b[i];
        for (i = l; i < m; i++) b[i] -= 2 * vl[i - l] * v_lB * hF[l];
    }
    for (i = n - 1; i >= 0; i--) {
        for (j = n - 1; j > i; j--) b[i] -= b[j] * A[i * n + j];
        if (fabsf(A[i * n + i]) < eps) {
            for (j = 0; j < n; j++) x[j] = 0.f;   /* cv::solve: if(!result) dst = Scalar(0) */
            return 0;
        }
        b[i] /= A[i * n + i];
    }
    for (i = 0; i < n; i++) x[i] = b[i];
    return 1;
}

void llo_cv_gemm_f32(int m, int k, int n, const float *A, const float *B, float *D)
{
    int i, j, l;
    for (i = 0; i < m; i++)
        for (j = 0; j < n; j++) {
            double s = 0;
            for (l = 0; l < k; l++) s += (double)A[i * k + l] * (double)B[l * n + j];
            D[i * n + j] = (float)s;
        }
}

#define S(r, c) ((double)A[(r) * n + (c)])

int llo_cv_inv_f32(int n, const float *A, float *D)
{
    int i, j, k;
    if (n == 1) {
        double d = A[0];
        if (d != 0.) { D[0] = (float)(1. / d); return 1; }
        D[0] = 0; return 0;
    }
    if (n == 2) {
        double d = S(0,0) * S(1,1) - S(0,1) * S(1,0);
        if (d != 0.) {
            d = 1. / d;
            double t0 = S(0,0) * d, t1 = S(1,1) * d;
            D[3] = (float)t0; D[0] = (float)t1;
            t0 = -S(0,1) * d; t1 = -S(1,0) * d;
            D[1] = (float)t0; D[2] = (float)t1;
            return 1;
        }
        memset(D, 0, 4 * sizeof(float)); return 0;
    }
    if (n == 3) {
        double d = S(0,0) * (S(1,1) * S(2,2) - S(1,2) * S(2,1)) -
                   S(0,1) * (S(1,0) * S(2,2) - S(1,2) * S(2,0)) +
                   S(0,2) * (S(1,0) * S(2,1) - S(1,1) * S(2,0));
        if (d != 0.) {
            double t[9];
            d = 1. / d;
            t[0] = (S(1,1) * S(2,2) - S(1,2) * S(2,1)) * d;
            t[1] = (S(0,2) * S(2,1) - S(0,1) * S(2,2)) * d;
            t[2] = (S(0,1) * S(1,2) - S(0,2) * S(1,1)) * d;
            t[3] = (S(1,2) * S(2,0) - S(1,0) * S(2,2)) * d;
            t[4] = (S(0,0) * S(2,2) - S(0,2) * S(2,0)) * d;
            t[5] = (S(0,2) * S(1,0) - S(0,0) * S(1,2)) * d;
            t[6] = (S(1,0) * S(2,1) - S(1,1) * S(2,0)) * d;
            t[7] = (S(0,1) * S(2,0) - S(0,0) * S(2,1)) * d;
            t[8] = (S(0,0) * S(1,1) - S(0,1) * S(1,0)) * d;
            for (i = 0; i < 9; i++) D[i] = (float)t[i];
            return 1;
        }
        memset(D, 0, 9 * sizeof(float)); return 0;
    }
    if (n > LLO_MAXN) return 0;
    /* LU with partial pivoting on a copy, right-hand side = identity */
    {
        float a[LLO_MAXN * LLO_MAXN], b[LLO_MAXN * LLO_MAXN];
        const float eps = FLT_EPSILON * 10;
        memcpy(a, A, sizeof(float) * n * n);
        for (i = 0; i < n; i++)
            for (j = 0; j < n; j++) b[i * n + j] = (i == j) ? 1.f : 0.f;
        for (i = 0; i < n; i++) {
            k = i;
            for (j = i + 1; j < n; j++)
                if (fabsf(a[j * n + i]) > fabsf(a[k * n + i])) k = j;
            if (fabsf(a[k * n + i]) < eps) { memset(D, 0, sizeof(float) * n * n); return 0; }
            if (k != i) {
                for (j = i; j < n; j++) { float t = a[i * n + j]; a[i * n + j] = a[k * n + j]; a[k * n + j] = t; }
                for (j = 0; j < n; j++) { float t = b[i * n + j]; b[i * n + j] = b[k * n + j]; b[k * n + j] = t; }
            }
            float d = -1 / a[i * n + i];
            for (j = i + 1; j < n; j++) {
                float alpha = a[j * n + i] * d;
                for (k = i + 1; k < n; k++) a[j * n + k] += alpha * a[i * n + k];
                for (k = 0; k < n; k++) b[j * n + k] += alpha * b[i * n + k];
            }
        }
        for (i = n - 1; i >= 0; i--)
            for (j = 0; j < n; j++) {
                float s = b[i * n + j];
                for (k = i + 1; k < n; k++) s -= a[i * n + k] * b[k * n + j];
                b[i * n + j] = s / a[i * n + i];
            }
        memcpy(D, b, sizeof(float) * n * n);
        return 1;
    }
}
#undef S
