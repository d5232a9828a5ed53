// s2m_dev.cuh — device functions of the scan-to-map iteration shared by the persistent single-registration
// kernel (s2m.cu) and the batched multi-registration kernels (batch.cu): the per-query fits of
// cornerOptimization (MO:1102-1170) / surfOptimization (MO:1184-1223), the Jacobian row of LMOptimization
// (MO:1252-1271) and the LM step with the degeneracy analysis (MO:1273-1326).
#pragma once
#include "s2m.cuh"
#include "linalg.cuh"
#include "glibc_sincosf.cuh"

namespace llb {
// cornerOptimization body for one query (MO:1102-1170).  Returns true when the row is accepted.
__device__ __forceinline__ bool corner_fit(const float (&nx)[5], const float (&ny)[5], const float (&nz)[5],
                                           float x0, float y0, float z0, float4 &coeff)
{
    float cx = 0, cy = 0, cz = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) { cx += nx[j]; cy += ny[j]; cz += nz[j]; }
    cx /= 5; cy /= 5; cz /= 5;

    float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        float ax = nx[j] - cx, ay = ny[j] - cy, az = nz[j] - cz;
        a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
        a22 += ay * ay; a23 += ay * az;
        a33 += az * az;
    }
    a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;

    float D1[3], V1[9];
    cv_eigen3(a11, a12, a13, a22, a23, a33, D1, V1);
    if (!(D1[0] > 3 * D1[1])) return false;

    float x1 = (float)((double)cx + 0.1 * (double)V1[0]);
    float y1 = (float)((double)cy + 0.1 * (double)V1[1]);
    float z1 = (float)((double)cz + 0.1 * (double)V1[2]);
    float x2 = (float)((double)cx - 0.1 * (double)V1[0]);
    float y2 = (float)((double)cy - 0.1 * (double)V1[1]);
    float z2 = (float)((double)cz - 0.1 * (double)V1[2]);

    float m11 = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1);
    float m22 = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1);
    float m33 = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
    float a012 = sqrtf(m11 * m11 + m22 * m22 + m33 * m33);
    float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
    float la = ((y1 - y2) * m11 + (z1 - z2) * m22) / a012 / l12;
    float lb = -((x1 - x2) * m11 - (z1 - z2) * m33) / a012 / l12;
    float lc = -((x1 - x2) * m22 + (y1 - y2) * m33) / a012 / l12;
    float ld2 = a012 / l12;

    float s = (float)(1.0 - 0.9 * (double)fabsf(ld2));
    coeff = make_float4(s * la, s * lb, s * lc, s * ld2);
    return (double)s > 0.1;
}

// surfOptimization body for one query (MO:1184-1223)
__device__ __forceinline__ bool surf_fit(const float (&nx)[5], const float (&ny)[5], const float (&nz)[5],
                                         float x0, float y0, float z0, float4 &coeff)
{
    float A0[15], B0[5] = { -1.f, -1.f, -1.f, -1.f, -1.f }, X0[3];
#pragma unroll
    for (int j = 0; j < 5; j++) { A0[3 * j] = nx[j]; A0[3 * j + 1] = ny[j]; A0[3 * j + 2] = nz[j]; }
    cv_solve_qr<5, 3>(A0, B0, X0);

    float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1;
    float ps = sqrtf(pa * pa + pb * pb + pc * pc);
    pa /= ps; pb /= ps; pc /= ps; pd /= ps;

    bool planeValid = true;
#pragma unroll
    for (int j = 0; j < 5; j++)
        if ((double)fabsf(pa * nx[j] + pb * ny[j] + pc * nz[j] + pd) > 0.2) planeValid = false;
    if (!planeValid) return false;

    float pd2 = pa * x0 + pb * y0 + pc * z0 + pd;
    float s = (float)(1.0 - 0.9 * (double)fabsf(pd2) / (double)sqrtf(sqrtf(x0 * x0 + y0 * y0 + z0 * z0)));
    coeff = make_float4(s * pa, s * pb, s * pc, s * pd2);
    return (double)s > 0.1;
}

// sin / cos of a float pose angle as the reference build gets them (MO:498-506, MO:1241-1246: sin(float) is glibc's
// sinf, which is NOT always the correctly rounded value): glibc's operation sequence restated (glibc_sincosf.cuh), so the
// poses are the reference's bit for bit.
__device__ __forceinline__ float ref_sinf(float x) { return glibcm::sinf_(x); }
__device__ __forceinline__ float ref_cosf(float x) { return glibcm::cosf_(x); }
// entry k of S2mState::cs (cRoll sRoll cPitch sPitch cYaw sYaw) for the pose T
__device__ __forceinline__ float pose_trig(const float *T, int k) { return (k & 1) ? ref_sinf(T[k >> 1]) : ref_cosf(T[k >> 1]); }

__device__ inline void update_sincos(S2mState *st)
{
#pragma unroll
    for (int k = 0; k < 6; k++) st->cs[k] = pose_trig(st->T, k);
}

// LMOptimization tail MO:1273-1326 on the 28 reduced sums (one thread)
// Certificate that every eigenvalue of the symmetric matrix A (6x6, float) exceeds `bound`:
// Cholesky of (A - bound*I) in fp64 succeeds iff that matrix is positive definite.
__device__ inline bool all_eigenvalues_above(const float *A, double bound)
{
    double L[36];
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = (double)A[j * 6 + j] - bound;
#pragma unroll
        for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k];
        if (!(d > 0.0)) return false;
        const double r = sqrt(d);
        L[j * 6 + j] = r;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            double s = (double)A[i * 6 + j];
#pragma unroll
            for (int k = 0; k < j; k++) s -= L[i * 6 + k] * L[j * 6 + k];
            L[i * 6 + j] = s / r;
        }
    }
    return true;
}

// iteration-0 degeneracy analysis MO:1278-1299 on AtA: cv::eigen, zero the rows of the small
// eigenvalues, matP = V^-1 * V2.  ~80 us for one thread (Jacobi in local memory): only run when the
// cheap certificate below cannot rule degeneracy out, or when matP is asked for.
__device__ inline void degeneracy_full(S2mState *st, const float *AtA, float thresh)
{
    float A[36], E[6], V[36], V2[36], Vinv[36];
    for (int i = 0; i < 36; i++) A[i] = AtA[i];
    cv_eigen<6>(A, E, V);
    for (int i = 0; i < 36; i++) V2[i] = V[i];
    int deg = 0;
    for (int i = 5; i >= 0; i--) {
        if (E[i] < thresh) {
            for (int j = 0; j < 6; j++) V2[i * 6 + j] = 0.f;
            deg = 1;
        } else break;
    }
    st->is_degenerate = deg;
    cv_inv_lu<6>(V, Vinv);
    cv_gemm<6, 6, 6>(Vinv, V2, st->matP);
    st->matP_valid = 1;
}

__device__ inline void lm_solve(S2mState *st, const double *sum, int iter, const S2mParams &prm, bool do_trig)
{
    const int n_corr = (int)sum[27];
    st->n_corr = n_corr;
    st->iters = iter + 1;
    if (n_corr < prm.min_corr) return;                       // MO:1238: pose untouched, not converged

    float AtA[36], AtB[6], A[36], B[6], X[6];
    int k = 0;
    for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++, k++) { float v = (float)sum[k]; AtA[i * 6 + j] = v; AtA[j * 6 + i] = v; }
    for (int i = 0; i < 6; i++) AtB[i] = (float)sum[21 + i];
    for (int i = 0; i < 36; i++) { A[i] = AtA[i]; st->AtA[i] = AtA[i]; }
    for (int i = 0; i < 6; i++) { B[i] = AtB[i]; st->AtB[i] = AtB[i]; }
    cv_solve_qr<6, 6>(A, B, X);

    if (iter == 0) {
        // isDegenerate <=> the float Jacobi reports an eigenvalue < thresh.  Jacobi's absolute eigenvalue
        // error is O(n eps ||A||) <= ~1e-5 trace(A); if A - (thresh + 1e-4 trace) I is positive definite
        // every Jacobi eigenvalue is >= thresh for sure: not degenerate, and matP is never read
        // (MO:1301), so the eigen-decomposition is deferred until somebody asks for matP.
        double tr = 0.0;
        for (int i = 0; i < 6; i++) { tr += (double)AtA[i * 6 + i]; }
        for (int i = 0; i < 36; i++) st->AtA0[i] = AtA[i];
        if (all_eigenvalues_above(AtA, (double)prm.degeneracy_thresh + 1e-4 * tr + 1.0)) {
            st->is_degenerate = 0;
            st->matP_valid = 0;
        } else {
            degeneracy_full(st, AtA, prm.degeneracy_thresh);
        }
    }
    if (st->is_degenerate) {
        float X2[6];
        for (int i = 0; i < 6; i++) X2[i] = X[i];
        cv_gemm<6, 6, 1>(st->matP, X2, X);
    }
    for (int i = 0; i < 6; i++) { st->T[i] += X[i]; st->X[i] = X[i]; }
    if (do_trig) update_sincos(st);

    double r0 = (double)(X[0] * 57.29578f), r1 = (double)(X[1] * 57.29578f), r2 = (double)(X[2] * 57.29578f);
    double t0 = (double)(X[3] * 100), t1 = (double)(X[4] * 100), t2 = (double)(X[5] * 100);
    float deltaR = (float)sqrt(r0 * r0 + r1 * r1 + r2 * r2);
    float deltaT = (float)sqrt(t0 * t0 + t1 * t1 + t2 * t2);
    if ((double)deltaR < (double)prm.converge_deg && (double)deltaT < (double)prm.converge_cm) st->converged = 1;
}

// index pairs of the 28 accumulated products of v = {arx, ary, arz, cx, cy, cz, b, 1}
__device__ __forceinline__ void pair_of(int k, int &ia, int &ib)
{
    ia = 7; ib = 7;                                          // k == 27: row count
    int c = 0;
    for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++, c++)
            if (c == k) { ia = i; ib = j; }
    if (k >= 21 && k < 27) { ia = k - 21; ib = 6; }
}

// one row of matA / matB in LMOptimization (MO:1252-1271): v = {arx, ary, arz, coeff.x, coeff.y, coeff.z, -coeff.w, 1}
__device__ __forceinline__ void jacobian_row(float crx, float srx, float cry, float sry, float crz, float srz,
                                             float px, float py, float pz, const float4 &coeff, float (&v)[8])
{
    v[0] = (crx * sry * srz * px + crx * crz * sry * py - srx * sry * pz) * coeff.x
         + (-srx * srz * px - crz * srx * py - crx * pz) * coeff.y
         + (crx * cry * srz * px + crx * cry * crz * py - cry * srx * pz) * coeff.z;
    v[1] = ((cry * srx * srz - crz * sry) * px + (sry * srz + cry * crz * srx) * py + crx * cry * pz) * coeff.x
         + ((-cry * crz - srx * sry * srz) * px + (cry * srz - crz * srx * sry) * py - crx * sry * pz) * coeff.z;
    v[2] = ((crz * srx * sry - cry * srz) * px + (-cry * crz - srx * sry * srz) * py) * coeff.x
         + (crx * crz * px - crx * srz * py) * coeff.y
         + ((sry * srz + cry * crz * srx) * px + (crz * sry - cry * srx * srz) * py) * coeff.z;
    v[3] = coeff.x; v[4] = coeff.y; v[5] = coeff.z;
    v[6] = -coeff.w;
    v[7] = 1.f;
}

// pointAssociateToMap MO:513-527 with the six cached sin/cos of MO:498-506
__device__ __forceinline__ void associate_to_map(float crx, float srx, float cry, float sry, float crz, float srz,
                                                 float tX, float tY, float tZ, const float4 &po,
                                                 float &sx, float &sy, float &sz)
{
    const float x1 = crz * po.x - srz * po.y;
    const float y1 = srz * po.x + crz * po.y;
    const float z1 = po.z;
    const float y2 = crx * y1 - srx * z1;
    const float z2 = srx * y1 + crx * z1;
    sx = cry * x1 + sry * z2 + tX;
    sy = y2 + tY;
    sz = -sry * x1 + cry * z2 + tZ;
}

}  // namespace llb
