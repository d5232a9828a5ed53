/*
 * llo_loop.c — CPU ORACLE (test infrastructure) for the loop-closure alignment of mapOptimization:
 * pcl::IterativeClosestPoint<PointXYZI, PointXYZI>::align as performLoopClosure configures it (MO:892-902:
 * max correspondence distance 100, 100 iterations, transformation epsilon 1e-6, Euclidean fitness epsilon 1e-6,
 * no RANSAC) and getFitnessScore (MO:904).
 *
 * PARITY UNPINNED: PCL and Eigen exist nowhere offline and the reference has no fixtures for this path.  Restated
 * from the published algorithm of PCL 1.8 (the version the reference's README names):
 *   registration/impl/icp.hpp computeTransformation + transformCloud, correspondence_estimation.hpp
 *   determineCorrespondences (exact 1-NN per source point, kept when d^2 <= max^2), default_convergence_criteria.hpp
 *   hasConverged (iterations / transformation epsilon / absolute 1e-12 and relative MSE), transformation_estimation_svd.hpp
 *   -> pcl::umeyama (= Eigen::umeyama without scaling), registration.hpp getFitnessScore.
 * Known deviations from a real PCL build: Eigen sums the means and the 3x3 covariance in float with its packet order
 * (not reproducible without Eigen); here they are accumulated in fp64 and the 3x3 SVD is a one-sided Jacobi in fp64.
 * The rotation of the Umeyama solution is the polar factor of the covariance and does not depend on the SVD algorithm;
 * the float-sum noise of a real PCL run is ~1e-4 m per step (thousands of ~50 m coordinates), i.e. below the 1e-3 m
 * step at which its own convergence test stops.  Equal nearest-neighbour distances: smaller index (FLANN's order is
 * tree-dependent).
 */
#include "llo.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* exact 1-NN, squared L2 as flann::L2_Simple accumulates it (float, x then y then z); ties: smaller index */
static void nn1(const llo_kdtree *tree, const llo_point *q, int *idx, float *d2)
{
    float qq[3] = { q->x, q->y, q->z };
    llo_kdtree_knn(tree, qq, 1, idx, d2);
}

/* V (3x3, columns = right singular vectors), s[3], U (3x3) of A = U diag(s) V^T; one-sided Jacobi in fp64,
 * singular values sorted descending */
static void svd3(const double A[9], double U[9], double s[3], double V[9])
{
    double B[9]; memcpy(B, A, sizeof(B));
    for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double a = 0, b = 0, c = 0;
                for (int k = 0; k < 3; k++) { a += B[3 * k + p] * B[3 * k + p]; b += B[3 * k + q] * B[3 * k + q]; c += B[3 * k + p] * B[3 * k + q]; }
                off = fmax(off, fabs(c) / (sqrt(a * b) + DBL_MIN));
                if (fabs(c) <= 1e-300) continue;
                const double zeta = (b - a) / (2.0 * c);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int k = 0; k < 3; k++) {
                    const double bp = B[3 * k + p], bq = B[3 * k + q];
                    B[3 * k + p] = cs * bp - sn * bq; B[3 * k + q] = sn * bp + cs * bq;
                    const double vp = V[3 * k + p], vq = V[3 * k + q];
                    V[3 * k + p] = cs * vp - sn * vq; V[3 * k + q] = sn * vp + cs * vq;
                }
            }
        if (off < 1e-15) break;
    }
    int ord[3] = { 0, 1, 2 };
    double n[3];
    for (int j = 0; j < 3; j++) n[j] = sqrt(B[j] * B[j] + B[3 + j] * B[3 + j] + B[6 + j] * B[6 + j]);
    for (int i = 0; i < 2; i++) for (int j = i + 1; j < 3; j++) if (n[ord[j]] > n[ord[i]]) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
    double Vs[9], Us[9];
    for (int j = 0; j < 3; j++) {
        const int o = ord[j];
        s[j] = n[o];
        for (int k = 0; k < 3; k++) { Vs[3 * k + j] = V[3 * k + o]; Us[3 * k + j] = n[o] > 0 ? B[3 * k + o] / n[o] : 0.0; }
    }
    /* a vanishing singular value leaves its left vector undefined: complete U to an orthonormal basis */
    if (s[2] <= 1e-12 * s[0]) {
        if (s[1] <= 1e-12 * s[0]) {         /* rank <= 1: any basis orthogonal to u0 */
            double a[3] = { Us[0], Us[3], Us[6] };
            if (s[0] <= 0) { a[0] = 1; a[1] = 0; a[2] = 0; Us[0] = 1; Us[3] = 0; Us[6] = 0; }
            double e[3] = { 0, 0, 0 }; e[fabs(a[0]) < 0.9 ? 0 : 1] = 1.0;
            double d = e[0] * a[0] + e[1] * a[1] + e[2] * a[2];
            double u1[3] = { e[0] - d * a[0], e[1] - d * a[1], e[2] - d * a[2] };
            double l = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
            for (int k = 0; k < 3; k++) Us[3 * k + 1] = u1[k] / l;
        }
        const double a[3] = { Us[0], Us[3], Us[6] }, b[3] = { Us[1], Us[4], Us[7] };
        Us[2] = a[1] * b[2] - a[2] * b[1]; Us[5] = a[2] * b[0] - a[0] * b[2]; Us[8] = a[0] * b[1] - a[1] * b[0];
    }
    memcpy(U, Us, sizeof(Us)); memcpy(V, Vs, sizeof(Vs));
}

static double det3(const double M[9])
{
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

/* pcl::umeyama(src, dst, with_scaling = false) from the 17 sums of the correspondences: n, sum p (source), sum q
 * (target), sum q p^T.  Rt: row-major 4x4 float */
void llo_umeyama_from_sums(double n, const double sp[3], const double sq[3], const double sqp[9], float Rt[16])
{
    double pm[3], qm[3], sigma[9];
    for (int a = 0; a < 3; a++) { pm[a] = sp[a] / n; qm[a] = sq[a] / n; }
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) sigma[3 * r + c] = sqp[3 * r + c] / n - qm[r] * pm[c];
    double U[9], s[3], V[9], S[3] = { 1, 1, 1 };
    svd3(sigma, U, s, V);
    if (det3(sigma) < 0) S[2] = -1;
    int rank = 0;
    for (int i = 0; i < 3; i++) if (!(fabs(s[i]) <= fabs(s[0]) * 1e-12)) rank++;   /* isMuchSmallerThan(d_i, d_0) */
    if (rank == 2) S[2] = (det3(U) * det3(V) > 0) ? 1 : -1;
    double R[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
        double v = 0; for (int k = 0; k < 3; k++) v += U[3 * r + k] * S[k] * V[3 * c + k];
        R[3 * r + c] = v;
    }
    for (int i = 0; i < 16; i++) Rt[i] = (i % 5 == 0) ? 1.f : 0.f;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) Rt[4 * r + c] = (float)R[3 * r + c];
        Rt[4 * r + 3] = (float)(qm[r] - (R[3 * r] * pm[0] + R[3 * r + 1] * pm[1] + R[3 * r + 2] * pm[2]));
    }
}

/* IterativeClosestPoint::transformCloud: pt_t = tr * (x, y, z, 1) in float, columns added left to right */
static void transform_cloud(llo_point *p, int n, const float T[16])
{
    for (int i = 0; i < n; i++) {
        const float x = p[i].x, y = p[i].y, z = p[i].z;
        p[i].x = ((T[0] * x + T[1] * y) + T[2] * z) + T[3];
        p[i].y = ((T[4] * x + T[5] * y) + T[6] * z) + T[7];
        p[i].z = ((T[8] * x + T[9] * y) + T[10] * z) + T[11];
    }
}

static void matmul4(const float A[16], const float B[16], float C[16])
{
    float out[16];
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++)
        out[4 * r + c] = ((A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c]) + A[4 * r + 2] * B[8 + c]) + A[4 * r + 3] * B[12 + c];
    memcpy(C, out, sizeof(out));
}

/* one ICP iteration on the CURRENT (already transformed) source cloud: correspondences, the 17 sums, the mean squared
 * correspondence distance.  Returns the number of correspondences. */
int llo_icp_correspondence_sums(const llo_point *cur, int ns, const llo_kdtree *tree, const llo_point *tgt, double max_d2,
                                double sp[3], double sq[3], double sqp[9], double *mse, int *nn_idx /* may be NULL */)
{
    int cnt = 0; double sum_d = 0;
    memset(sp, 0, 3 * sizeof(double)); memset(sq, 0, 3 * sizeof(double)); memset(sqp, 0, 9 * sizeof(double));
    for (int i = 0; i < ns; i++) {
        int j; float d2;
        nn1(tree, &cur[i], &j, &d2);
        if (nn_idx) nn_idx[i] = ((double)d2 > max_d2) ? -1 : j;
        if ((double)d2 > max_d2) continue;
        const double p[3] = { cur[i].x, cur[i].y, cur[i].z }, q[3] = { tgt[j].x, tgt[j].y, tgt[j].z };
        for (int a = 0; a < 3; a++) { sp[a] += p[a]; sq[a] += q[a]; for (int b = 0; b < 3; b++) sqp[3 * a + b] += q[a] * p[b]; }
        sum_d += (double)d2; cnt++;
    }
    *mse = cnt > 0 ? sum_d / (double)cnt : 0.0;
    return cnt;
}

/* result->T: final_transformation_ (row-major); state: 0 not converged (too few correspondences), 1 iterations,
 * 2 transform, 3 abs MSE, 4 rel MSE (DefaultConvergenceCriteria::ConvergenceState) */
int llo_icp_align(const llo_point *src, int ns, const llo_point *tgt, int nt, int max_iterations, double max_corr_dist,
                  double transformation_epsilon, double euclidean_fitness_epsilon, float T_final[16], int *converged,
                  int *iterations, int *state, double *fitness)
{
    for (int i = 0; i < 16; i++) T_final[i] = (i % 5 == 0) ? 1.f : 0.f;
    *converged = 0; *iterations = 0; *state = 0; *fitness = DBL_MAX;
    if (ns <= 0 || nt <= 0) return 0;
    llo_kdtree *tree = llo_kdtree_build(tgt, nt);
    llo_point *cur = (llo_point *)malloc(sizeof(llo_point) * (size_t)ns);
    memcpy(cur, src, sizeof(llo_point) * (size_t)ns);
    const double max_d2 = max_corr_dist * max_corr_dist;
    const double rot_thr = 1.0 - transformation_epsilon, trans_thr = transformation_epsilon;
    double prev_mse = DBL_MAX;
    int it = 0, conv = 0, st = 0;
    for (;;) {
        double sp[3], sq[3], sqp[9], mse;
        const int cnt = llo_icp_correspondence_sums(cur, ns, tree, tgt, max_d2, sp, sq, sqp, &mse, NULL);
        if (cnt < 3) { conv = 0; st = 0; break; }           /* min_number_correspondences_ */
        float Tr[16];
        llo_umeyama_from_sums((double)cnt, sp, sq, sqp, Tr);
        transform_cloud(cur, ns, Tr);
        matmul4(Tr, T_final, T_final);
        it++;
        /* DefaultConvergenceCriteria::hasConverged */
        if (it >= max_iterations) { conv = 1; st = 1; break; }
        const double cos_angle = 0.5 * ((double)Tr[0] + (double)Tr[5] + (double)Tr[10] - 1.0);
        const double tsq = (double)Tr[3] * Tr[3] + (double)Tr[7] * Tr[7] + (double)Tr[11] * Tr[11];
        if (cos_angle >= rot_thr && tsq <= trans_thr) { conv = 1; st = 2; break; }
        if (fabs(mse - prev_mse) < 1e-12) { conv = 1; st = 3; break; }
        if (fabs(mse - prev_mse) / prev_mse < euclidean_fitness_epsilon) { conv = 1; st = 4; break; }
        prev_mse = mse;
    }
    /* getFitnessScore: the ORIGINAL source through final_transformation_, mean squared 1-NN distance */
    memcpy(cur, src, sizeof(llo_point) * (size_t)ns);
    transform_cloud(cur, ns, T_final);
    double fs = 0; int nr = 0;
    for (int i = 0; i < ns; i++) { int j; float d2; nn1(tree, &cur[i], &j, &d2); fs += (double)d2; nr++; }
    *fitness = nr > 0 ? fs / nr : DBL_MAX;
    *converged = conv; *iterations = it; *state = st;
    free(cur); llo_kdtree_free(tree);
    return 0;
}
