/*
 * llo_odometry.c — ORACLE (test infrastructure): CPU restatement of the
 * featureAssociation scan-to-scan matcher of the reference
 * (/root/reference/LeGO-LOAM/src/featureAssociation.cpp = FA):
 *   TransformToStart                    FA:860-883
 *   findCorrespondingCornerFeatures     FA:1044-1153
 *   findCorrespondingSurfFeatures       FA:1155-1268
 *   calculateTransformationSurf         FA:1270-1377
 *   calculateTransformationCorner       FA:1379-1478
 *   updateTransformation                FA:1666-1695
 * Quirks C1-C8, C14, C20 of SURVEY.md Appendix C are reproduced and marked.
 */
#include "llo.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NEAREST_SQ 25.0f   /* nearestFeatureSearchSqDist, UT:125 */

typedef struct { llo_point *p; int n, cap; } cloud;

static void cloud_set(cloud *c, const llo_point *p, int n)
{
    if (n > c->cap) { c->cap = n + 16; c->p = (llo_point *)realloc(c->p, sizeof(llo_point) * (size_t)c->cap); }
    if (n > 0) memcpy(c->p, p, sizeof(llo_point) * (size_t)n);
    c->n = n;
}
static void cloud_push(cloud *c, llo_point p)
{
    if (c->n + 1 > c->cap) { c->cap = c->cap * 2 + 64; c->p = (llo_point *)realloc(c->p, sizeof(llo_point) * (size_t)c->cap); }
    c->p[c->n++] = p;
}

struct llo_featassoc {
    cloud sharp, flat;                 /* cornerPointsSharp, surfPointsFlat  FA:56-58 */
    cloud cornerLast, surfLast;        /* FA:161-162 */
    cloud treeCorner, treeSurf;        /* what the kd-trees were last built from (C20) */
    llo_kdtree *kdCorner, *kdSurf;     /* FA:166-167 */
    cloud ori, coeffSel;               /* FA:163-164 */
    float *cInd1, *cInd2, *sInd1, *sInd2, *sInd3; int indCap;   /* FA:145-152 (float, C3) */
    float cur[6];                      /* transformCur FA:154 */
    int isDegenerate; float matP[9];   /* FA:179-180 (shared by both solvers, C6) */
    int lastCornerNum, lastSurfNum;    /* FA:142-143 */
};

llo_featassoc *llo_featassoc_create(void) { return (llo_featassoc *)calloc(1, sizeof(llo_featassoc)); }

void llo_featassoc_destroy(llo_featassoc *f)
{
    if (!f) return;
    free(f->sharp.p); free(f->flat.p); free(f->cornerLast.p); free(f->surfLast.p);
    free(f->treeCorner.p); free(f->treeSurf.p); free(f->ori.p); free(f->coeffSel.p);
    free(f->cInd1); free(f->cInd2); free(f->sInd1); free(f->sInd2); free(f->sInd3);
    llo_kdtree_free(f->kdCorner); llo_kdtree_free(f->kdSurf);
    free(f);
}

void llo_featassoc_set_last(llo_featassoc *f, const llo_point *c, int nc, const llo_point *s, int ns, int force)
{
    cloud_set(&f->cornerLast, c, nc);
    cloud_set(&f->surfLast, s, ns);
    f->lastCornerNum = nc; f->lastSurfNum = ns;
    if (force || (nc > 10 && ns > 100)) {                         /* FA:1785 vs FA:1615 */
        cloud_set(&f->treeCorner, c, nc);
        cloud_set(&f->treeSurf, s, ns);
        llo_kdtree_free(f->kdCorner); llo_kdtree_free(f->kdSurf);
        f->kdCorner = llo_kdtree_build(f->treeCorner.p, nc);
        f->kdSurf = llo_kdtree_build(f->treeSurf.p, ns);
    }
}

void llo_featassoc_set_features(llo_featassoc *f, const llo_point *sharp, int nsharp, const llo_point *flat, int nflat)
{
    cloud_set(&f->sharp, sharp, nsharp);
    cloud_set(&f->flat, flat, nflat);
    int need = nsharp > nflat ? nsharp : nflat;
    if (need > f->indCap) {
        int old = f->indCap;
        f->indCap = need + 64;
        size_t b = sizeof(float) * (size_t)f->indCap;
        f->cInd1 = (float *)realloc(f->cInd1, b); f->cInd2 = (float *)realloc(f->cInd2, b);
        f->sInd1 = (float *)realloc(f->sInd1, b); f->sInd2 = (float *)realloc(f->sInd2, b);
        f->sInd3 = (float *)realloc(f->sInd3, b);
        for (int i = old; i < f->indCap; i++)
            f->cInd1[i] = f->cInd2[i] = f->sInd1[i] = f->sInd2[i] = f->sInd3[i] = -1.f;
    }
}

void llo_featassoc_set_transform(llo_featassoc *f, const float c[6]) { memcpy(f->cur, c, sizeof f->cur); }
void llo_featassoc_get_transform(const llo_featassoc *f, float c[6]) { memcpy(c, f->cur, sizeof f->cur); }
void llo_featassoc_get_degenerate(const llo_featassoc *f, int *d, float P[9]) { *d = f->isDegenerate; memcpy(P, f->matP, sizeof f->matP); }
void llo_featassoc_clear_correspondences(llo_featassoc *f) { f->ori.n = 0; f->coeffSel.n = 0; }

/* FA:860-883 */
static llo_point transform_to_start(const llo_featassoc *f, llo_point pi)
{
    float s = 10 * (pi.intensity - (int)pi.intensity);
    float rx = s * f->cur[0], ry = s * f->cur[1], rz = s * f->cur[2];
    float tx = s * f->cur[3], ty = s * f->cur[4], tz = s * f->cur[5];

    float x1 = llo_cosf(rz) * (pi.x - tx) + llo_sinf(rz) * (pi.y - ty);
    float y1 = -llo_sinf(rz) * (pi.x - tx) + llo_cosf(rz) * (pi.y - ty);
    float z1 = (pi.z - tz);

    float y2 = llo_cosf(rx) * y1 + llo_sinf(rx) * z1;
    float z2 = -llo_sinf(rx) * y1 + llo_cosf(rx) * z1;

    llo_point po;
    po.x = llo_cosf(ry) * x1 - llo_sinf(ry) * z2;
    po.y = y2;
    po.z = llo_sinf(ry) * x1 + llo_cosf(ry) * z2;
    po.intensity = pi.intensity;
    return po;
}

static inline float sqdist(const llo_point *a, const llo_point *b)
{
    return (a->x - b->x) * (a->x - b->x) + (a->y - b->y) * (a->y - b->y) + (a->z - b->z) * (a->z - b->z);
}

/* FA:1044-1153 */
void llo_featassoc_findCorrespondingCornerFeatures(llo_featassoc *f, int iterCount)
{
    const int nSharp = f->sharp.n;
    const llo_point *last = f->cornerLast.p;
    for (int i = 0; i < nSharp; i++) {
        llo_point pointSel = transform_to_start(f, f->sharp.p[i]);

        if (iterCount % 5 == 0) {                                   /* C4 */
            int ind[1] = { -1 }; float sq[1] = { 0 };
            llo_kdtree_knn(f->kdCorner, &pointSel.x, 1, ind, sq);
            int closest = -1, minInd2 = -1;
            if (sq[0] < NEAREST_SQ) {
                closest = ind[0];
                int closestScan = (int)last[closest].intensity;
                float minSq2 = NEAREST_SQ;                           /* one threshold for both directions (C2) */
                /* C1: forward bound is the CURRENT sharp count; clamp to the cloud for memory safety */
                int fwdEnd = nSharp < f->cornerLast.n ? nSharp : f->cornerLast.n;
                for (int j = closest + 1; j < fwdEnd; j++) {
                    if ((int)last[j].intensity > closestScan + 2.5) break;
                    float d = sqdist(&last[j], &pointSel);
                    if ((int)last[j].intensity > closestScan && d < minSq2) { minSq2 = d; minInd2 = j; }
                }
                for (int j = closest - 1; j >= 0; j--) {
                    if ((int)last[j].intensity < closestScan - 2.5) break;
                    float d = sqdist(&last[j], &pointSel);
                    if ((int)last[j].intensity < closestScan && d < minSq2) { minSq2 = d; minInd2 = j; }
                }
            }
            f->cInd1[i] = (float)closest;
            f->cInd2[i] = (float)minInd2;
        }

        if (f->cInd2[i] >= 0) {
            llo_point t1 = last[(int)f->cInd1[i]], t2 = last[(int)f->cInd2[i]];
            float x0 = pointSel.x, y0 = pointSel.y, z0 = pointSel.z;
            float x1 = t1.x, y1 = t1.y, z1 = t1.z, x2 = t2.x, y2 = t2.y, z2 = t2.z;

            float m11 = ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1));
            float m22 = ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1));
            float m33 = ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1));
            float a012 = sqrtf(m11 * m11 + m22 * m22 + m33 * m33);
            float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
            float la = ((y1 - y2) * m11 + (z1 - z2) * m22) / a012 / l12;
            float lb = -((x1 - x2) * m11 - (z1 - z2) * m33) / a012 / l12;
            float lc = -((x1 - x2) * m22 + (y1 - y2) * m33) / a012 / l12;
            float ld2 = a012 / l12;

            float s = 1;
            if (iterCount >= 5) s = (float)(1 - 1.8 * fabsf(ld2));
            if (s > 0.1 && ld2 != 0) {
                llo_point coeff = { s * la, s * lb, s * lc, s * ld2 };
                cloud_push(&f->ori, f->sharp.p[i]);
                cloud_push(&f->coeffSel, coeff);
            }
        }
    }
}

/* FA:1155-1268 */
void llo_featassoc_findCorrespondingSurfFeatures(llo_featassoc *f, int iterCount)
{
    const int nFlat = f->flat.n;
    const llo_point *last = f->surfLast.p;
    for (int i = 0; i < nFlat; i++) {
        llo_point pointSel = transform_to_start(f, f->flat.p[i]);

        if (iterCount % 5 == 0) {
            int ind[1] = { -1 }; float sq[1] = { 0 };
            llo_kdtree_knn(f->kdSurf, &pointSel.x, 1, ind, sq);
            int closest = -1, minInd2 = -1, minInd3 = -1;
            if (sq[0] < NEAREST_SQ) {
                closest = ind[0];
                int closestScan = (int)last[closest].intensity;
                float minSq2 = NEAREST_SQ, minSq3 = NEAREST_SQ;
                int fwdEnd = nFlat < f->surfLast.n ? nFlat : f->surfLast.n;     /* C1 */
                for (int j = closest + 1; j < fwdEnd; j++) {
                    if ((int)last[j].intensity > closestScan + 2.5) break;
                    float d = sqdist(&last[j], &pointSel);
                    if ((int)last[j].intensity <= closestScan) {
                        if (d < minSq2) { minSq2 = d; minInd2 = j; }
                    } else {
                        if (d < minSq3) { minSq3 = d; minInd3 = j; }
                    }
                }
                for (int j = closest - 1; j >= 0; j--) {
                    if ((int)last[j].intensity < closestScan - 2.5) break;
                    float d = sqdist(&last[j], &pointSel);
                    if ((int)last[j].intensity >= closestScan) {
                        if (d < minSq2) { minSq2 = d; minInd2 = j; }
                    } else {
                        if (d < minSq3) { minSq3 = d; minInd3 = j; }
                    }
                }
            }
            f->sInd1[i] = (float)closest;
            f->sInd2[i] = (float)minInd2;
            f->sInd3[i] = (float)minInd3;
        }

        if (f->sInd2[i] >= 0 && f->sInd3[i] >= 0) {
            llo_point t1 = last[(int)f->sInd1[i]], t2 = last[(int)f->sInd2[i]], t3 = last[(int)f->sInd3[i]];

            float pa = (t2.y - t1.y) * (t3.z - t1.z) - (t3.y - t1.y) * (t2.z - t1.z);
            float pb = (t2.z - t1.z) * (t3.x - t1.x) - (t3.z - t1.z) * (t2.x - t1.x);
            float pc = (t2.x - t1.x) * (t3.y - t1.y) - (t3.x - t1.x) * (t2.y - t1.y);
            float pd = -(pa * t1.x + pb * t1.y + pc * t1.z);
            float ps = sqrtf(pa * pa + pb * pb + pc * pc);
            pa /= ps; pb /= ps; pc /= ps; pd /= ps;

            float pd2 = pa * pointSel.x + pb * pointSel.y + pc * pointSel.z + pd;

            float s = 1;
            if (iterCount >= 5)
                s = (float)(1 - 1.8 * fabsf(pd2) /
                            sqrtf(sqrtf(pointSel.x * pointSel.x + pointSel.y * pointSel.y + pointSel.z * pointSel.z)));
            if (s > 0.1 && pd2 != 0) {
                llo_point coeff = { s * pa, s * pb, s * pc, s * pd2 };
                cloud_push(&f->ori, f->flat.p[i]);
                cloud_push(&f->coeffSel, coeff);
            }
        }
    }
}

/* shared tail of both 3-DoF solvers: FA:1324-1365 / FA:1425-1466 */
static void solve3(llo_featassoc *f, int iterCount, int N, const float *A, const float *B, float X[3])
{
    float *At = (float *)malloc(sizeof(float) * 3 * (size_t)N);
    float AtA[9], AtB[3];
    for (int i = 0; i < N; i++)
        for (int j = 0; j < 3; j++) At[(size_t)j * N + i] = A[3 * i + j];
    llo_cv_gemm_f32(3, N, 3, At, A, AtA);
    llo_cv_gemm_f32(3, N, 1, At, B, AtB);
    free(At);
    llo_cv_solve_qr_f32(3, 3, AtA, AtB, X);

    if (iterCount == 0) {
        float tmp[9], E[3], V[9], V2[9], Vinv[9];
        memcpy(tmp, AtA, sizeof tmp);
        llo_cv_eigen_f32(3, tmp, E, V);
        memcpy(V2, V, sizeof V2);
        f->isDegenerate = 0;
        for (int i = 2; i >= 0; i--) {                              /* C7, threshold 10 */
            if (E[i] < 10.f) {
                for (int j = 0; j < 3; j++) V2[3 * i + j] = 0;
                f->isDegenerate = 1;
            } else break;
        }
        llo_cv_inv_f32(3, V, Vinv);
        llo_cv_gemm_f32(3, 3, 3, Vinv, V2, f->matP);
    }
    if (f->isDegenerate) {
        float X2[3] = { X[0], X[1], X[2] };
        llo_cv_gemm_f32(3, 3, 1, f->matP, X2, X);
    }
}

static double fa_rad2deg(double r) { return r * 180.0 / M_PI; }     /* FA:1034-1037 */

/* FA:1270-1377. returns 0 (false) when converged (C8) */
int llo_featassoc_calculateTransformationSurf(llo_featassoc *f, int iterCount)
{
    int N = f->ori.n;
    float *A = (float *)malloc(sizeof(float) * 3 * (size_t)(N > 0 ? N : 1));
    float *B = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));

    float srx = llo_sinf(f->cur[0]), crx = llo_cosf(f->cur[0]);
    float sry = llo_sinf(f->cur[1]), cry = llo_cosf(f->cur[1]);
    float srz = llo_sinf(f->cur[2]), crz = llo_cosf(f->cur[2]);
    float tx = f->cur[3], ty = f->cur[4], tz = f->cur[5];

    float a1 = crx * sry * srz, a2 = crx * crz * sry, a3 = srx * sry, a4 = tx * a1 - ty * a2 - tz * a3;
    float a5 = srx * srz, a6 = crz * srx, a7 = ty * a6 - tz * crx - tx * a5;
    float a8 = crx * cry * srz, a9 = crx * cry * crz, a10 = cry * srx, a11 = tz * a10 + ty * a9 - tx * a8;

    float b1 = -crz * sry - cry * srx * srz, b2 = cry * crz * srx - sry * srz;
    float b5 = cry * crz - srx * sry * srz, b6 = cry * srz + crz * srx * sry;

    float c1 = -b6, c2 = b5, c3 = tx * b6 - ty * b5, c4 = -crx * crz, c5 = crx * srz, c6 = ty * c5 + tx * -c4;
    float c7 = b2, c8 = -b1, c9 = tx * -b2 - ty * -b1;

    for (int i = 0; i < N; i++) {
        llo_point p = f->ori.p[i], c = f->coeffSel.p[i];
        float arx = (-a1 * p.x + a2 * p.y + a3 * p.z + a4) * c.x
                  + (a5 * p.x - a6 * p.y + crx * p.z + a7) * c.y
                  + (a8 * p.x - a9 * p.y - a10 * p.z + a11) * c.z;
        float arz = (c1 * p.x + c2 * p.y + c3) * c.x
                  + (c4 * p.x - c5 * p.y + c6) * c.y
                  + (c7 * p.x + c8 * p.y + c9) * c.z;
        float aty = -b6 * c.x + c4 * c.y + b2 * c.z;
        float d2 = c.intensity;
        A[3 * i] = arx; A[3 * i + 1] = arz; A[3 * i + 2] = aty;
        B[i] = (float)(-0.05 * d2);                                 /* double literal */
    }
    float X[3];
    solve3(f, iterCount, N, A, B, X);
    free(A); free(B);

    f->cur[0] += X[0]; f->cur[2] += X[1]; f->cur[4] += X[2];
    for (int i = 0; i < 6; i++) if (isnan(f->cur[i])) f->cur[i] = 0;  /* C14 */

    double r0 = fa_rad2deg(X[0]), r1 = fa_rad2deg(X[1]);
    double t0 = (double)(X[2] * 100);
    float deltaR = (float)sqrt(r0 * r0 + r1 * r1);
    float deltaT = (float)sqrt(t0 * t0);
    if (deltaR < 0.1 && deltaT < 0.1) return 0;
    return 1;
}

/* FA:1379-1478. returns 0 (false) when converged (C8) */
int llo_featassoc_calculateTransformationCorner(llo_featassoc *f, int iterCount)
{
    int N = f->ori.n;
    float *A = (float *)malloc(sizeof(float) * 3 * (size_t)(N > 0 ? N : 1));
    float *B = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));

    float srx = llo_sinf(f->cur[0]), crx = llo_cosf(f->cur[0]);
    float sry = llo_sinf(f->cur[1]), cry = llo_cosf(f->cur[1]);
    float srz = llo_sinf(f->cur[2]), crz = llo_cosf(f->cur[2]);
    float tx = f->cur[3], ty = f->cur[4], tz = f->cur[5];

    float b1 = -crz * sry - cry * srx * srz, b2 = cry * crz * srx - sry * srz, b3 = crx * cry,
          b4 = tx * -b1 + ty * -b2 + tz * b3;
    float b5 = cry * crz - srx * sry * srz, b6 = cry * srz + crz * srx * sry, b7 = crx * sry,
          b8 = tz * b7 - ty * b6 - tx * b5;
    float c5 = crx * srz;

    for (int i = 0; i < N; i++) {
        llo_point p = f->ori.p[i], c = f->coeffSel.p[i];
        float ary = (b1 * p.x + b2 * p.y - b3 * p.z + b4) * c.x
                  + (b5 * p.x + b6 * p.y - b7 * p.z + b8) * c.z;
        float atx = -b5 * c.x + c5 * c.y + b1 * c.z;
        float atz = b7 * c.x - srx * c.y - b3 * c.z;
        float d2 = c.intensity;
        A[3 * i] = ary; A[3 * i + 1] = atx; A[3 * i + 2] = atz;
        B[i] = (float)(-0.05 * d2);
    }
    float X[3];
    solve3(f, iterCount, N, A, B, X);
    free(A); free(B);

    f->cur[1] += X[0]; f->cur[3] += X[1]; f->cur[5] += X[2];
    for (int i = 0; i < 6; i++) if (isnan(f->cur[i])) f->cur[i] = 0;

    double r0 = fa_rad2deg(X[0]);
    double t0 = (double)(X[1] * 100), t1 = (double)(X[2] * 100);
    float deltaR = (float)sqrt(r0 * r0);
    float deltaT = (float)sqrt(t0 * t0 + t1 * t1);
    if (deltaR < 0.1 && deltaT < 0.1) return 0;
    return 1;
}

/* FA:1666-1695 */
int llo_featassoc_updateTransformation(llo_featassoc *f)
{
    int it1 = 0, it2 = 0;
    if (f->lastCornerNum < 10 || f->lastSurfNum < 100) return 0;
    for (int it = 0; it < 25; it++) {
        llo_featassoc_clear_correspondences(f);
        llo_featassoc_findCorrespondingSurfFeatures(f, it);
        it1++;
        if (f->ori.n < 10) continue;
        if (llo_featassoc_calculateTransformationSurf(f, it) == 0) break;
    }
    for (int it = 0; it < 25; it++) {
        llo_featassoc_clear_correspondences(f);
        llo_featassoc_findCorrespondingCornerFeatures(f, it);
        it2++;
        if (f->ori.n < 10) continue;
        if (llo_featassoc_calculateTransformationCorner(f, it) == 0) break;
    }
    return it1 | (it2 << 16);
}

int llo_featassoc_get_correspondences(const llo_featassoc *f, llo_point *ori, llo_point *coeff, int cap)
{
    int n = f->ori.n < cap ? f->ori.n : cap;
    if (n > 0) { memcpy(ori, f->ori.p, sizeof(llo_point) * (size_t)n); memcpy(coeff, f->coeffSel.p, sizeof(llo_point) * (size_t)n); }
    return f->ori.n;
}

int llo_featassoc_get_search_ind(const llo_featassoc *f, int which, float *i1, float *i2, float *i3, int cap)
{
    int n = which == 0 ? f->sharp.n : f->flat.n;
    int c = n < cap ? n : cap;
    if (c > 0) {
        memcpy(i1, which == 0 ? f->cInd1 : f->sInd1, sizeof(float) * (size_t)c);
        memcpy(i2, which == 0 ? f->cInd2 : f->sInd2, sizeof(float) * (size_t)c);
        if (which == 1 && i3) memcpy(i3, f->sInd3, sizeof(float) * (size_t)c);
    }
    return n;
}
