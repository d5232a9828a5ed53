/*
 * llo_mapping.c — ORACLE (test infrastructure): CPU restatement of the
 * mapOptimization scan-to-map hot path of the reference
 * (/root/reference/LeGO-LOAM/src/mapOptmization.cpp = MO):
 *   map voxel DS                MO:1057-1064
 *   downsampleCurrentScan       MO:1067-1091
 *   pointAssociateToMap         MO:498-527
 *   cornerOptimization          MO:1093-1174
 *   surfOptimization            MO:1176-1227
 *   LMOptimization              MO:1229-1327
 *   scan2MapOptimization        MO:1329-1350
 *   transformUpdate (no IMU)    MO:463-496
 * Float / double promotion follows SURVEY.md Appendix B exactly; every float
 * expression keeps the reference's association order so that the result is
 * bit-identical to the reference compiled without FMA contraction.
 */
#include "llo.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { llo_point *p; int n, cap; } cloud;

static void cloud_reserve(cloud *c, int n)
{
    if (n > c->cap) {
        c->cap = n + n / 2 + 16;
        c->p = (llo_point *)realloc(c->p, sizeof(llo_point) * (size_t)c->cap);
    }
}
static void cloud_set(cloud *c, const llo_point *p, int n)
{
    cloud_reserve(c, n);
    if (n > 0) memcpy(c->p, p, sizeof(llo_point) * (size_t)n);
    c->n = n;
}
static void cloud_push(cloud *c, llo_point p)
{
    cloud_reserve(c, c->n + 1);
    c->p[c->n++] = p;
}
static void cloud_free(cloud *c) { free(c->p); c->p = 0; c->n = c->cap = 0; }

static void voxel_into(const cloud *in, float leaf, cloud *out)
{
    int ovf;
    cloud_reserve(out, in->n);
    out->n = llo_voxel_grid(in->p, in->n, leaf, out->p, &ovf);
}

struct llo_mapopt {
    cloud cornerLast, surfLast, outlierLast;                 /* MO:109-114 */
    cloud cornerLastDS, surfLastDS, outlierLastDS;           /* MO:111-115 */
    cloud surfTotalLast, surfTotalLastDS;                    /* MO:117-118 */
    cloud cornerFromMap, surfFromMap, cornerFromMapDS, surfFromMapDS; /* MO:123-126 */
    cloud ori, coeffSel;                                     /* MO:120-121 */
    llo_kdtree *kdCorner, *kdSurf;                           /* MO:128-129 */
    float tobe[6], sum[6], bef[6], aft[6];                   /* MO:173-178 */
    float cRoll, sRoll, cPitch, sPitch, cYaw, sYaw, tX, tY, tZ;   /* MO:219 */
    int isDegenerate; float matP[36];                        /* MO:202-203 */
    float AtA[36], AtB[6], X[6];
    int *knnC, *knnS; float *knnCd, *knnSd; int knnCn, knnSn;
};

llo_mapopt *llo_mapopt_create(void)
{
    llo_mapopt *m = (llo_mapopt *)calloc(1, sizeof(*m));   /* all poses 0, matP 0, !degenerate: MO:329-361 */
    return m;
}

void llo_mapopt_destroy(llo_mapopt *m)
{
    if (!m) return;
    cloud *cs[] = { &m->cornerLast, &m->surfLast, &m->outlierLast, &m->cornerLastDS, &m->surfLastDS,
                    &m->outlierLastDS, &m->surfTotalLast, &m->surfTotalLastDS, &m->cornerFromMap,
                    &m->surfFromMap, &m->cornerFromMapDS, &m->surfFromMapDS, &m->ori, &m->coeffSel };
    for (unsigned i = 0; i < sizeof cs / sizeof cs[0]; i++) cloud_free(cs[i]);
    llo_kdtree_free(m->kdCorner); llo_kdtree_free(m->kdSurf);
    free(m->knnC); free(m->knnS); free(m->knnCd); free(m->knnSd);
    free(m);
}

void llo_mapopt_set_map_ds(llo_mapopt *m, const llo_point *c, int mc, const llo_point *s, int ms)
{
    cloud_set(&m->cornerFromMapDS, c, mc);
    cloud_set(&m->surfFromMapDS, s, ms);
}

void llo_mapopt_set_map_raw(llo_mapopt *m, const llo_point *c, int rc, const llo_point *s, int rs)
{
    cloud_set(&m->cornerFromMap, c, rc);
    cloud_set(&m->surfFromMap, s, rs);
    voxel_into(&m->cornerFromMap, 0.2f, &m->cornerFromMapDS);   /* MO:1058-1060, leaf MO:249 */
    voxel_into(&m->surfFromMap, 0.4f, &m->surfFromMapDS);       /* MO:1062-1064, leaf MO:250 */
}

void llo_mapopt_set_scan(llo_mapopt *m, const llo_point *c, int nc, const llo_point *s, int ns,
                         const llo_point *o, int no)
{
    cloud_set(&m->cornerLast, c, nc);
    cloud_set(&m->surfLast, s, ns);
    cloud_set(&m->outlierLast, o, no);
}

void llo_mapopt_set_pose(llo_mapopt *m, const float t[6]) { memcpy(m->tobe, t, sizeof m->tobe); }
void llo_mapopt_get_pose(const llo_mapopt *m, float t[6]) { memcpy(t, m->tobe, sizeof m->tobe); }
void llo_mapopt_set_transform_sum(llo_mapopt *m, const float s[6]) { memcpy(m->sum, s, sizeof m->sum); }
void llo_mapopt_get_bef_aft(const llo_mapopt *m, float bef[6], float aft[6])
{
    memcpy(bef, m->bef, sizeof m->bef); memcpy(aft, m->aft, sizeof m->aft);
}
void llo_mapopt_get_degenerate(const llo_mapopt *m, int *d, float P[36])
{
    *d = m->isDegenerate; memcpy(P, m->matP, sizeof m->matP);
}

/* MO:1067-1091 */
void llo_mapopt_downsampleCurrentScan(llo_mapopt *m)
{
    voxel_into(&m->cornerLast, 0.2f, &m->cornerLastDS);
    voxel_into(&m->surfLast, 0.4f, &m->surfLastDS);
    voxel_into(&m->outlierLast, 0.4f, &m->outlierLastDS);
    m->surfTotalLast.n = 0;
    cloud_reserve(&m->surfTotalLast, m->surfLastDS.n + m->outlierLastDS.n);
    for (int i = 0; i < m->surfLastDS.n; i++) cloud_push(&m->surfTotalLast, m->surfLastDS.p[i]);
    for (int i = 0; i < m->outlierLastDS.n; i++) cloud_push(&m->surfTotalLast, m->outlierLastDS.p[i]);
    voxel_into(&m->surfTotalLast, 0.4f, &m->surfTotalLastDS);   /* double down-sampling, quirk C12 */
}

void llo_mapopt_build_kdtrees(llo_mapopt *m)
{
    llo_kdtree_free(m->kdCorner); llo_kdtree_free(m->kdSurf);
    m->kdCorner = llo_kdtree_build(m->cornerFromMapDS.p, m->cornerFromMapDS.n);
    m->kdSurf = llo_kdtree_build(m->surfFromMapDS.p, m->surfFromMapDS.n);
}

void llo_mapopt_clear_correspondences(llo_mapopt *m) { m->ori.n = 0; m->coeffSel.n = 0; }

/* MO:498-511 — cosf/sinf because utility.h pulls std:: overloads into scope */
static void update_sincos(llo_mapopt *m)
{
    m->cRoll = llo_cosf(m->tobe[0]);  m->sRoll = llo_sinf(m->tobe[0]);
    m->cPitch = llo_cosf(m->tobe[1]); m->sPitch = llo_sinf(m->tobe[1]);
    m->cYaw = llo_cosf(m->tobe[2]);   m->sYaw = llo_sinf(m->tobe[2]);
    m->tX = m->tobe[3]; m->tY = m->tobe[4]; m->tZ = m->tobe[5];
}

/* MO:513-527: Rz then Rx then Ry then translate */
static llo_point associate_to_map(const llo_mapopt *m, llo_point pi)
{
    llo_point po;
    float x1 = m->cYaw * pi.x - m->sYaw * pi.y;
    float y1 = m->sYaw * pi.x + m->cYaw * pi.y;
    float z1 = pi.z;
    float y2 = m->cRoll * y1 - m->sRoll * z1;
    float z2 = m->sRoll * y1 + m->cRoll * z1;
    po.x = m->cPitch * x1 + m->sPitch * z2 + m->tX;
    po.y = y2 + m->tY;
    po.z = -m->sPitch * x1 + m->cPitch * z2 + m->tZ;
    po.intensity = pi.intensity;
    return po;
}

static void knn_diag_alloc(int **idx, float **d2, int *cur, int n)
{
    *idx = (int *)realloc(*idx, sizeof(int) * 5 * (size_t)(n > 0 ? n : 1));
    *d2 = (float *)realloc(*d2, sizeof(float) * 5 * (size_t)(n > 0 ? n : 1));
    *cur = n;
}

/* MO:1093-1174 */
void llo_mapopt_cornerOptimization(llo_mapopt *m, int iterCount)
{
    (void)iterCount;                                   /* unused in the reference too (C10) */
    update_sincos(m);
    const llo_point *map = m->cornerFromMapDS.p;
    knn_diag_alloc(&m->knnC, &m->knnCd, &m->knnCn, m->cornerLastDS.n);
    for (int i = 0; i < m->cornerLastDS.n; i++) {
        llo_point pointOri = m->cornerLastDS.p[i];
        llo_point pointSel = associate_to_map(m, pointOri);
        int ind[5] = { -1, -1, -1, -1, -1 }; float sq[5] = { 0, 0, 0, 0, 0 };
        int found = llo_kdtree_knn(m->kdCorner, &pointSel.x, 5, ind, sq);
        memcpy(&m->knnC[5 * i], ind, sizeof ind); memcpy(&m->knnCd[5 * i], sq, sizeof sq);
        if (found < 5 || !(sq[4] < 1.0)) continue;

        float cx = 0, cy = 0, cz = 0;
        for (int j = 0; j < 5; j++) { cx += map[ind[j]].x; cy += map[ind[j]].y; cz += map[ind[j]].z; }
        cx /= 5; cy /= 5; cz /= 5;

        float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
        for (int j = 0; j < 5; j++) {
            float ax = map[ind[j]].x - cx, ay = map[ind[j]].y - cy, az = map[ind[j]].z - cz;
            a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
            a22 += ay * ay; a23 += ay * az;
            a33 += az * az;
        }
        a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;

        float A1[9] = { a11, a12, a13, a12, a22, a23, a13, a23, a33 }, D1[3], V1[9];
        llo_cv_eigen_f32(3, A1, D1, V1);                           /* MO:1126 */

        if (!(D1[0] > 3 * D1[1])) continue;                        /* MO:1128 */

        float x0 = pointSel.x, y0 = pointSel.y, z0 = pointSel.z;
        /* 0.1 is a double literal: the sum is evaluated in double (Appendix B) */
        float x1 = (float)(cx + 0.1 * V1[0]), y1 = (float)(cy + 0.1 * V1[1]), z1 = (float)(cz + 0.1 * V1[2]);
        float x2 = (float)(cx - 0.1 * V1[0]), y2 = (float)(cy - 0.1 * V1[1]), z2 = (float)(cz - 0.1 * V1[2]);

        /* components of (p0-p1) x (p0-p2) */
        float m11 = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1);
        float m22 = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1);
        float m33 = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
        float a012 = sqrtf(m11 * m11 + m22 * m22 + m33 * m33);
        float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
        float la = ((y1 - y2) * m11 + (z1 - z2) * m22) / a012 / l12;
        float lb = -((x1 - x2) * m11 - (z1 - z2) * m33) / a012 / l12;
        float lc = -((x1 - x2) * m22 + (y1 - y2) * m33) / a012 / l12;
        float ld2 = a012 / l12;

        float s = (float)(1 - 0.9 * fabsf(ld2));                   /* MO:1160, double */
        llo_point coeff = { s * la, s * lb, s * lc, s * ld2 };
        if (s > 0.1) {                                             /* float vs double 0.1 */
            cloud_push(&m->ori, pointOri);
            cloud_push(&m->coeffSel, coeff);
        }
    }
}

/* MO:1176-1227 */
void llo_mapopt_surfOptimization(llo_mapopt *m, int iterCount)
{
    (void)iterCount;
    update_sincos(m);
    const llo_point *map = m->surfFromMapDS.p;
    const float B0[5] = { -1, -1, -1, -1, -1 };                    /* MO:353 */
    knn_diag_alloc(&m->knnS, &m->knnSd, &m->knnSn, m->surfTotalLastDS.n);
    for (int i = 0; i < m->surfTotalLastDS.n; i++) {
        llo_point pointOri = m->surfTotalLastDS.p[i];
        llo_point pointSel = associate_to_map(m, pointOri);
        int ind[5] = { -1, -1, -1, -1, -1 }; float sq[5] = { 0, 0, 0, 0, 0 };
        int found = llo_kdtree_knn(m->kdSurf, &pointSel.x, 5, ind, sq);
        memcpy(&m->knnS[5 * i], ind, sizeof ind); memcpy(&m->knnSd[5 * i], sq, sizeof sq);
        if (found < 5 || !(sq[4] < 1.0)) continue;

        float A0[15], X0[3];
        for (int j = 0; j < 5; j++) {
            A0[3 * j] = map[ind[j]].x; A0[3 * j + 1] = map[ind[j]].y; A0[3 * j + 2] = map[ind[j]].z;
        }
        llo_cv_solve_qr_f32(5, 3, A0, B0, X0);                     /* MO:1189 */

        float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1;
        float ps = sqrtf(pa * pa + pb * pb + pc * pc);
        pa /= ps; pb /= ps; pc /= ps; pd /= ps;

        int planeValid = 1;
        for (int j = 0; j < 5; j++) {
            if (fabsf(pa * map[ind[j]].x + pb * map[ind[j]].y + pc * map[ind[j]].z + pd) > 0.2) {
                planeValid = 0;
                break;
            }
        }
        if (!planeValid) continue;

        float pd2 = pa * pointSel.x + pb * pointSel.y + pc * pointSel.z + pd;
        float s = (float)(1 - 0.9 * fabsf(pd2) /
                          sqrtf(sqrtf(pointSel.x * pointSel.x + pointSel.y * pointSel.y + pointSel.z * pointSel.z)));
        llo_point coeff = { s * pa, s * pb, s * pc, s * pd2 };
        if (s > 0.1) {
            cloud_push(&m->ori, pointOri);
            cloud_push(&m->coeffSel, coeff);
        }
    }
}

/* MO:1229-1327. returns 1 when converged */
int llo_mapopt_LMOptimization(llo_mapopt *m, int iterCount)
{
    float srx = llo_sinf(m->tobe[0]), crx = llo_cosf(m->tobe[0]);
    float sry = llo_sinf(m->tobe[1]), cry = llo_cosf(m->tobe[1]);
    float srz = llo_sinf(m->tobe[2]), crz = llo_cosf(m->tobe[2]);

    int N = m->ori.n;
    if (N < 50) return 0;                                          /* MO:1238 */

    float *A = (float *)malloc(sizeof(float) * 6 * (size_t)N);
    float *At = (float *)malloc(sizeof(float) * 6 * (size_t)N);
    float *B = (float *)malloc(sizeof(float) * (size_t)N);
    for (int i = 0; i < N; i++) {
        llo_point p = m->ori.p[i], c = m->coeffSel.p[i];
        float arx = (crx * sry * srz * p.x + crx * crz * sry * p.y - srx * sry * p.z) * c.x
                  + (-srx * srz * p.x - crz * srx * p.y - crx * p.z) * c.y
                  + (crx * cry * srz * p.x + crx * cry * crz * p.y - cry * srx * p.z) * c.z;
        float ary = ((cry * srx * srz - crz * sry) * p.x + (sry * srz + cry * crz * srx) * p.y + crx * cry * p.z) * c.x
                  + ((-cry * crz - srx * sry * srz) * p.x + (cry * srz - crz * srx * sry) * p.y - crx * sry * p.z) * c.z;
        float arz = ((crz * srx * sry - cry * srz) * p.x + (-cry * crz - srx * sry * srz) * p.y) * c.x
                  + (crx * crz * p.x - crx * srz * p.y) * c.y
                  + ((sry * srz + cry * crz * srx) * p.x + (crz * sry - cry * srx * srz) * p.y) * c.z;
        A[6 * i + 0] = arx; A[6 * i + 1] = ary; A[6 * i + 2] = arz;
        A[6 * i + 3] = c.x; A[6 * i + 4] = c.y; A[6 * i + 5] = c.z;
        B[i] = -c.intensity;
    }
    for (int i = 0; i < N; i++)
        for (int j = 0; j < 6; j++) At[(size_t)j * N + i] = A[6 * i + j];
    llo_cv_gemm_f32(6, N, 6, At, A, m->AtA);                       /* MO:1274 */
    llo_cv_gemm_f32(6, N, 1, At, B, m->AtB);                       /* MO:1275 */
    free(A); free(At); free(B);

    float X[6];
    llo_cv_solve_qr_f32(6, 6, m->AtA, m->AtB, X);                  /* MO:1276 */

    if (iterCount == 0) {                                          /* MO:1278-1299 */
        float tmp[36], E[6], V[36], V2[36], Vinv[36];
        memcpy(tmp, m->AtA, sizeof tmp);
        llo_cv_eigen_f32(6, tmp, E, V);
        memcpy(V2, V, sizeof V2);
        m->isDegenerate = 0;
        for (int i = 5; i >= 0; i--) {
            if (E[i] < 100.f) {
                for (int j = 0; j < 6; j++) V2[6 * i + j] = 0;
                m->isDegenerate = 1;
            } else break;
        }
        llo_cv_inv_f32(6, V, Vinv);
        llo_cv_gemm_f32(6, 6, 6, Vinv, V2, m->matP);
    }
    if (m->isDegenerate) {                                         /* MO:1301-1305 */
        float X2[6];
        memcpy(X2, X, sizeof X2);
        llo_cv_gemm_f32(6, 6, 1, m->matP, X2, X);
    }
    for (int i = 0; i < 6; i++) m->tobe[i] += X[i];
    memcpy(m->X, X, sizeof X);

    /* pcl::rad2deg(float) = x * 57.29578f (float); pow(float,2) is evaluated in double */
    double r0 = (double)(X[0] * 57.29578f), r1 = (double)(X[1] * 57.29578f), r2 = (double)(X[2] * 57.29578f);
    double t0 = (double)(X[3] * 100), t1 = (double)(X[4] * 100), t2 = (double)(X[5] * 100);
    float deltaR = (float)sqrt(r0 * r0 + r1 * r1 + r2 * r2);
    float deltaT = (float)sqrt(t0 * t0 + t1 * t1 + t2 * t2);
    if (deltaR < 0.05 && deltaT < 0.05) return 1;
    return 0;
}

/* MO:1329-1350 (+ transformUpdate MO:463-496 with no IMU message ever received, C15/C23) */
int llo_mapopt_scan2MapOptimization(llo_mapopt *m)
{
    int iters = 0;
    if (m->cornerFromMapDS.n > 10 && m->surfFromMapDS.n > 100) {
        llo_mapopt_build_kdtrees(m);
        for (int it = 0; it < 10; it++) {
            llo_mapopt_clear_correspondences(m);
            llo_mapopt_cornerOptimization(m, it);
            llo_mapopt_surfOptimization(m, it);
            iters++;
            if (llo_mapopt_LMOptimization(m, it)) break;
        }
        for (int i = 0; i < 6; i++) { m->bef[i] = m->sum[i]; m->aft[i] = m->tobe[i]; }
    }
    return iters;
}

static int copy_out(const cloud *c, llo_point *out, int cap)
{
    int n = c->n < cap ? c->n : cap;
    if (out && n > 0) memcpy(out, c->p, sizeof(llo_point) * (size_t)n);
    return c->n;
}

int llo_mapopt_get_scan_ds(const llo_mapopt *m, int which, llo_point *out, int cap)
{
    const cloud *c = which == 0 ? &m->cornerLastDS : which == 1 ? &m->surfLastDS
                   : which == 2 ? &m->outlierLastDS : &m->surfTotalLastDS;
    return copy_out(c, out, cap);
}
int llo_mapopt_get_map_ds(const llo_mapopt *m, int which, llo_point *out, int cap)
{
    return copy_out(which == 0 ? &m->cornerFromMapDS : &m->surfFromMapDS, out, cap);
}
int llo_mapopt_get_correspondences(const llo_mapopt *m, llo_point *ori, llo_point *coeff, int cap)
{
    copy_out(&m->ori, ori, cap);
    return copy_out(&m->coeffSel, coeff, cap);
}
int llo_mapopt_get_knn(const llo_mapopt *m, int which, int *idx5, float *d2_5, int cap)
{
    int n = which == 0 ? m->knnCn : m->knnSn;
    int c = n < cap ? n : cap;
    if (c > 0) {
        memcpy(idx5, which == 0 ? m->knnC : m->knnS, sizeof(int) * 5 * (size_t)c);
        memcpy(d2_5, which == 0 ? m->knnCd : m->knnSd, sizeof(float) * 5 * (size_t)c);
    }
    return n;
}
void llo_mapopt_get_normal_eq(const llo_mapopt *m, float AtA[36], float AtB[6], float X[6])
{
    memcpy(AtA, m->AtA, sizeof m->AtA); memcpy(AtB, m->AtB, sizeof m->AtB); memcpy(X, m->X, sizeof m->X);
}
