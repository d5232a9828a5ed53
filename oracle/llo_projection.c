/*
 * llo_projection.c — CPU ORACLE (test infrastructure only, see llo.h): restatement of the reference's imageProjection
 * node (LeGO-LOAM/src/imageProjection.cpp = IP), the producer of what featureAssociation consumes (the
 * next row of the path):
 *   findStartEndAngle   IP:199-211
 *   projectPointCloud   IP:213-257   (useCloudRing = true, UT:60: the row is the point's ring)
 *   groundRemoval       IP:259-310
 *   labelComponents     IP:370-448   (BFS with the reference's queue order and its lineCountFlag quirk: the seed's row
 *                                     only counts when another pushed point lies in it)
 *   cloudSegmentation   IP:312-368
 * Pinned bit-for-bit against the compiled reference (oracle/_ref/libref_ip.so) in tests/test_oracle_projection.py.
 * Mixed precision follows the C++ promotions of the reference (float members against double M_PI expressions);
 * sin / atan2 / sqrt on float arguments are the float overloads (utility.h: <cmath> + using namespace std).
 */
#include "llo.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct llo_projection {
    int n_scan, horizon, ground_scan_ind;
    float ang_res_x, ang_res_y, sensor_min_range, sensor_mount_angle, segment_theta, alpha_x, alpha_y;
    int valid_point_num, valid_line_num;
    float *range_mat; int8_t *ground_mat; int32_t *label_mat;
    llo_point *full;              /* fullCloud, intensity -1 = no return */
    uint16_t *qx, *qy, *px, *py;
    int label_count;
    /* outputs */
    llo_point *seg, *outlier; int n_seg, n_outlier;
    int32_t *start_ring, *end_ring; float start_ori, end_ori, ori_diff;
    uint8_t *ground_flag; uint32_t *col_ind; float *seg_range;
};

llo_projection *llo_projection_create(int n_scan, int horizon, float ang_res_x, float ang_res_y, int ground_scan_ind)
{
    llo_projection *p = (llo_projection *)calloc(1, sizeof(*p));
    const size_t cap = (size_t)n_scan * horizon;
    p->n_scan = n_scan; p->horizon = horizon; p->ground_scan_ind = ground_scan_ind;
    p->ang_res_x = ang_res_x; p->ang_res_y = ang_res_y;
    p->sensor_min_range = 1.0f; p->sensor_mount_angle = 0.0f;                 /* UT:111-112 */
    p->segment_theta = (float)(60.0 / 180.0 * M_PI);                          /* UT:113 */
    p->valid_point_num = 5; p->valid_line_num = 3;                            /* UT:114-115 */
    p->alpha_x = (float)(ang_res_x / 180.0 * M_PI);                           /* UT:116-117 */
    p->alpha_y = (float)(ang_res_y / 180.0 * M_PI);
    p->range_mat = (float *)malloc(sizeof(float) * cap); p->ground_mat = (int8_t *)malloc(cap);
    p->label_mat = (int32_t *)malloc(sizeof(int32_t) * cap); p->full = (llo_point *)malloc(sizeof(llo_point) * cap);
    p->qx = (uint16_t *)malloc(2 * cap); p->qy = (uint16_t *)malloc(2 * cap);
    p->px = (uint16_t *)malloc(2 * cap); p->py = (uint16_t *)malloc(2 * cap);
    p->seg = (llo_point *)malloc(sizeof(llo_point) * cap); p->outlier = (llo_point *)malloc(sizeof(llo_point) * cap);
    p->start_ring = (int32_t *)calloc((size_t)n_scan, 4); p->end_ring = (int32_t *)calloc((size_t)n_scan, 4);
    p->ground_flag = (uint8_t *)calloc(cap, 1); p->col_ind = (uint32_t *)calloc(cap, 4); p->seg_range = (float *)calloc(cap, 4);
    return p;
}

void llo_projection_destroy(llo_projection *p)
{
    if (!p) return;
    free(p->range_mat); free(p->ground_mat); free(p->label_mat); free(p->full); free(p->qx); free(p->qy); free(p->px);
    free(p->py); free(p->seg); free(p->outlier); free(p->start_ring); free(p->end_ring); free(p->ground_flag);
    free(p->col_ind); free(p->seg_range); free(p);
}

/* IP:370-448 */
static void label_components(llo_projection *p, int row, int col)
{
    const int H = p->horizon, N = p->n_scan;
    static const int dxs[4] = { -1, 0, 0, 1 }, dys[4] = { 0, 1, -1, 0 };     /* neighborIterator, IP:132-136 */
    uint8_t line_flag[256];
    memset(line_flag, 0, sizeof line_flag);
    p->qx[0] = (uint16_t)row; p->qy[0] = (uint16_t)col;
    int qsize = 1, qstart = 0, qend = 1;
    p->px[0] = (uint16_t)row; p->py[0] = (uint16_t)col;
    int pushed = 1;
    while (qsize > 0) {
        const int fx = p->qx[qstart], fy = p->qy[qstart];
        --qsize; ++qstart;
        p->label_mat[fx * H + fy] = p->label_count;
        for (int k = 0; k < 4; k++) {
            int tx = fx + dxs[k], ty = fy + dys[k];
            if (tx < 0 || tx >= N) continue;
            if (ty < 0) ty = H - 1;
            if (ty >= H) ty = 0;
            if (p->label_mat[tx * H + ty] != 0) continue;
            const float a = p->range_mat[fx * H + fy], b = p->range_mat[tx * H + ty];
            const float d1 = a > b ? a : b;          /* std::max / std::min */
            const float d2 = b < a ? b : a;
            const float alpha = dxs[k] == 0 ? p->alpha_x : p->alpha_y;
            const float angle = atan2f(d2 * sinf(alpha), (d1 - d2 * cosf(alpha)));
            if (angle > p->segment_theta) {
                p->qx[qend] = (uint16_t)tx; p->qy[qend] = (uint16_t)ty;
                ++qsize; ++qend;
                p->label_mat[tx * H + ty] = p->label_count;
                line_flag[tx] = 1;
                p->px[pushed] = (uint16_t)tx; p->py[pushed] = (uint16_t)ty;
                ++pushed;
            }
        }
    }
    int feasible = 0;
    if (pushed >= 30) feasible = 1;
    else if (pushed >= p->valid_point_num) {
        int lines = 0;
        for (int i = 0; i < N; i++) if (line_flag[i]) ++lines;
        if (lines >= p->valid_line_num) feasible = 1;
    }
    if (feasible) ++p->label_count;
    else for (int i = 0; i < pushed; i++) p->label_mat[p->px[i] * H + p->py[i]] = 999999;
}

void llo_projection_process(llo_projection *p, const llo_point *cloud, const uint16_t *ring, int n)
{
    const int H = p->horizon, N = p->n_scan;
    const size_t cap = (size_t)N * H;
    /* resetParameters IP:144-157 */
    for (size_t i = 0; i < cap; i++) {
        p->range_mat[i] = FLT_MAX; p->ground_mat[i] = 0; p->label_mat[i] = 0;
        p->full[i].x = p->full[i].y = p->full[i].z = NAN; p->full[i].intensity = -1;
    }
    p->label_count = 1;
    p->n_seg = p->n_outlier = 0;
    if (n <= 0) return;
    /* findStartEndAngle IP:199-211 (the cloud is free of NaN points, IP:170) */
    p->start_ori = -atan2f(cloud[0].y, cloud[0].x);
    p->end_ori = (float)(-atan2f(cloud[n - 1].y, cloud[n - 1].x) + 2 * M_PI);
    if (p->end_ori - p->start_ori > 3 * M_PI) p->end_ori = (float)(p->end_ori - 2 * M_PI);
    else if (p->end_ori - p->start_ori < M_PI) p->end_ori = (float)(p->end_ori + 2 * M_PI);
    p->ori_diff = p->end_ori - p->start_ori;
    /* projectPointCloud IP:213-257 */
    for (int i = 0; i < n; i++) {
        llo_point pt = cloud[i];
        const size_t row = ring[i];
        if (row >= (size_t)N) continue;
        const float horizon_angle = (float)(atan2f(pt.x, pt.y) * 180 / M_PI);
        size_t column = (size_t)(-round((horizon_angle - 90.0) / p->ang_res_x) + H / 2);
        if (column >= (size_t)H) column -= H;
        if (column >= (size_t)H) continue;
        const float range = sqrtf(pt.x * pt.x + pt.y * pt.y + pt.z * pt.z);
        if (range < p->sensor_min_range) continue;
        p->range_mat[row * H + column] = range;
        pt.intensity = (float)((float)row + (float)column / 10000.0);
        p->full[column + row * H] = pt;
    }
    /* groundRemoval IP:259-310 */
    for (int j = 0; j < H; ++j)
        for (int i = 0; i < p->ground_scan_ind; ++i) {
            const size_t lo = (size_t)j + (size_t)i * H, up = (size_t)j + (size_t)(i + 1) * H;
            if (p->full[lo].intensity == -1 || p->full[up].intensity == -1) { p->ground_mat[i * H + j] = -1; continue; }
            const float dx = p->full[up].x - p->full[lo].x, dy = p->full[up].y - p->full[lo].y, dz = p->full[up].z - p->full[lo].z;
            const float angle = (float)(atan2f(dz, sqrtf(dx * dx + dy * dy)) * 180 / M_PI);
            if (fabsf(angle - p->sensor_mount_angle) <= 10) { p->ground_mat[i * H + j] = 1; p->ground_mat[(i + 1) * H + j] = 1; }
        }
    for (size_t i = 0; i < cap; i++)
        if (p->ground_mat[i] == 1 || p->range_mat[i] == FLT_MAX) p->label_mat[i] = -1;
    /* cloudSegmentation IP:312-368 */
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < H; ++j)
            if (p->label_mat[i * H + j] == 0) label_components(p, i, j);
    int size = 0;
    for (int i = 0; i < N; ++i) {
        p->start_ring[i] = size - 1 + 5;
        for (int j = 0; j < H; ++j) {
            const int lab = p->label_mat[i * H + j];
            const int gnd = p->ground_mat[i * H + j] == 1;
            if (lab > 0 || gnd) {
                if (lab == 999999) {
                    if (i > p->ground_scan_ind && j % 5 == 0) p->outlier[p->n_outlier++] = p->full[j + i * H];
                    continue;
                }
                if (gnd) { if (j % 5 != 0 && j > 5 && j < H - 5) continue; }
                p->ground_flag[size] = (uint8_t)gnd;
                p->col_ind[size] = (uint32_t)j;
                p->seg_range[size] = p->range_mat[i * H + j];
                p->seg[size] = p->full[j + i * H];
                ++size;
            }
        }
        p->end_ring[i] = size - 1 - 5;
    }
    p->n_seg = size;
}

int llo_projection_get_cloud(const llo_projection *p, int which, llo_point *out, int cap)
{
    const llo_point *src = which == 0 ? p->seg : p->outlier;
    const int n = which == 0 ? p->n_seg : p->n_outlier;
    for (int i = 0; i < n && i < cap; i++) out[i] = src[i];
    return n;
}

void llo_projection_get_info(const llo_projection *p, int *start_ring, int *end_ring, float ori[3], uint8_t *ground,
                             uint32_t *col, float *range, int n)
{
    for (int i = 0; i < p->n_scan; i++) { start_ring[i] = p->start_ring[i]; end_ring[i] = p->end_ring[i]; }
    ori[0] = p->start_ori; ori[1] = p->end_ori; ori[2] = p->ori_diff;
    for (int i = 0; i < n; i++) { ground[i] = p->ground_flag[i]; col[i] = p->col_ind[i]; range[i] = p->seg_range[i]; }
}

void llo_projection_get_images(const llo_projection *p, float *range_mat, int8_t *ground_mat, int32_t *label_mat)
{
    const size_t cap = (size_t)p->n_scan * p->horizon;
    memcpy(range_mat, p->range_mat, sizeof(float) * cap); memcpy(ground_mat, p->ground_mat, cap);
    memcpy(label_mat, p->label_mat, sizeof(int32_t) * cap);
}
