/*
 * llo_features.c — CPU ORACLE (test infrastructure only, see llo.h): restatement of the feature
 * extraction that precedes the odometry matcher:
 *   adjustDistortion (no-IMU branch)  FA:491-619
 *   calculateSmoothness               FA:621-641
 *   markOccludedPoints                FA:643-678
 *   extractFeatures                   FA:680-784
 * and of the third-party primitive they depend on for their exact result: libstdc++'s std::sort
 * (bits/stl_algo.h: introsort with median-of-3 to first, unguarded Hoare partition, threshold 16, final
 * insertion sort, heap sort on depth exhaustion).  The sort is NOT stable and the reference compares by
 * curvature only (FA:57-61), so which of several equal-curvature points is picked depends on the exact
 * sequence of swaps: the restatement reproduces it move for move.  Pinned against the compiled reference
 * (oracle/_ref/libref_fa.so) and against libstdc++ itself in tests/test_oracle_features.py.
 *
 * State that survives from sweep to sweep in the reference is kept here as well: calculateSmoothness writes
 * indices [5, n-5) only, while sector 0 of the first ring starts at index 4 (IP:318), so element 4 of
 * cloudSmoothness (initially {0, 0}) and the entries 0..4 of the per-point arrays are never reset.
 */
#include "llo.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float value; uint32_t ind; } smooth_t;

struct llo_features {
    int n_scan, horizon, cap;
    float *curv; int *picked; int *label; smooth_t *smooth;
    float edge_threshold, surf_threshold, scan_period, leaf;
};

/* ---------------------------------------------------------------- std::sort, libstdc++ */
static inline int less_(const smooth_t *a, const smooth_t *b) { return a->value < b->value; }
static inline void swap_(smooth_t *a, smooth_t *b) { smooth_t t = *a; *a = *b; *b = t; }

static void push_heap_(smooth_t *first, long hole, long top, smooth_t value)
{
    long parent = (hole - 1) / 2;
    while (hole > top && less_(first + parent, &value)) {
        first[hole] = first[parent]; hole = parent; parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void adjust_heap_(smooth_t *first, long hole, long len, smooth_t value)
{
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less_(first + child, first + (child - 1))) child--;
        first[hole] = first[child]; hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1]; hole = child - 1;
    }
    push_heap_(first, hole, top, value);
}
static void heap_sort_(smooth_t *first, smooth_t *last)
{   /* std::__partial_sort(first, last, last): __heap_select = make_heap (no element beyond middle), then __sort_heap */
    const long len = last - first;
    if (len >= 2) {
        long parent = (len - 2) / 2;
        for (;;) {
            smooth_t v = first[parent];
            adjust_heap_(first, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        smooth_t v = *last; *last = *first;
        adjust_heap_(first, 0, last - first, v);
    }
}
static void move_median_to_first_(smooth_t *result, smooth_t *a, smooth_t *b, smooth_t *c)
{
    if (less_(a, b)) {
        if (less_(b, c)) swap_(result, b);
        else if (less_(a, c)) swap_(result, c);
        else swap_(result, a);
    } else if (less_(a, c)) swap_(result, a);
    else if (less_(b, c)) swap_(result, c);
    else swap_(result, b);
}
static smooth_t *unguarded_partition_(smooth_t *first, smooth_t *last, smooth_t *pivot)
{
    for (;;) {
        while (less_(first, pivot)) ++first;
        --last;
        while (less_(pivot, last)) --last;
        if (!(first < last)) return first;
        swap_(first, last);
        ++first;
    }
}
static void introsort_loop_(smooth_t *first, smooth_t *last, long depth_limit)
{
    while (last - first > 16) {
        if (depth_limit == 0) { heap_sort_(first, last); return; }
        --depth_limit;
        smooth_t *mid = first + (last - first) / 2;
        move_median_to_first_(first, first + 1, mid, last - 1);
        smooth_t *cut = unguarded_partition_(first + 1, last, first);
        introsort_loop_(cut, last, depth_limit);
        last = cut;
    }
}
static void unguarded_linear_insert_(smooth_t *last)
{
    smooth_t val = *last;
    smooth_t *next = last - 1;
    while (less_(&val, next)) { *last = *next; last = next; --next; }
    *last = val;
}
static void insertion_sort_(smooth_t *first, smooth_t *last)
{
    if (first == last) return;
    for (smooth_t *i = first + 1; i != last; ++i) {
        if (less_(i, first)) {
            smooth_t val = *i;
            memmove(first + 1, first, (size_t)(i - first) * sizeof(smooth_t));
            *first = val;
        } else unguarded_linear_insert_(i);
    }
}
static void final_insertion_sort_(smooth_t *first, smooth_t *last)
{
    if (last - first > 16) {
        insertion_sort_(first, first + 16);
        for (smooth_t *i = first + 16; i != last; ++i) unguarded_linear_insert_(i);
    } else insertion_sort_(first, last);
}
static long lg_(long n) { long k = 0; while (n > 1) { n >>= 1; k++; } return k; }

static void std_sort_(smooth_t *first, smooth_t *last, long depth_limit)
{
    if (first == last) return;
    introsort_loop_(first, last, depth_limit < 0 ? lg_(last - first) * 2 : depth_limit);
    final_insertion_sort_(first, last);
}

void llo_std_sort_by_value(float *value, uint32_t *ind, int n, int depth_limit)
{
    smooth_t *v = (smooth_t *)malloc(sizeof(smooth_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) { v[i].value = value[i]; v[i].ind = ind[i]; }
    std_sort_(v, v + n, depth_limit);
    for (int i = 0; i < n; i++) { value[i] = v[i].value; ind[i] = v[i].ind; }
    free(v);
}

/* ---------------------------------------------------------------- state */
llo_features *llo_features_create(int n_scan, int horizon)
{
    llo_features *f = (llo_features *)calloc(1, sizeof(*f));
    f->n_scan = n_scan; f->horizon = horizon; f->cap = n_scan * horizon;
    f->curv = (float *)calloc((size_t)f->cap, sizeof(float));       /* FA:210-212 (defined as zero, see ref_fa.cpp) */
    f->picked = (int *)calloc((size_t)f->cap, sizeof(int));
    f->label = (int *)calloc((size_t)f->cap, sizeof(int));
    f->smooth = (smooth_t *)calloc((size_t)f->cap, sizeof(smooth_t));  /* FA:223: value-initialised {0, 0} */
    f->edge_threshold = 0.1f; f->surf_threshold = 0.1f;             /* UT:116-117 */
    f->scan_period = 0.1f;                                           /* UT:107 */
    f->leaf = 0.2f;                                                  /* FA:214 */
    return f;
}
void llo_features_destroy(llo_features *f)
{
    if (!f) return;
    free(f->curv); free(f->picked); free(f->label); free(f->smooth); free(f);
}
void llo_features_get_state(const llo_features *f, int n, float *curv, int *picked, int *label)
{
    for (int i = 0; i < n; i++) { curv[i] = f->curv[i]; picked[i] = f->picked[i]; label[i] = f->label[i]; }
}

/* ---------------------------------------------------------------- the four functions */
/* FA:491-617 with imuPointerLast < 0 */
void llo_adjust_distortion(llo_point *cloud, int n, float start_ori, float end_ori, float ori_diff, float scan_period)
{
    int half_passed = 0;
    for (int i = 0; i < n; i++) {
        llo_point p;
        p.x = cloud[i].y; p.y = cloud[i].z; p.z = cloud[i].x;
        float ori = -atan2f(p.x, p.z);
        if (!half_passed) {
            if (ori < start_ori - M_PI / 2) ori += 2 * M_PI;
            else if (ori > start_ori + M_PI * 3 / 2) ori -= 2 * M_PI;
            if (ori - start_ori > M_PI) half_passed = 1;
        } else {
            ori += 2 * M_PI;
            if (ori < end_ori - M_PI * 3 / 2) ori += 2 * M_PI;
            else if (ori > end_ori + M_PI / 2) ori -= 2 * M_PI;
        }
        float rel_time = (ori - start_ori) / ori_diff;
        p.intensity = (int)cloud[i].intensity + scan_period * rel_time;
        cloud[i] = p;
    }
}

static inline int col_diff_(const uint32_t *col, int a, int b) { return abs((int)(col[a] - col[b])); }

static void mark_neighbors_(llo_features *f, const uint32_t *col, int ind)
{   /* FA:727-740 == FA:758-773 */
    f->picked[ind] = 1;
    for (int l = 1; l <= 5; l++) {
        if (col_diff_(col, ind + l, ind + l - 1) > 10) break;
        f->picked[ind + l] = 1;
    }
    for (int l = -1; l >= -5; l--) {
        /* a stale record may name point 0 (see the header): the reference then reads and writes before its arrays
         * (undefined behaviour whose effects no later statement reads); the restatement stops at the array start */
        if (ind + l < 0) break;
        if (col_diff_(col, ind + l, ind + l + 1) > 10) break;
        f->picked[ind + l] = 1;
    }
}

/* ground / col / range must be readable up to index n + 5 and at every index a stale smoothness record may hold;
 * the caller passes arrays of f->cap entries, zero beyond n (as ref_fa_set_segmented builds them).
 * out[0..3]: cornerPointsSharp, cornerPointsLessSharp, surfPointsFlat, surfPointsLessFlat; each with room for n. */
void llo_features_extract(llo_features *f, llo_point *cloud, int n, const int *start_ring, const int *end_ring,
                          float start_ori, float end_ori, float ori_diff,
                          const uint8_t *ground, const uint32_t *col, const float *range,
                          llo_point *const out[4], int n_out[4])
{
    llo_adjust_distortion(cloud, n, start_ori, end_ori, ori_diff, f->scan_period);
    /* calculateSmoothness FA:621-641 */
    for (int i = 5; i < n - 5; i++) {
        float d = range[i - 5] + range[i - 4] + range[i - 3] + range[i - 2] + range[i - 1] - range[i] * 10
                + range[i + 1] + range[i + 2] + range[i + 3] + range[i + 4] + range[i + 5];
        f->curv[i] = d * d;
        f->picked[i] = 0; f->label[i] = 0;
        f->smooth[i].value = f->curv[i]; f->smooth[i].ind = (uint32_t)i;
    }
    /* markOccludedPoints FA:643-678 */
    for (int i = 5; i < n - 6; ++i) {
        float depth1 = range[i], depth2 = range[i + 1];
        int cd = col_diff_(col, i + 1, i);
        if (cd < 10) {
            if (depth1 - depth2 > 0.3) { for (int k = -5; k <= 0; k++) f->picked[i + k] = 1; }
            else if (depth2 - depth1 > 0.3) { for (int k = 1; k <= 6; k++) f->picked[i + k] = 1; }
        }
        float diff1 = fabsf(range[i - 1] - range[i]);
        float diff2 = fabsf(range[i + 1] - range[i]);
        if (diff1 > 0.02 * range[i] && diff2 > 0.02 * range[i]) f->picked[i] = 1;
    }
    /* extractFeatures FA:680-784 */
    n_out[0] = n_out[1] = n_out[2] = n_out[3] = 0;
    llo_point *scan = (llo_point *)malloc(sizeof(llo_point) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < f->n_scan; i++) {
        int n_scan_pts = 0;
        for (int j = 0; j < 6; j++) {
            int sp = (start_ring[i] * (6 - j) + end_ring[i] * j) / 6;
            int ep = (start_ring[i] * (5 - j) + end_ring[i] * (j + 1)) / 6 - 1;
            if (sp >= ep) continue;
            std_sort_(f->smooth + sp, f->smooth + ep, -1);           /* [sp, ep): element ep keeps its place */
            int largest = 0;
            for (int k = ep; k >= sp; k--) {
                int ind = (int)f->smooth[k].ind;
                if (f->picked[ind] == 0 && f->curv[ind] > f->edge_threshold && ground[ind] == 0) {
                    largest++;
                    if (largest <= 2) { f->label[ind] = 2; out[0][n_out[0]++] = cloud[ind]; out[1][n_out[1]++] = cloud[ind]; }
                    else if (largest <= 20) { f->label[ind] = 1; out[1][n_out[1]++] = cloud[ind]; }
                    else break;
                    mark_neighbors_(f, col, ind);
                }
            }
            int smallest = 0;
            for (int k = sp; k <= ep; k++) {
                int ind = (int)f->smooth[k].ind;
                if (f->picked[ind] == 0 && f->curv[ind] < f->surf_threshold && ground[ind] != 0) {
                    f->label[ind] = -1;
                    out[2][n_out[2]++] = cloud[ind];
                    smallest++;
                    if (smallest >= 4) break;
                    mark_neighbors_(f, col, ind);
                }
            }
            for (int k = sp; k <= ep; k++)
                if (f->label[k] <= 0) scan[n_scan_pts++] = cloud[k];
        }
        int ovf = 0;
        if (n_scan_pts > 0) n_out[3] += llo_voxel_grid(scan, n_scan_pts, f->leaf, out[3] + n_out[3], &ovf);   /* FA:778-782 */
    }
    free(scan);
}

/* TransformToEnd FA:885-953 for every point of a cloud, in place, with the IMU terms of a node that never received an
 * IMU message (imuRollStart = ... = 0, shifts 0: cos -> 1, sin -> 0, kept in the expressions as the reference has them).
 * sin/cos through llo_sinf/llo_cosf (llo_set_trig_mode). */
void llo_transform_to_end(const float T[6], llo_point *cloud, int n)
{
    const float zero = 0.f;
    const float cosImuRollStart = llo_cosf(zero), cosImuPitchStart = llo_cosf(zero), cosImuYawStart = llo_cosf(zero);
    const float sinImuRollStart = llo_sinf(zero), sinImuPitchStart = llo_sinf(zero), sinImuYawStart = llo_sinf(zero);
    const float imuShiftFromStartX = 0.f, imuShiftFromStartY = 0.f, imuShiftFromStartZ = 0.f;
    const float imuYawLast = 0.f, imuPitchLast = 0.f, imuRollLast = 0.f;
    for (int i = 0; i < n; i++) {
        const llo_point *pi = &cloud[i];
        float s = 10 * (pi->intensity - (int)pi->intensity);
        float rx = s * T[0], ry = s * T[1], rz = s * T[2], tx = s * T[3], ty = s * T[4], tz = s * T[5];
        float x1 = llo_cosf(rz) * (pi->x - tx) + llo_sinf(rz) * (pi->y - ty);
        float y1 = -llo_sinf(rz) * (pi->x - tx) + llo_cosf(rz) * (pi->y - ty);
        float z1 = (pi->z - tz);
        float x2 = x1;
        float y2 = llo_cosf(rx) * y1 + llo_sinf(rx) * z1;
        float z2 = -llo_sinf(rx) * y1 + llo_cosf(rx) * z1;
        float x3 = llo_cosf(ry) * x2 - llo_sinf(ry) * z2;
        float y3 = y2;
        float z3 = llo_sinf(ry) * x2 + llo_cosf(ry) * z2;
        rx = T[0]; ry = T[1]; rz = T[2]; tx = T[3]; ty = T[4]; tz = T[5];
        float x4 = llo_cosf(ry) * x3 + llo_sinf(ry) * z3;
        float y4 = y3;
        float z4 = -llo_sinf(ry) * x3 + llo_cosf(ry) * z3;
        float x5 = x4;
        float y5 = llo_cosf(rx) * y4 - llo_sinf(rx) * z4;
        float z5 = llo_sinf(rx) * y4 + llo_cosf(rx) * z4;
        float x6 = llo_cosf(rz) * x5 - llo_sinf(rz) * y5 + tx;
        float y6 = llo_sinf(rz) * x5 + llo_cosf(rz) * y5 + ty;
        float z6 = z5 + tz;
        float x7 = cosImuRollStart * (x6 - imuShiftFromStartX) - sinImuRollStart * (y6 - imuShiftFromStartY);
        float y7 = sinImuRollStart * (x6 - imuShiftFromStartX) + cosImuRollStart * (y6 - imuShiftFromStartY);
        float z7 = z6 - imuShiftFromStartZ;
        float x8 = x7;
        float y8 = cosImuPitchStart * y7 - sinImuPitchStart * z7;
        float z8 = sinImuPitchStart * y7 + cosImuPitchStart * z7;
        float x9 = cosImuYawStart * x8 + sinImuYawStart * z8;
        float y9 = y8;
        float z9 = -sinImuYawStart * x8 + cosImuYawStart * z8;
        float x10 = llo_cosf(imuYawLast) * x9 - llo_sinf(imuYawLast) * z9;
        float y10 = y9;
        float z10 = llo_sinf(imuYawLast) * x9 + llo_cosf(imuYawLast) * z9;
        float x11 = x10;
        float y11 = llo_cosf(imuPitchLast) * y10 + llo_sinf(imuPitchLast) * z10;
        float z11 = -llo_sinf(imuPitchLast) * y10 + llo_cosf(imuPitchLast) * z10;
        llo_point po;
        po.x = llo_cosf(imuRollLast) * x11 + llo_sinf(imuRollLast) * y11;
        po.y = -llo_sinf(imuRollLast) * x11 + llo_cosf(imuRollLast) * y11;
        po.z = z11;
        po.intensity = (int)pi->intensity;
        cloud[i] = po;
    }
}
