/*
 * llo_kdtree.c — ORACLE (test infrastructure): exact k-nearest-neighbour search
 * with the contract of pcl::KdTreeFLANN<pcl::PointXYZI> as the reference uses it
 * (setInputCloud MO:1333-1334, FA:1615-1616, FA:1786-1787; nearestKSearch
 * MO:1099, MO:1181 with k=5 and FA:1054, FA:1165 with k=1).
 *
 * PCL/FLANN are un-vendored and absent offline.  The published behaviour
 * (SURVEY.md Appendix A.2): flann::KDTreeSingleIndex (leaf_max_size 15,
 * reorder=true), distance flann::L2_Simple<float> = sequential float sum of
 * squared differences over x,y,z, exact search (checks=-1, eps=0), results
 * sorted ascending.  Any exact kNN is therefore equivalent up to equal-distance
 * ties; this oracle defines ties as "smaller point index first" and the tree
 * below returns exactly what llo_knn_bruteforce returns.
 *
 * The tree is a bounding-box kd-tree in the FLANN single-index style (split on
 * the widest dimension at the box midpoint clamped to the data, leaves of <= 15
 * reordered points) so that its cost profile is representative of the reference
 * when it is used as the CPU baseline.
 */
#include "llo.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define LEAF_MAX 15

typedef struct {
    int left, right;       /* children (internal) or [left,right) point range (leaf) */
    int cutdim;            /* -1 for a leaf */
    float divlow, divhigh;
} kd_node;

struct llo_kdtree {
    int n;
    float *pts;            /* reordered xyz, 3*n */
    int *ids;              /* original index of reordered point */
    kd_node *nodes;
    int n_nodes, cap_nodes;
    float bbox[6];
};

static inline float l2_simple(const float *a, const float *b)
{
    float d = 0.f, diff;
    diff = a[0] - b[0]; d += diff * diff;
    diff = a[1] - b[1]; d += diff * diff;
    diff = a[2] - b[2]; d += diff * diff;
    return d;
}

/* result set: ascending by (d2, idx) */
static inline void rs_insert(int k, int *cnt, int *idx, float *d2, float d, int id)
{
    int c = *cnt;
    if (c == k) {
        if (d > d2[k - 1] || (d == d2[k - 1] && id > idx[k - 1])) return;
    }
    int i = (c < k) ? c : k - 1;
    while (i > 0 && (d2[i - 1] > d || (d2[i - 1] == d && idx[i - 1] > id))) {
        d2[i] = d2[i - 1]; idx[i] = idx[i - 1]; i--;
    }
    d2[i] = d; idx[i] = id;
    if (c < k) *cnt = c + 1;
}

int llo_knn_bruteforce(const llo_point *pts, int n, const float *q, int k, int *idx, float *d2)
{
    int cnt = 0;
    if (k > n) k = n;
    for (int i = 0; i < n; i++) {
        float d = l2_simple(q, &pts[i].x);
        rs_insert(k, &cnt, idx, d2, d, i);
    }
    return cnt;
}

static int new_node(llo_kdtree *t)
{
    if (t->n_nodes == t->cap_nodes) {
        t->cap_nodes = t->cap_nodes ? t->cap_nodes * 2 : 1024;
        t->nodes = (kd_node *)realloc(t->nodes, sizeof(kd_node) * (size_t)t->cap_nodes);
    }
    return t->n_nodes++;
}

/* src: input points (stride 4 floats); ind: permutation being partitioned */
static int build_rec(llo_kdtree *t, const llo_point *src, int *ind, int lo, int hi, float *bbox)
{
    int ni = new_node(t);
    if (hi - lo <= LEAF_MAX) {
        t->nodes[ni].cutdim = -1;
        t->nodes[ni].left = lo; t->nodes[ni].right = hi;
        for (int a = 0; a < 3; a++) { bbox[2 * a] = FLT_MAX; bbox[2 * a + 1] = -FLT_MAX; }
        for (int i = lo; i < hi; i++) {
            const float *p = &src[ind[i]].x;
            for (int a = 0; a < 3; a++) {
                if (p[a] < bbox[2 * a]) bbox[2 * a] = p[a];
                if (p[a] > bbox[2 * a + 1]) bbox[2 * a + 1] = p[a];
            }
        }
        return ni;
    }
    /* widest dimension of the current box, by actual data spread */
    int cutdim = 0; float best = -1.f, mn = 0, mx = 0;
    for (int a = 0; a < 3; a++) {
        float lo_a = FLT_MAX, hi_a = -FLT_MAX;
        for (int i = lo; i < hi; i++) {
            float v = (&src[ind[i]].x)[a];
            if (v < lo_a) lo_a = v;
            if (v > hi_a) hi_a = v;
        }
        if (hi_a - lo_a > best) { best = hi_a - lo_a; cutdim = a; mn = lo_a; mx = hi_a; }
    }
    float cut = (bbox[2 * cutdim] + bbox[2 * cutdim + 1]) * 0.5f;
    if (cut < mn) cut = mn; else if (cut > mx) cut = mx;
    /* three-way partition: < cut | == cut | > cut */
    int i = lo, j = hi - 1;
    for (;;) {
        while (i <= j && (&src[ind[i]].x)[cutdim] < cut) i++;
        while (i <= j && (&src[ind[j]].x)[cutdim] >= cut) j--;
        if (i > j) break;
        int tmp = ind[i]; ind[i] = ind[j]; ind[j] = tmp; i++; j--;
    }
    int lim1 = i;
    j = hi - 1;
    for (;;) {
        while (i <= j && (&src[ind[i]].x)[cutdim] <= cut) i++;
        while (i <= j && (&src[ind[j]].x)[cutdim] > cut) j--;
        if (i > j) break;
        int tmp = ind[i]; ind[i] = ind[j]; ind[j] = tmp; i++; j--;
    }
    int lim2 = i;
    int half = lo + (hi - lo) / 2, mid;
    if (lim1 > half) mid = lim1; else if (lim2 < half) mid = lim2; else mid = half;
    if (mid == lo || mid == hi) mid = half;   /* all equal along cutdim: split by count */

    float lb[6], rb[6];
    memcpy(lb, bbox, sizeof lb); memcpy(rb, bbox, sizeof rb);
    lb[2 * cutdim + 1] = cut; rb[2 * cutdim] = cut;
    int l = build_rec(t, src, ind, lo, mid, lb);
    int r = build_rec(t, src, ind, mid, hi, rb);
    kd_node *nd = &t->nodes[ni];
    nd->cutdim = cutdim; nd->left = l; nd->right = r;
    nd->divlow = lb[2 * cutdim + 1]; nd->divhigh = rb[2 * cutdim];
    for (int a = 0; a < 3; a++) {
        bbox[2 * a] = lb[2 * a] < rb[2 * a] ? lb[2 * a] : rb[2 * a];
        bbox[2 * a + 1] = lb[2 * a + 1] > rb[2 * a + 1] ? lb[2 * a + 1] : rb[2 * a + 1];
    }
    return ni;
}

llo_kdtree *llo_kdtree_build(const llo_point *pts, int n)
{
    llo_kdtree *t = (llo_kdtree *)calloc(1, sizeof(*t));
    t->n = n;
    if (n <= 0) return t;
    int *ind = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) ind[i] = i;
    for (int a = 0; a < 3; a++) { t->bbox[2 * a] = FLT_MAX; t->bbox[2 * a + 1] = -FLT_MAX; }
    for (int i = 0; i < n; i++)
        for (int a = 0; a < 3; a++) {
            float v = (&pts[i].x)[a];
            if (v < t->bbox[2 * a]) t->bbox[2 * a] = v;
            if (v > t->bbox[2 * a + 1]) t->bbox[2 * a + 1] = v;
        }
    float bb[6]; memcpy(bb, t->bbox, sizeof bb);
    build_rec(t, pts, ind, 0, n, bb);
    memcpy(t->bbox, bb, sizeof bb);
    t->pts = (float *)malloc(sizeof(float) * 3 * (size_t)n);
    t->ids = ind;
    for (int i = 0; i < n; i++) {
        t->pts[3 * i] = pts[ind[i]].x; t->pts[3 * i + 1] = pts[ind[i]].y; t->pts[3 * i + 2] = pts[ind[i]].z;
    }
    return t;
}

void llo_kdtree_free(llo_kdtree *t)
{
    if (!t) return;
    free(t->pts); free(t->ids); free(t->nodes); free(t);
}

typedef struct { const llo_kdtree *t; const float *q; int k, cnt; int *idx; float *d2; } kd_search;

/* mindist is a LOWER BOUND of the true distance to the cell (computed in double
 * so rounding can never prune a point the float distance would accept) */
static void search_rec(kd_search *s, int ni, double mindist, double dists[3])
{
    const kd_node *nd = &s->t->nodes[ni];
    if (nd->cutdim < 0) {
        for (int i = nd->left; i < nd->right; i++) {
            float d = l2_simple(s->q, &s->t->pts[3 * i]);
            rs_insert(s->k, &s->cnt, s->idx, s->d2, d, s->t->ids[i]);
        }
        return;
    }
    int a = nd->cutdim;
    double val = s->q[a];
    double diff1 = val - (double)nd->divlow, diff2 = val - (double)nd->divhigh;
    int best, other; double cut;
    if (diff1 + diff2 < 0) { best = nd->left; other = nd->right; cut = diff2 * diff2; }
    else { best = nd->right; other = nd->left; cut = diff1 * diff1; }
    search_rec(s, best, mindist, dists);
    double saved = dists[a];
    double md = mindist + cut - saved;
    dists[a] = cut;
    /* "<=" keeps equal-distance candidates reachable so ties resolve by index;
     * the 1e-6 relative slack absorbs float rounding of the leaf distances */
    if (s->cnt < s->k || md * (1.0 - 1e-6) <= (double)s->d2[s->cnt - 1])
        search_rec(s, other, md, dists);
    dists[a] = saved;
}

int llo_kdtree_knn(const llo_kdtree *t, const float *q, int k, int *idx, float *d2)
{
    if (t->n <= 0) return 0;
    if (k > t->n) k = t->n;
    kd_search s = { t, q, k, 0, idx, d2 };
    double dists[3] = { 0, 0, 0 }, mind = 0;
    for (int a = 0; a < 3; a++) {
        double v = q[a];
        if (v < t->bbox[2 * a]) dists[a] = (v - t->bbox[2 * a]) * (v - t->bbox[2 * a]);
        else if (v > t->bbox[2 * a + 1]) dists[a] = (v - t->bbox[2 * a + 1]) * (v - t->bbox[2 * a + 1]);
        mind += dists[a];
    }
    search_rec(&s, 0, mind, dists);
    return s.cnt;
}
