/*
 * llo_voxel.c — ORACLE (test infrastructure): restatement of
 * pcl::VoxelGrid<pcl::PointXYZI>::applyFilter with setLeafSize(l,l,l) and every
 * other setting at its default, as the reference uses it:
 *   MO:249-251 (leaf 0.2 corner, 0.4 surf / outlier), MO:1058-1063 (local map),
 *   MO:1070-1089 (current scan, 4 filters), FA:779-780.
 * PCL is an un-vendored, un-pinned dependency (find_package(PCL),
 * /root/reference/LeGO-LOAM/CMakeLists.txt:24) and is not present offline, so
 * this follows the published PCL 1.7/1.8 algorithm:
 *   inverse_leaf = 1.0f/leaf; getMinMax3D; int64 overflow check -> pass-through;
 *   min_b = (int)floor(min*inv); ijk = (int)(floor(x*inv) - (float)min_b);
 *   idx = ijk . (1, div_x, div_x*div_y); sort by idx; one centroid per idx in
 *   ascending idx order; xyz AND intensity averaged with float sums and a true
 *   division by (float)count.
 * Definition chosen where PCL leaves it open: PCL's std::sort is unstable, so
 * the summation order inside a voxel is implementation-defined there; here it
 * is ascending input index (stable).  "parity unpinned" for this primitive.
 */
#include "llo.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int32_t idx; int32_t pt; } vox_pair;

/* (idx, input index) ascending = a STABLE sort by idx: LSD radix sort, 11-bit digits over the significant bits.
 * PCL sorts with std::sort (inlined comparator); a qsort with a function-pointer comparator was ~2x slower than that and
 * made the CPU baseline of the mapping cycle pessimistic (round-1 review) - the radix sort errs on the CPU's side. */
static void vox_sort(vox_pair *v, int n, int32_t max_idx)
{
    vox_pair *tmp = (vox_pair *)malloc(sizeof(vox_pair) * (size_t)n);
    vox_pair *src = v, *dst = tmp;
    for (int shift = 0; shift < 32 && ((int64_t)max_idx >> shift) != 0; shift += 11) {
        int cnt[2049];
        memset(cnt, 0, sizeof(cnt));
        for (int i = 0; i < n; i++) cnt[(((uint32_t)src[i].idx >> shift) & 2047u) + 1]++;
        for (int k = 0; k < 2048; k++) cnt[k + 1] += cnt[k];
        for (int i = 0; i < n; i++) dst[cnt[((uint32_t)src[i].idx >> shift) & 2047u]++] = src[i];
        vox_pair *t = src; src = dst; dst = t;
    }
    if (src != v) memcpy(v, src, sizeof(vox_pair) * (size_t)n);
    free(tmp);
}

/* points are read as base[i * stride + {0, 1, 2, off_i}] = x, y, z, intensity and written the same way: stride 4 / off 3
 * for llo_point, stride 8 / off 4 for pcl::PointXYZI's 32-byte layout (the tier-B shim filters the caller's clouds in
 * place, as PCL does, instead of converting them - two copies less per filter on the CPU baseline) */
static int voxel_grid_core(const float *in, int stride, int off_i, int n, float leaf, float *out, int ostride, int ooff_i,
                           int *overflow)
{
#define PX(i) in[(size_t)(i) * stride]
#define PY(i) in[(size_t)(i) * stride + 1]
#define PZ(i) in[(size_t)(i) * stride + 2]
#define PI(i) in[(size_t)(i) * stride + off_i]
    if (overflow) *overflow = 0;
    if (n <= 0) return 0;

    const float inv = 1.0f / leaf;
    float mn[3] = { PX(0), PY(0), PZ(0) }, mx[3] = { PX(0), PY(0), PZ(0) };
    for (int i = 1; i < n; i++) {
        const float p[3] = { PX(i), PY(i), PZ(i) };
        for (int a = 0; a < 3; a++) {
            if (p[a] < mn[a]) mn[a] = p[a];
            if (p[a] > mx[a]) mx[a] = p[a];
        }
    }
    int64_t d[3];
    for (int a = 0; a < 3; a++) d[a] = (int64_t)((mx[a] - mn[a]) * inv) + 1;
    if (d[0] * d[1] * d[2] > (int64_t)INT32_MAX) {
        for (int i = 0; i < n; i++) {
            float *o = out + (size_t)i * ostride;
            o[0] = PX(i); o[1] = PY(i); o[2] = PZ(i); o[ooff_i] = PI(i);
        }
        if (overflow) *overflow = 1;
        return n;
    }
    int min_b[3], max_b[3], div_b[3], mul[3];
    for (int a = 0; a < 3; a++) {
        min_b[a] = (int)floorf(mn[a] * inv);
        max_b[a] = (int)floorf(mx[a] * inv);
        div_b[a] = max_b[a] - min_b[a] + 1;
    }
    mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];

    vox_pair *v = (vox_pair *)malloc(sizeof(vox_pair) * (size_t)n);
    for (int i = 0; i < n; i++) {
        int ijk0 = (int)(floorf(PX(i) * inv) - (float)min_b[0]);
        int ijk1 = (int)(floorf(PY(i) * inv) - (float)min_b[1]);
        int ijk2 = (int)(floorf(PZ(i) * inv) - (float)min_b[2]);
        v[i].idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
        v[i].pt = i;
    }
    vox_sort(v, n, (int32_t)((int64_t)div_b[0] * div_b[1] * div_b[2] - 1));

    int m = 0, i = 0;
    while (i < n) {
        int j = i;
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
        while (j < n && v[j].idx == v[i].idx) {
            const int p = v[j].pt;
            sx += PX(p); sy += PY(p); sz += PZ(p); si += PI(p);
            j++;
        }
        float cnt = (float)(j - i);
        float *o = out + (size_t)m * ostride;
        o[0] = sx / cnt; o[1] = sy / cnt; o[2] = sz / cnt; o[ooff_i] = si / cnt;
        m++;
        i = j;
    }
    free(v);
    return m;
#undef PX
#undef PY
#undef PZ
#undef PI
}

int llo_voxel_grid(const llo_point *in, int n, float leaf, llo_point *out, int *overflow)
{
    return voxel_grid_core((const float *)in, 4, 3, n, leaf, (float *)out, 4, 3, overflow);
}

/* the same on clouds in pcl::PointXYZI's layout (8 floats per point: x y z _ intensity _ _ _); out must not alias in */
int llo_voxel_grid_pcl(const float *in32, int n, float leaf, float *out32, int *overflow)
{
    return voxel_grid_core(in32, 8, 4, n, leaf, out32, 8, 4, overflow);
}
