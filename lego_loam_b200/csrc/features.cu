// features.cu — see features.cuh.  Every statement cites the reference line it restates; arithmetic types follow the
// reference's C++ promotion rules (float members compared against double M_PI expressions, FA:506-519).
#include "features.cuh"
#include "std_sort.cuh"
#include "glibc_atan2f.cuh"
#include "glibc_sincosf.cuh"
#include <climits>
#include <cstring>
#include <cstddef>
#include <vector>
#include <algorithm>
#include <stdexcept>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace llb {

namespace {

constexpr int FE_TPB = 256;
constexpr int FE_RING_THREADS = 192;      // 6 warps: one per sector for the sorts
constexpr double FE_PI = 3.14159265358979323846;

__device__ __forceinline__ unsigned fe_col(const FeView &v, int i) { return (i >= 0 && i < v.n) ? __ldg(v.col + i) : 0u; }
__device__ __forceinline__ int fe_ground(const FeView &v, int i) { return (i >= 0 && i < v.n) ? (int)__ldg(v.ground + i) : 0; }
__device__ __forceinline__ int fe_col_diff(const FeView &v, int a, int b)
{   // std::abs(int(ColInd[a] - ColInd[b])) on uint32 members, FA:651
    return abs((int)(fe_col(v, a) - fe_col(v, b)));
}

// orientation of a point before the half-sweep test, FA:504-509
__device__ __forceinline__ float fe_ori_first_half(float ori, float start_ori)
{
    if ((double)ori < (double)start_ori - FE_PI / 2) ori = (float)((double)ori + 2 * FE_PI);
    else if ((double)ori > (double)start_ori + FE_PI * 3 / 2) ori = (float)((double)ori - 2 * FE_PI);
    return ori;
}

// every kernel of the step reads its sweep from a device table: blockIdx.y = sweep (one entry for a single context, one per
// slot for a batch)
__global__ void __launch_bounds__(FE_TPB) fe_point_kernel(const FeView *__restrict__ table)
{
    const FeView v = table[blockIdx.y];
    if (blockIdx.x == 0 && threadIdx.x < 9) v.hdr->pad[threadIdx.x] = 0;       // pad[2] + prof[7]
    // the record calculateSmoothness never rewrites (see fe_ring_kernel), as it is before this sweep's sorts
    if (blockIdx.x == 0 && threadIdx.x == 0) v.hdr->stale_ind = (int)(unsigned)v.smooth[4];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += gridDim.x * blockDim.x) {
        const float4 q = __ldg(v.cloud_in + i);
        // point.x = y, point.z = x; ori = -atan2(point.x, point.z) (float overload = glibc's atan2f, restated), FA:500-504
        const float ori = -glibcm::atan2f_(q.y, q.x);
        v.ori[i] = ori;
        const float a = fe_ori_first_half(ori, v.start_ori);
        if ((double)(a - v.start_ori) > FE_PI) atomicMin(&v.hdr->first_half, i);      // FA:511-512
        if (i >= 5 && i < v.n - 5) {                                                    // FA:624-640
            const float r0 = __ldg(v.range + i);
            float d = __ldg(v.range + i - 5) + __ldg(v.range + i - 4);
            d = d + __ldg(v.range + i - 3); d = d + __ldg(v.range + i - 2); d = d + __ldg(v.range + i - 1);
            d = d - r0 * 10;
            d = d + __ldg(v.range + i + 1); d = d + __ldg(v.range + i + 2); d = d + __ldg(v.range + i + 3);
            d = d + __ldg(v.range + i + 4); d = d + __ldg(v.range + i + 5);
            const float c = d * d;
            v.curv[i] = c; v.picked[i] = 0; v.label[i] = 0;
            v.smooth[i] = ((unsigned long long)__float_as_uint(c) << 32) | (unsigned)i;
        }
    }
}

__global__ void __launch_bounds__(FE_TPB) fe_mark_kernel(const FeView *__restrict__ table)
{
    const FeView v = table[blockIdx.y];
    const int half = v.hdr->first_half;       // the point at which halfPassed becomes true still takes the first branch
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += gridDim.x * blockDim.x) {
        const float4 q = __ldg(v.cloud_in + i);
        float ori = v.ori[i];
        if (i <= half) ori = fe_ori_first_half(ori, v.start_ori);
        else {                                                                          // FA:514-520
            ori = (float)((double)ori + 2 * FE_PI);
            if ((double)ori < (double)v.end_ori - FE_PI * 3 / 2) ori = (float)((double)ori + 2 * FE_PI);
            else if ((double)ori > (double)v.end_ori + FE_PI / 2) ori = (float)((double)ori - 2 * FE_PI);
        }
        const float rel = (ori - v.start_ori) / v.ori_diff;                            // FA:522
        float4 p;
        p.x = q.y; p.y = q.z; p.z = q.x;                                               // FA:500-502
        p.w = (float)(int)q.w + v.prm.scan_period * rel;                               // FA:523
        v.cloud_adj[i] = p;
        if (v.imu) v.ori[i] = rel;                                                      // relTime for fe_imu_kernel
        if (i >= 5 && i < v.n - 6) {                                                    // FA:647-677
            const float d1 = __ldg(v.range + i), d2 = __ldg(v.range + i + 1);
            if (fe_col_diff(v, i + 1, i) < 10) {
                if ((double)(d1 - d2) > 0.3) { for (int k = -5; k <= 0; k++) v.picked[i + k] = 1; }
                else if ((double)(d2 - d1) > 0.3) { for (int k = 1; k <= 6; k++) v.picked[i + k] = 1; }
            }
            const float f1 = fabsf(__ldg(v.range + i - 1) - d1), f2 = fabsf(d2 - d1);
            if ((double)f1 > 0.02 * (double)d1 && (double)f2 > 0.02 * (double)d1) v.picked[i] = 1;
        }
    }
}

// ---- the IMU branch of adjustDistortion (FA:525-613).  glibc's sinf / cosf restated (glibc_sincosf.cuh).
__device__ __forceinline__ float fe_cosf(float x) { return glibcm::cosf_(x); }
__device__ __forceinline__ float fe_sinf(float x) { return glibcm::sinf_(x); }

struct FeImuAt { float roll, pitch, yaw, velo[3], shift[3], ang[3]; };

// IMU state at timeScanCur + pointTime: the reference's pointer walk from imuPointerLastIteration (it restarts for every
// point, FA:527-533) and the linear interpolation between the two entries around that time (FA:535-565, FA:580-594)
__device__ __forceinline__ void fe_imu_at(const FeImu &m, float pointTime, FeImuAt &o)
{
    const double t = m.time_scan_cur + pointTime;
    int front = m.pointer_last_iteration;
    while (front != m.pointer_last) {
        if (t < m.time[front]) break;
        front = (front + 1) % FE_IMU_QUEUE;
    }
    if (t > m.time[front]) {
        o.roll = m.roll[front]; o.pitch = m.pitch[front]; o.yaw = m.yaw[front];
#pragma unroll
        for (int a = 0; a < 3; a++) { o.velo[a] = m.velo[a][front]; o.shift[a] = m.shift[a][front]; o.ang[a] = m.angular[a][front]; }
    } else {
        const int back = (front + FE_IMU_QUEUE - 1) % FE_IMU_QUEUE;
        const float rf = (float)((t - m.time[back]) / (m.time[front] - m.time[back]));
        const float rb = (float)((m.time[front] - m.time_scan_cur - pointTime) / (m.time[front] - m.time[back]));
        o.roll = m.roll[front] * rf + m.roll[back] * rb;
        o.pitch = m.pitch[front] * rf + m.pitch[back] * rb;
        if ((double)(m.yaw[front] - m.yaw[back]) > FE_PI)
            o.yaw = (float)((double)(m.yaw[front] * rf) + ((double)m.yaw[back] + 2 * FE_PI) * (double)rb);
        else if ((double)(m.yaw[front] - m.yaw[back]) < -FE_PI)
            o.yaw = (float)((double)(m.yaw[front] * rf) + ((double)m.yaw[back] - 2 * FE_PI) * (double)rb);
        else
            o.yaw = m.yaw[front] * rf + m.yaw[back] * rb;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            o.velo[a] = m.velo[a][front] * rf + m.velo[a][back] * rb;
            o.shift[a] = m.shift[a][front] * rf + m.shift[a][back] * rb;
            o.ang[a] = m.angular[a][front] * rf + m.angular[a][back] * rb;
        }
    }
}

__global__ void __launch_bounds__(FE_TPB) fe_imu_kernel(const FeView *__restrict__ table)
{
    const FeView v = table[blockIdx.y];
    if (!v.imu || v.n <= 0) return;
    const FeImu &m = *v.imu;
    if (m.pointer_last < 0) return;                                                    // FA:525
    // the first point defines the start state (FA:567-578); every thread derives it again (a few hundred instructions)
    FeImuAt st;
    fe_imu_at(m, v.ori[0] * v.prm.scan_period, st);
    const float cRs = fe_cosf(st.roll), cPs = fe_cosf(st.pitch), cYs = fe_cosf(st.yaw);      // updateImuRollPitchYawStartSinCos
    const float sRs = fe_sinf(st.roll), sPs = fe_sinf(st.pitch), sYs = fe_sinf(st.yaw);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += gridDim.x * blockDim.x) {
        FeImuOut *out = v.imu_out;
        if (i == 0) {
            out->start[0] = st.roll; out->start[1] = st.pitch; out->start[2] = st.yaw;
            for (int a = 0; a < 3; a++) { out->start[3 + a] = st.velo[a]; out->start[6 + a] = st.shift[a]; out->angular_cur[a] = st.ang[a]; }
            out->valid = 1;
            if (v.n == 1) { out->cur[0] = st.roll; out->cur[1] = st.pitch; out->cur[2] = st.yaw; out->has_velo = 0; }
            continue;
        }
        FeImuAt c;
        fe_imu_at(m, v.ori[i] * v.prm.scan_period, c);
        // VeloToStartIMU FA:346-363
        float vx = c.velo[0] - st.velo[0], vy = c.velo[1] - st.velo[1], vz = c.velo[2] - st.velo[2];
        {
            const float x1 = cYs * vx - sYs * vz, y1 = vy, z1 = sYs * vx + cYs * vz;
            const float x2 = x1, y2 = cPs * y1 + sPs * z1, z2 = -sPs * y1 + cPs * z1;
            vx = cRs * x2 + sRs * y2; vy = -sRs * x2 + cRs * y2; vz = z2;
        }
        // TransformToStartIMU FA:365-388 (imuShiftFromStart*Cur stay 0: ShiftToStartIMU is never called in the reference)
        float4 p = v.cloud_adj[i];
        {
            const float cr = fe_cosf(c.roll), sr = fe_sinf(c.roll), cp = fe_cosf(c.pitch), sp = fe_sinf(c.pitch);
            const float cy = fe_cosf(c.yaw), sy = fe_sinf(c.yaw);
            const float shx = 0.f, shy = 0.f, shz = 0.f;
            const float x1 = cr * p.x - sr * p.y, y1 = sr * p.x + cr * p.y, z1 = p.z;
            const float x2 = x1, y2 = cp * y1 - sp * z1, z2 = sp * y1 + cp * z1;
            const float x3 = cy * x2 + sy * z2, y3 = y2, z3 = -sy * x2 + cy * z2;
            const float x4 = cYs * x3 - sYs * z3, y4 = y3, z4 = sYs * x3 + cYs * z3;
            const float x5 = x4, y5 = cPs * y4 + sPs * z4, z5 = -sPs * y4 + cPs * z4;
            p.x = cRs * x5 + sRs * y5 + shx;
            p.y = -sRs * x5 + cRs * y5 + shy;
            p.z = z5 + shz;
        }
        v.cloud_adj[i] = p;
        if (i == v.n - 1) {
            out->cur[0] = c.roll; out->cur[1] = c.pitch; out->cur[2] = c.yaw;
            out->velo_from_start_cur[0] = vx; out->velo_from_start_cur[1] = vy; out->velo_from_start_cur[2] = vz;
            out->has_velo = 1;
        }
    }
}

// the ring's slice of the per-point arrays in shared memory; a stale record (see fe_ring_kernel) may name a point outside
// of it, which goes to global memory
struct RingWin {
    unsigned char *pk, *gr; unsigned short *col; float *curv;
    int w0, w1; volatile int *g_picked; int cap;
};
__device__ __forceinline__ int pk_get(const RingWin &w, int i)
{
    if (i >= w.w0 && i < w.w1) return w.pk[i - w.w0];
    return (i >= 0 && i < w.cap) ? w.g_picked[i] : 1;
}
__device__ __forceinline__ void pk_set(const RingWin &w, int i)
{
    if (i >= w.w0 && i < w.w1) w.pk[i - w.w0] = 1;
    else if (i >= 0 && i < w.cap) w.g_picked[i] = 1;
}
__device__ __forceinline__ int win_col_diff(const FeView &v, const RingWin &w, int a, int b)
{
    const unsigned ca = (a >= w.w0 && a < w.w1) ? (unsigned)w.col[a - w.w0] : fe_col(v, a);
    const unsigned cb = (b >= w.w0 && b < w.w1) ? (unsigned)w.col[b - w.w0] : fe_col(v, b);
    return abs((int)(ca - cb));
}
__device__ __forceinline__ float win_curv(const FeView &v, const RingWin &w, int i)
{
    return (i >= w.w0 && i < w.w1) ? w.curv[i - w.w0] : v.curv[i];
}
__device__ __forceinline__ int win_ground(const FeView &v, const RingWin &w, int i)
{
    return (i >= w.w0 && i < w.w1) ? (int)w.gr[i - w.w0] : fe_ground(v, i);
}
// FA:727-740 == FA:758-773: a pick at point i marks i, the forward steps i+1 .. i+nf and the backward steps i-1 .. i-nb,
// each direction up to 5 steps or the first column gap above 10.  nf | nb << 4 per point is a property of the sweep.
__device__ __forceinline__ int fe_reach(const FeView &v, const RingWin &w, int i)
{
    int nf = 0, nb = 0;
    while (nf < 5 && win_col_diff(v, w, i + nf + 1, i + nf) <= 10) nf++;
    while (nb < 5 && i - nb - 1 >= 0 && win_col_diff(v, w, i - nb - 1, i - nb) <= 10) nb++;   // i - nb - 1 < 0: see fe_ring_kernel
    return nf | (nb << 4);
}

// ---- libstdc++ std::sort by one warp (std_sort.cuh: closed-form partitions + independent leaf ranges; checked on the
// ---- host against libstdc++ in tests/test_host_std_sort.py).  All lanes call with identical arguments.
__device__ int warp_partition(stdsort::rec_t *a, int first, int last, int pivot, unsigned short *Lbuf, unsigned short *Rbuf, int lane)
{
    const stdsort::rec_t pv = a[pivot];
    const int m = last - first;
    const unsigned lt = (1u << lane) - 1;
    int nL = 0, nR = 0;
    for (int base = 0; base < m; base += 32) {
        const int i = base + lane;
        const bool fl = i < m && !stdsort::less(a[first + i], pv);              // left stoppers, increasing position
        const unsigned bl = __ballot_sync(FULL, fl);
        if (fl) Lbuf[nL + __popc(bl & lt)] = (unsigned short)i;
        nL += __popc(bl);
        const int j = m - 1 - i;
        const bool fr = i < m && !stdsort::less(pv, a[first + j]);              // right stoppers, decreasing position
        const unsigned br = __ballot_sync(FULL, fr);
        if (fr) Rbuf[nR + __popc(br & lt)] = (unsigned short)j;
        nR += __popc(br);
    }
    __syncwarp();
    const int mm = min(nL, nR);
    int K = 0;
    for (int base = 0; base < mm; base += 32) {
        const int k = base + lane;
        const unsigned b = __ballot_sync(FULL, k < mm && Lbuf[k] < Rbuf[k]);
        K += __popc(b);
        if (b != FULL) break;                                                   // the condition is monotone in k
    }
    for (int k = lane; k < K; k += 32) stdsort::swp(a + first + Lbuf[k], a + first + Rbuf[k]);
    int cut;
    if (K == 0) cut = first + Lbuf[0];
    else {
        const int rk = first + Rbuf[K - 1];
        cut = (K < nL && first + Lbuf[K] < rk) ? first + Lbuf[K] : rk;
    }
    __syncwarp();
    return cut;
}

__device__ void warp_std_sort(stdsort::rec_t *a, stdsort::rec_t *tmp, int n, unsigned short *Lbuf, unsigned short *Rbuf,
                              int lane, int *prof)
{
    if (n <= 1) return;
    const long long t0 = clock64();
    int depth = 0;
    for (int m = n; m > 1; m >>= 1) depth += 2;
    int stk_first[40], stk_last[40], stk_depth[40];       // <= 2 * log2(n) pending right parts
    int sp = 0, first = 0, last = n;
    for (;;) {
        while (last - first > 16) {
            if (depth == 0) {
                if (lane == 0) stdsort::heap_sort(a + first, last - first);
                for (int p = first + lane; p < last; p += 32) { Lbuf[p] = (unsigned short)p; Rbuf[p] = (unsigned short)(p + 1); }   // in order already
                __syncwarp();
                last = first;
                break;
            }
            --depth;
            if (lane == 0) stdsort::median_to_first(a + first, a + first + 1, a + first + (last - first) / 2, a + last - 1);
            __syncwarp();
            const int cut = warp_partition(a, first + 1, last, first, Lbuf + first + 1, Rbuf + first + 1, lane);
            stk_first[sp] = cut; stk_last[sp] = last; stk_depth[sp] = depth; sp++;
            last = cut;
        }
        // a terminal range: its records remember its bounds (the partition scratch of a range stays inside that range)
        for (int p = first + lane; p < last; p += 32) { Lbuf[p] = (unsigned short)first; Rbuf[p] = (unsigned short)last; }
        if (sp == 0) break;
        sp--;
        first = stk_first[sp]; last = stk_last[sp]; depth = stk_depth[sp];
    }
    __syncwarp();
    const long long t1 = clock64();
    // __final_insertion_sort = a stable sort of every terminal range on its own (std_sort.cuh): one record per lane, its
    // place inside the range by counting the records that go before it
    for (int base = 0; base < n; base += 32) {
        const int p = base + lane;
        if (p < n) {
            const stdsort::rec_t rec = a[p];
            const int s = Lbuf[p], e = Rbuf[p];
            int r = 0;
#pragma unroll
            for (int i = 0; i < 16; i++) {           // terminal ranges hold at most 16 records
                const int q = s + i;
                const stdsort::rec_t x = a[min(q, n - 1)];
                r += (q < e && (stdsort::less(x, rec) || (!stdsort::less(rec, x) && q < p))) ? 1 : 0;
            }
            const int dest = s + r;
            tmp[dest] = rec;
        }
    }
    __syncwarp();
    for (int p = lane; p < n; p += 32) a[p] = tmp[p];
    __syncwarp();
    if (lane == 0) { atomicMax(prof + 0, (int)(t1 - t0)); atomicMax(prof + 1, (int)(clock64() - t1)); }
}

__device__ __forceinline__ void fe_sector(int st, int en, int j, int &sp, int &ep)
{   // FA:693-694
    sp = (st * (6 - j) + en * j) / 6;
    ep = (st * (5 - j) + en * (j + 1)) / 6 - 1;
}

__global__ void __launch_bounds__(FE_RING_THREADS) fe_ring_kernel(const FeView *__restrict__ table)
{
    const FeView v = table[blockIdx.y];
    extern __shared__ __align__(16) unsigned long long s_rec[];                       // [horizon + 8]
    const int wcap = v.horizon + 32;
    unsigned long long *s_tmp = s_rec + v.horizon + 8;                                 // [horizon + 8]
    float *s_curv = reinterpret_cast<float *>(s_tmp + v.horizon + 8);                  // [wcap] each
    unsigned short *s_col = reinterpret_cast<unsigned short *>(s_curv + wcap);
    unsigned char *s_pk = reinterpret_cast<unsigned char *>(s_col + wcap);
    unsigned char *s_gr = s_pk + wcap;
    __shared__ int s_tot[FE_RING_THREADS / 32];
    __shared__ int s_sharp[FE_SHARP_PER_RING], s_lsharp[FE_LSHARP_PER_RING], s_flat[FE_FLAT_PER_RING], s_n[3];
    __shared__ int s_owner, s_late, s_walk[6];
    const int ring = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int st = __ldg(v.start_ring + ring), en = __ldg(v.end_ring + ring);
    // the record at position 4 is never rewritten by calculateSmoothness (FA:624 starts at 5) while sector 0 of the
    // first populated ring starts there (IP:318): it is state from earlier sweeps and may name a point of another
    // ring.  Only then (every ring can tell from the copy fe_point_kernel took of the record) the ring that sorts it
    // ("owner") finishes its picks before the other rings read their flags.
    if (tid == 0) {
        int owner = -1, late = 0;
        for (int r = 0; r < v.n_scan; r++) {
            const int rs = __ldg(v.start_ring + r), re = __ldg(v.end_ring + r);
            int sp, ep; fe_sector(rs, re, 0, sp, ep);
            if (sp == 4 && sp < ep) {
                owner = r;
                const int stale = v.hdr->stale_ind, ow0 = max(0, rs - 6), ow1 = max(ow0, min(v.cap, re + 7));
                late = (stale >= ow0 && stale < ow1) ? 0 : 1;
                break;
            }
        }
        s_owner = owner; s_late = late;
    }
    const int w0 = max(0, st - 6), w1 = max(w0, min(v.cap, en + 7));
    const int nrec = max(0, en - st);                    // positions [st, en - 1] = [sp_0, ep_5]
    for (int k = tid; k < nrec; k += FE_RING_THREADS) s_rec[k] = v.smooth[st + k];
    __syncthreads();
    const int owner = s_owner;
    const long long t_start = clock64();
    // ---- the six sorts, one warp each
    if (warp < 6) {      // scratch: the shared arrays of the ring slice, which are filled only after the sorts
        int sp, ep; fe_sector(st, en, warp, sp, ep);
        if (sp < ep)                                                                     // FA:699: [sp, ep)
            warp_std_sort(s_rec + (sp - st), s_tmp + (sp - st), ep - sp, s_col + (sp - st), reinterpret_cast<unsigned short *>(s_curv) + (sp - st),
                          lane, v.hdr->prof);
    }
    if (owner >= 0 && owner != ring && s_late && tid == 0) {
        volatile int *flag = &v.hdr->release_seq;
        while (*flag != v.seq) __nanosleep(100);
        __threadfence();
    }
    __syncthreads();
    const long long t_sorted = clock64();
    for (int k = w0 + tid; k < w1; k += FE_RING_THREADS) {
        s_pk[k - w0] = (unsigned char)(((volatile int *)v.picked)[k] != 0);
        s_curv[k - w0] = v.curv[k];
        s_col[k - w0] = (unsigned short)fe_col(v, k);
        s_gr[k - w0] = (unsigned char)fe_ground(v, k);
    }
    __syncthreads();
    RingWin w{ s_pk, s_gr, s_col, s_curv, w0, w1, (volatile int *)v.picked, v.cap };
    unsigned char *s_reach = reinterpret_cast<unsigned char *>(s_tmp);        // [wcap]; the sort scratch is free again
    for (int k = w0 + tid; k < w1; k += FE_RING_THREADS) s_reach[k - w0] = (unsigned char)fe_reach(v, w, k);
    // a sector without non-ground (ground) points has no edge (flat) candidates: its loop is not walked at all.  The
    // sector that holds the stale record is always walked (the record may name a point outside the sector).
    if (warp < 6) {
        int sp, ep; fe_sector(st, en, warp, sp, ep);
        bool any_g = false, any_ng = false;
        for (int k = sp + lane; k <= ep && sp < ep; k += 32) { const bool g = s_gr[k - w0] != 0; any_g |= g; any_ng |= !g; }
        any_g = __any_sync(FULL, any_g); any_ng = __any_sync(FULL, any_ng);
        if (lane == 0) s_walk[warp] = (owner == ring && warp == 0) ? 3 : ((any_ng ? 1 : 0) | (any_g ? 2 : 0));
    }
    __syncthreads();
    // ---- greedy picks, sectors in order (a pick of sector j may block neighbours that belong to sector j+1)
    if (warp == 0) {
        int nsharp = 0, nls = 0, nflat = 0;
        long long c_large = 0, c_flat = 0;
        for (int j = 0; j < 6; j++) {
            int sp, ep; fe_sector(st, en, j, sp, ep);
            if (sp >= ep) continue;
            int cnt = 0;
            const long long tl0 = clock64();
            const int walk = s_walk[j];
            for (int khi = ep; khi >= sp && cnt < 20 && (walk & 1); khi -= 32) {         // FA:701-742
                const int k = khi - lane;
                const bool valid = k >= sp;
                const int ind = valid ? (int)(unsigned)s_rec[k - st] : w0;
                bool stat = valid && win_curv(v, w, ind) > v.prm.edge_threshold && win_ground(v, w, ind) == 0;
                // flags set by earlier chunks / sectors come from shared memory once; inside the chunk every lane follows
                // the marks of the picks in registers
                bool el = stat && pk_get(w, ind) == 0;
                const int my_reach = !stat ? 0 : (ind >= w0 && ind < w1) ? (int)s_reach[ind - w0] : fe_reach(v, w, ind);
                for (;;) {
                    const unsigned b = __ballot_sync(FULL, el);
                    if (!b) break;
                    const int f = __ffs(b) - 1;
                    const int pind = __shfl_sync(FULL, ind, f), pr = __shfl_sync(FULL, my_reach, f);
                    const int nf = pr & 15, nb = pr >> 4;
                    if (lane == f) {      // labels and the points themselves are written after the picks, by all threads
                        if (cnt < 2) s_sharp[nsharp] = ind;
                        s_lsharp[nls] = ind;
                    }
                    if (lane <= nf + nb) pk_set(w, pind - nb + lane);
                    if (cnt < 2) nsharp++;
                    nls++; cnt++;
                    if (cnt >= 20) break;
                    if (lane <= f || (ind >= pind - nb && ind <= pind + nf)) el = false;
                }
                __syncwarp();
            }
            cnt = 0;
            const long long tl1 = clock64();
            c_large += tl1 - tl0;
            bool done = false;
            for (int klo = sp; klo <= ep && !done && (walk & 2); klo += 32) {            // FA:744-775
                const int k = klo + lane;
                const bool valid = k <= ep;
                const int ind = valid ? (int)(unsigned)s_rec[k - st] : w0;
                bool stat = valid && win_curv(v, w, ind) < v.prm.surf_threshold && win_ground(v, w, ind) != 0;
                bool el = stat && pk_get(w, ind) == 0;
                const int my_reach = !stat ? 0 : (ind >= w0 && ind < w1) ? (int)s_reach[ind - w0] : fe_reach(v, w, ind);
                for (;;) {
                    const unsigned b = __ballot_sync(FULL, el);
                    if (!b) break;
                    const int f = __ffs(b) - 1;
                    const int pind = __shfl_sync(FULL, ind, f), pr = __shfl_sync(FULL, my_reach, f);
                    const int nf = pr & 15, nb = pr >> 4;
                    if (lane == f) s_flat[nflat] = ind;
                    nflat++; cnt++;
                    if (cnt >= 4) { done = true; break; }                 // the 4th pick breaks before the marks, FA:752-756
                    if (lane <= nf + nb) pk_set(w, pind - nb + lane);
                    if (lane <= f || (ind >= pind - nb && ind <= pind + nf)) el = false;
                }
                __syncwarp();
            }
            c_flat += clock64() - tl1;
        }
        if (lane == 0) { atomicMax(v.hdr->prof + 2, (int)c_large); atomicMax(v.hdr->prof + 3, (int)c_flat); }
        if (lane == 0) {
            v.r_cnt[ring * 4 + 0] = nsharp; v.r_cnt[ring * 4 + 1] = nls; v.r_cnt[ring * 4 + 2] = nflat;
            s_n[0] = nsharp; s_n[1] = nls; s_n[2] = nflat;
        }
        __threadfence();
        __syncwarp();
        if (lane == 0 && owner == ring && s_late) atomicExch(&v.hdr->release_seq, v.seq);
        if (lane == 0) {        // slowest ring: cycles of the sort phase (incl. the wait for the owner ring) and of the picks
            atomicMax(&v.hdr->pad[0], (int)(t_sorted - t_start)); atomicMax(&v.hdr->pad[1], (int)(clock64() - t_sorted));
        }
    }
    __syncthreads();
    // cloudLabel: 1 for every edge pick, then 2 for the first two of each sector, -1 for the flat picks (FA:708-750)
    for (int i = tid; i < s_n[1]; i += FE_RING_THREADS) v.label[s_lsharp[i]] = 1;
    __syncthreads();
    for (int i = tid; i < s_n[0]; i += FE_RING_THREADS) v.label[s_sharp[i]] = 2;
    for (int i = tid; i < s_n[2]; i += FE_RING_THREADS) v.label[s_flat[i]] = -1;
    for (int i = tid; i < s_n[0]; i += FE_RING_THREADS) v.r_sharp[ring * FE_SHARP_PER_RING + i] = v.cloud_adj[s_sharp[i]];
    for (int i = tid; i < s_n[1]; i += FE_RING_THREADS) v.r_lsharp[ring * FE_LSHARP_PER_RING + i] = v.cloud_adj[s_lsharp[i]];
    for (int i = tid; i < s_n[2]; i += FE_RING_THREADS) v.r_flat[ring * FE_FLAT_PER_RING + i] = v.cloud_adj[s_flat[i]];
    // state back: flags set by this ring, records in sorted order
    for (int k = w0 + tid; k < w1; k += FE_RING_THREADS) if (s_pk[k - w0]) v.picked[k] = 1;
    for (int k = tid; k < nrec; k += FE_RING_THREADS) v.smooth[st + k] = s_rec[k];
    // ---- surfPointsLessFlatScan: points with label <= 0 of the sectors that were processed, in index order, FA:777-781
    __syncthreads();                 // labels of this ring are in place
    int base = 0;
    for (int j = 0; j < 6; j++) {
        int sp, ep; fe_sector(st, en, j, sp, ep);
        if (sp >= ep) continue;
        for (int k0 = sp; k0 <= ep; k0 += FE_RING_THREADS) {
            const int k = k0 + tid;
            const bool flag = k <= ep && v.label[k] <= 0;
            const unsigned b = __ballot_sync(FULL, flag);
            if (lane == 0) s_tot[warp] = __popc(b);
            __syncthreads();
            int off = 0, tot = 0;
#pragma unroll
            for (int q = 0; q < FE_RING_THREADS / 32; q++) { if (q < warp) off += s_tot[q]; tot += s_tot[q]; }
            if (flag) v.r_lf_scan[(size_t)ring * v.horizon + base + off + __popc(b & ((1u << lane) - 1))] = v.cloud_adj[k];
            base += tot;
            __syncthreads();
        }
    }
    if (tid == 0) v.r_cnt[ring * 4 + 3] = base;
}

__global__ void __launch_bounds__(1024) fe_concat_kernel(const FeView *__restrict__ table)
{
    const FeView v = table[blockIdx.y];
    __shared__ int s_off[4][FE_MAX_RINGS + 1];
    const int tid = threadIdx.x;
    __shared__ int s_c[4][FE_MAX_RINGS];
    for (int i = tid; i < 4 * v.n_scan; i += blockDim.x) {
        const int which = i & 3, r = i >> 2;
        s_c[which][r] = which < 3 ? v.r_cnt[r * 4 + which] : v.r_lf_ds_cnt[r];
    }
    __syncthreads();
    if (tid < 4) {
        int acc = 0;
        for (int r = 0; r < v.n_scan; r++) { s_off[tid][r] = acc; acc += s_c[tid][r]; }
        s_off[tid][v.n_scan] = acc;
        v.hdr->counts[tid] = acc;
    }
    __syncthreads();
    const float4 *src[4] = { v.r_sharp, v.r_lsharp, v.r_flat, v.r_lf_ds };
    const int stride[4] = { FE_SHARP_PER_RING, FE_LSHARP_PER_RING, FE_FLAT_PER_RING, v.horizon };
    for (int which = 0; which < 4; which++) {
        const int total = s_off[which][v.n_scan];
        for (int i = tid; i < total; i += blockDim.x) {
            int lo = 0, hi = v.n_scan;                    // last ring with offset <= i
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[which][mid] <= i) lo = mid; else hi = mid; }
            v.out[which][i] = src[which][(size_t)lo * stride[which] + (i - s_off[which][lo])];
        }
    }
    __syncthreads();
    if (tid == 0) {
        *v.out_hdr = *v.hdr;                               // counts + cycle counters travel with the clouds: one D2H
        v.hdr->first_half = INT_MAX;                       // ready for the next sweep
    }
}

struct FeEndJob { const float4 *in[2]; float4 *out[2]; int n[2]; float T[6]; FeEndImu imu; };

// TransformToEnd FA:885-953 (glibc's sinf / cosf restated: the clouds come out as the reference's own, bit for bit).  Without
// IMU messages every IMU angle and shift is 0; the factors stay in the expressions so that signed zeros come out alike.
__global__ void __launch_bounds__(FE_TPB) fe_to_end_kernel(FeEndJob jb)
{
    const float *T = jb.T;
    const float cosImuRollStart = jb.imu.cs_start[0], sinImuRollStart = jb.imu.cs_start[1];
    const float cosImuPitchStart = jb.imu.cs_start[2], sinImuPitchStart = jb.imu.cs_start[3];
    const float cosImuYawStart = jb.imu.cs_start[4], sinImuYawStart = jb.imu.cs_start[5];
    const float shX = jb.imu.shift_from_start[0], shY = jb.imu.shift_from_start[1], shZ = jb.imu.shift_from_start[2];
    const float cYawL = fe_cosf(jb.imu.last[2]), sYawL = fe_sinf(jb.imu.last[2]), cPitchL = fe_cosf(jb.imu.last[1]), sPitchL = fe_sinf(jb.imu.last[1]);
    const float cRollL = fe_cosf(jb.imu.last[0]), sRollL = fe_sinf(jb.imu.last[0]);
    const float cRy = fe_cosf(T[1]), sRy = fe_sinf(T[1]), cRx = fe_cosf(T[0]), sRx = fe_sinf(T[0]);
    const float cRz = fe_cosf(T[2]), sRz = fe_sinf(T[2]);
    const int total = jb.n[0] + jb.n[1];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int which = i < jb.n[0] ? 0 : 1, k = which ? i - jb.n[0] : i;
        const float4 pi = jb.in[which][k];
        const float s = 10 * (pi.w - (int)pi.w);
        float rx = s * T[0], ry = s * T[1], rz = s * T[2], tx = s * T[3], ty = s * T[4], tz = s * T[5];
        const float crz = fe_cosf(rz), srz = fe_sinf(rz), crx = fe_cosf(rx), srx = fe_sinf(rx), cry = fe_cosf(ry), sry = fe_sinf(ry);
        const float x1 = crz * (pi.x - tx) + srz * (pi.y - ty);
        const float y1 = -srz * (pi.x - tx) + crz * (pi.y - ty);
        const float z1 = (pi.z - tz);
        const float x2 = x1;
        const float y2 = crx * y1 + srx * z1;
        const float z2 = -srx * y1 + crx * z1;
        const float x3 = cry * x2 - sry * z2;
        const float y3 = y2;
        const float z3 = sry * x2 + cry * z2;
        tx = T[3]; ty = T[4]; tz = T[5];
        const float x4 = cRy * x3 + sRy * z3;
        const float y4 = y3;
        const float z4 = -sRy * x3 + cRy * z3;
        const float x5 = x4;
        const float y5 = cRx * y4 - sRx * z4;
        const float z5 = sRx * y4 + cRx * z4;
        const float x6 = cRz * x5 - sRz * y5 + tx;
        const float y6 = sRz * x5 + cRz * y5 + ty;
        const float z6 = z5 + tz;
        const float x7 = cosImuRollStart * (x6 - shX) - sinImuRollStart * (y6 - shY);
        const float y7 = sinImuRollStart * (x6 - shX) + cosImuRollStart * (y6 - shY);
        const float z7 = z6 - shZ;
        const float x8 = x7;
        const float y8 = cosImuPitchStart * y7 - sinImuPitchStart * z7;
        const float z8 = sinImuPitchStart * y7 + cosImuPitchStart * z7;
        const float x9 = cosImuYawStart * x8 + sinImuYawStart * z8;
        const float y9 = y8;
        const float z9 = -sinImuYawStart * x8 + cosImuYawStart * z8;
        const float x10 = cYawL * x9 - sYawL * z9;
        const float y10 = y9;
        const float z10 = sYawL * x9 + cYawL * z9;
        const float x11 = x10;
        const float y11 = cPitchL * y10 + sPitchL * z10;
        const float z11 = -sPitchL * y10 + cPitchL * z10;
        float4 po;
        po.x = cRollL * x11 + sRollL * y11;
        po.y = -sRollL * x11 + cRollL * y11;
        po.z = z11;
        po.w = (float)(int)pi.w;
        jb.out[which][k] = po;
    }
}

size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
// pcl::PointXYZI (32 B: x y z _ intensity _ _ _) -> float4 {x, y, z, intensity}; this repacking is the host cost of a sweep
void fe_pack_cloud(const float *cloud32, int n, float *out16)
{
#if defined(__SSE2__)
    for (int i = 0; i < n; i++) {
        const float *q = cloud32 + 8 * (size_t)i;
        const __m128 a = _mm_loadu_ps(q), b = _mm_load_ss(q + 4);
        const __m128 t = _mm_shuffle_ps(a, b, _MM_SHUFFLE(0, 0, 2, 2));             // z z i i
        _mm_store_ps(out16 + 4 * (size_t)i, _mm_shuffle_ps(a, t, _MM_SHUFFLE(2, 0, 1, 0)));   // x y z i
    }
#else
    for (int i = 0; i < n; i++) {
        const float *q = cloud32 + 8 * (size_t)i;
        out16[4 * i] = q[0]; out16[4 * i + 1] = q[1]; out16[4 * i + 2] = q[2]; out16[4 * i + 3] = q[4];
    }
#endif
}

size_t fe_input_bytes(int n, int n_scan)
{   // must match the layout of FeatureExtractor::stage
    const size_t o_cloud = FE_HEAD, o_range = align16(o_cloud + sizeof(float4) * n), o_col = align16(o_range + 4 * (size_t)n),
                 o_start = align16(o_col + 4 * (size_t)n), o_end = align16(o_start + 4 * (size_t)n_scan),
                 o_ground = align16(o_end + 4 * (size_t)n_scan);
    return align16(o_ground + (size_t)n + 16);
}

}  // namespace

size_t FeatureExtractor::input_bytes(int n, int n_scan) { return fe_input_bytes(n, n_scan); }

size_t FeatureExtractor::out_block_bytes(int n_scan, int horizon)
{
    return 64 + sizeof(float4) * (size_t)n_scan * (FE_SHARP_PER_RING + FE_LSHARP_PER_RING + FE_FLAT_PER_RING)
         + sizeof(float4) * (size_t)n_scan * horizon;
}

void FeatureExtractor::init(int n_scan, int horizon, cudaStream_t s, unsigned char *out_dev_ext)
{
    release();
    n_scan_ = n_scan; horizon_ = horizon; cap_ = n_scan * horizon;
    const size_t cap = (size_t)cap_;
    cloud_adj_.ensure(cap); ori_.ensure(cap); curv_.ensure(cap); picked_.ensure(cap); label_.ensure(cap); smooth_.ensure(cap);
    r_sharp_.ensure((size_t)n_scan * FE_SHARP_PER_RING); r_lsharp_.ensure((size_t)n_scan * FE_LSHARP_PER_RING);
    r_flat_.ensure((size_t)n_scan * FE_FLAT_PER_RING); r_lf_scan_.ensure(cap); r_lf_ds_.ensure(cap);
    r_cnt_.ensure((size_t)n_scan * 4); r_lf_ds_cnt_.ensure(n_scan); hdr_.ensure(1); jobs_.ensure(n_scan);
    // the reference's arrays start as whatever `new` returns (FA:210-212); zero, as the oracle harness defines them
    LLB_CUDA(cudaMemsetAsync(curv_.p, 0, sizeof(float) * cap, s));
    LLB_CUDA(cudaMemsetAsync(picked_.p, 0, sizeof(int) * cap, s));
    LLB_CUDA(cudaMemsetAsync(label_.p, 0, sizeof(int) * cap, s));
    LLB_CUDA(cudaMemsetAsync(smooth_.p, 0, sizeof(unsigned long long) * cap, s));       // FA:223: {0, 0}
    FeHeader h{}; h.first_half = INT_MAX; h.release_seq = 0;
    LLB_CUDA(cudaMemcpyAsync(hdr_.p, &h, sizeof(h), cudaMemcpyHostToDevice, s));
    jobs_host_.assign(n_scan, SmallJob{});
    std::vector<SmallJob> &jobs = jobs_host_;
    for (int r = 0; r < n_scan; r++) {
        SmallJob &j = jobs[r];
        j.in.a = r_lf_scan_.p + (size_t)r * horizon; j.in.na_dev = r_cnt_.p + r * 4 + 3; j.in.na = horizon;
        j.in.b = nullptr; j.in.nb_dev = nullptr; j.in.nb = 0;
        j.leaf = prm.leaf; j.out = r_lf_ds_.p + (size_t)r * horizon; j.n_out = r_lf_ds_cnt_.p + r;
    }
    LLB_CUDA(cudaMemcpyAsync(jobs_.p, jobs.data(), sizeof(SmallJob) * n_scan, cudaMemcpyHostToDevice, s));
    LLB_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < 2; k++) LLB_CUDA(cudaEventCreateWithFlags(&in_ev_[k], cudaEventDisableTiming));
    out_off_[0] = 64;
    out_off_[1] = out_off_[0] + sizeof(float4) * n_scan * FE_SHARP_PER_RING;
    out_off_[2] = out_off_[1] + sizeof(float4) * n_scan * FE_LSHARP_PER_RING;
    out_off_[3] = out_off_[2] + sizeof(float4) * n_scan * FE_FLAT_PER_RING;
    if (out_dev_ext) { out_dev_ = out_dev_ext; out_pin_ = nullptr; }
    else {
        pin_out_.ensure(out_off_[3] + sizeof(float4) * cap); out_block_.ensure(out_off_[3] + sizeof(float4) * cap);
        out_dev_ = out_block_.p; out_pin_ = pin_out_.p;
    }
    const int smem = (horizon + 8) * 16 + (horizon + 32) * 8;
    if (smem > 48 * 1024)
        LLB_CUDA(cudaFuncSetAttribute(fe_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    seq_ = 0; n_ = 0;
}

void FeatureExtractor::release()
{
    for (int k = 0; k < 2; k++) { if (in_ev_[k]) cudaEventDestroy(in_ev_[k]); in_ev_[k] = nullptr; in_busy_[k] = false; pin_in_[k].release(); }
    in_dev_.release(); cloud_adj_.release(); r_sharp_.release(); r_lsharp_.release(); r_flat_.release(); r_lf_scan_.release();
    r_lf_ds_.release(); out_block_.release();
    ori_.release(); curv_.release(); picked_.release(); label_.release(); r_cnt_.release(); r_lf_ds_cnt_.release();
    smooth_.release(); hdr_.release(); jobs_.release(); pin_out_.release();
    n_scan_ = 0;
}

const float4 *FeatureExtractor::host_cloud(int which) const
{
    return reinterpret_cast<const float4 *>(out_pin_ + out_off_[which]);
}

// the sweep goes into one pinned block: [FeView, FE_HEAD B][cloud as float4][range][column][ring bounds][ground flags];
// copy_in() sends it with one H2D, the view at its head is this sweep's entry of the kernels' table
FeView FeatureExtractor::stage(const float *cloud32, int n, const int *start_ring, const int *end_ring, float start_ori,
                               float end_ori, float ori_diff, const unsigned char *ground, const unsigned *col,
                               const float *range, unsigned char *hp_ext, unsigned char *dp_ext)
{
    static_assert(sizeof(FeView) <= FE_HEAD, "FeView must fit the head of the input block");
    const size_t o_cloud = FE_HEAD, o_range = align16(o_cloud + sizeof(float4) * n), o_col = align16(o_range + 4 * (size_t)n),
                 o_start = align16(o_col + 4 * (size_t)n), o_end = align16(o_start + 4 * (size_t)n_scan_),
                 o_ground = align16(o_end + 4 * (size_t)n_scan_), total = align16(o_ground + (size_t)n + 16);
    int rb = 0;
    unsigned char *hp = hp_ext, *dp = dp_ext;
    if (!hp_ext) {
        rb = ring_; ring_ ^= 1;
        if (in_busy_[rb]) { LLB_CUDA(cudaEventSynchronize(in_ev_[rb])); in_busy_[rb] = false; }
        pin_in_[rb].ensure(total); in_dev_.ensure(total);
        hp = pin_in_[rb].p; dp = in_dev_.p;
    }
    float *hc = reinterpret_cast<float *>(hp + o_cloud);
    fe_pack_cloud(cloud32, n, hc);
    std::memcpy(hp + o_range, range, 4 * (size_t)n);
    std::memcpy(hp + o_col, col, 4 * (size_t)n);
    std::memcpy(hp + o_start, start_ring, 4 * (size_t)n_scan_);
    std::memcpy(hp + o_end, end_ring, 4 * (size_t)n_scan_);
    std::memcpy(hp + o_ground, ground, (size_t)n);

    FeView v{};
    v.cloud_in = reinterpret_cast<const float4 *>(dp + o_cloud); v.cloud_adj = cloud_adj_.p;
    v.n = n; v.n_scan = n_scan_; v.horizon = horizon_; v.cap = cap_;
    v.start_ring = reinterpret_cast<const int *>(dp + o_start); v.end_ring = reinterpret_cast<const int *>(dp + o_end);
    v.ground = dp + o_ground; v.col = reinterpret_cast<const unsigned *>(dp + o_col);
    v.range = reinterpret_cast<const float *>(dp + o_range);
    v.start_ori = start_ori; v.end_ori = end_ori; v.ori_diff = ori_diff;
    v.ori = ori_.p; v.curv = curv_.p; v.picked = picked_.p; v.label = label_.p; v.smooth = smooth_.p; v.hdr = hdr_.p;
    v.r_sharp = r_sharp_.p; v.r_lsharp = r_lsharp_.p; v.r_flat = r_flat_.p; v.r_lf_scan = r_lf_scan_.p; v.r_lf_ds = r_lf_ds_.p;
    v.r_cnt = r_cnt_.p; v.r_lf_ds_cnt = r_lf_ds_cnt_.p;
    for (int k = 0; k < 4; k++) v.out[k] = reinterpret_cast<float4 *>(out_dev_ + out_off_[k]);
    v.out_hdr = reinterpret_cast<FeHeader *>(out_dev_);
    v.prm = prm; v.seq = ++seq_;
    v.imu = imu_set_ ? imu_dev_.p : nullptr; v.imu_out = imu_out_dev_.p;
    std::memcpy(hp, &v, sizeof(v));
    n_ = n; staged_ = rb; staged_bytes_ = total;
    return v;
}

void FeatureExtractor::set_imu(const FeImu *imu_host)
{
    imu_dev_.ensure(1); pin_imu_.ensure(1); imu_out_dev_.ensure(1); pin_imu_out_.ensure(1);
    imu_set_ = imu_host != nullptr && imu_host->pointer_last >= 0;
    if (imu_set_) *pin_imu_.p = *imu_host;                   // (the previous sweep's copy has long been sent)
}

void FeatureExtractor::copy_in(cudaStream_t s)
{
    if (imu_set_) LLB_CUDA(cudaMemcpyAsync(imu_dev_.p, pin_imu_.p, sizeof(FeImu), cudaMemcpyHostToDevice, s));
    if (imu_out_dev_.p) LLB_CUDA(cudaMemsetAsync(imu_out_dev_.p, 0, sizeof(FeImuOut), s));
    LLB_CUDA(cudaMemcpyAsync(in_dev_.p, pin_in_[staged_].p, staged_bytes_, cudaMemcpyHostToDevice, s));
    LLB_CUDA(cudaEventRecord(in_ev_[staged_], s)); in_busy_[staged_] = true;
}

void FeatureExtractor::copy_out(cudaStream_t s)
{   // header (counts) + the four clouds in one block, one D2H (the less-flat cloud is bounded by n)
    LLB_CUDA(cudaMemcpyAsync(out_pin_, out_dev_, out_bytes_used(n_), cudaMemcpyDeviceToHost, s));
    if (imu_out_dev_.p) LLB_CUDA(cudaMemcpyAsync(pin_imu_out_.p, imu_out_dev_.p, sizeof(FeImuOut), cudaMemcpyDeviceToHost, s));
}

int FeatureExtractor::launch(const FeView *table_dev, int count, int n_max, int n_scan, int horizon, const SmallJob *jobs_dev,
                             cudaStream_t s, bool imu)
{
    const int gx = std::max(1, std::min(div_up(n_max, FE_TPB), std::max(8, 148 * 4 / std::max(count, 1))));
    fe_point_kernel<<<dim3(gx, count), FE_TPB, 0, s>>>(table_dev);
    fe_mark_kernel<<<dim3(gx, count), FE_TPB, 0, s>>>(table_dev);
    if (imu) fe_imu_kernel<<<dim3(gx, count), FE_TPB, 0, s>>>(table_dev);
    fe_ring_kernel<<<dim3(n_scan, count), FE_RING_THREADS, (horizon + 8) * 16 + (horizon + 32) * 8, s>>>(table_dev);
    launch_voxel_cta_jobs(jobs_dev, n_scan * count, (horizon + 1023) & ~1023, s);
    fe_concat_kernel<<<dim3(1, count), 1024, 0, s>>>(table_dev);
    LLB_CUDA(cudaGetLastError());
    return imu ? 6 : 5;
}

int FeatureExtractor::extract(const float *cloud32, int n, const int *start_ring, const int *end_ring, float start_ori,
                              float end_ori, float ori_diff, const unsigned char *ground, const unsigned *col,
                              const float *range, cudaStream_t s)
{
    stage(cloud32, n, start_ring, end_ring, start_ori, end_ori, ori_diff, ground, col, range);
    copy_in(s);
    const int launches = launch(reinterpret_cast<const FeView *>(in_dev_.p), 1, n, n_scan_, horizon_, jobs_.p, s, imu_set_);
    copy_out(s);
    return launches;
}

int FeatureExtractor::extract_dev(const float4 *cloud_dev, int n, const int *start_ring_dev, const int *end_ring_dev, float start_ori,
                                  float end_ori, float ori_diff, const unsigned char *ground_dev, const unsigned *col_dev,
                                  const float *range_dev, cudaStream_t s)
{
    const int rb = ring_; ring_ ^= 1;
    if (in_busy_[rb]) { LLB_CUDA(cudaEventSynchronize(in_ev_[rb])); in_busy_[rb] = false; }
    pin_in_[rb].ensure(FE_HEAD); in_dev_.ensure(FE_HEAD);
    FeView v{};
    v.cloud_in = cloud_dev; v.cloud_adj = cloud_adj_.p;
    v.n = n; v.n_scan = n_scan_; v.horizon = horizon_; v.cap = cap_;
    v.start_ring = start_ring_dev; v.end_ring = end_ring_dev;
    v.ground = ground_dev; v.col = col_dev; v.range = range_dev;
    v.start_ori = start_ori; v.end_ori = end_ori; v.ori_diff = ori_diff;
    v.ori = ori_.p; v.curv = curv_.p; v.picked = picked_.p; v.label = label_.p; v.smooth = smooth_.p; v.hdr = hdr_.p;
    v.r_sharp = r_sharp_.p; v.r_lsharp = r_lsharp_.p; v.r_flat = r_flat_.p; v.r_lf_scan = r_lf_scan_.p; v.r_lf_ds = r_lf_ds_.p;
    v.r_cnt = r_cnt_.p; v.r_lf_ds_cnt = r_lf_ds_cnt_.p;
    for (int k = 0; k < 4; k++) v.out[k] = reinterpret_cast<float4 *>(out_dev_ + out_off_[k]);
    v.out_hdr = reinterpret_cast<FeHeader *>(out_dev_);
    v.prm = prm; v.seq = ++seq_;
    v.imu = imu_set_ ? imu_dev_.p : nullptr; v.imu_out = imu_out_dev_.p;
    std::memcpy(pin_in_[rb].p, &v, sizeof(v));
    n_ = n; staged_ = rb; staged_bytes_ = FE_HEAD;
    copy_in(s);
    const int launches = launch(reinterpret_cast<const FeView *>(in_dev_.p), 1, n, n_scan_, horizon_, jobs_.p, s, imu_set_);
    copy_out(s);
    return launches;
}

// ---------------------------------------------------------------- batch: one sweep per slot, the same five launches
void FeatureBatch::init(int slots, int n_scan, int horizon, cudaStream_t s)
{
    release();
    out_stride_ = (FeatureExtractor::out_block_bytes(n_scan, horizon) + 255) & ~(size_t)255;
    out_dev_.ensure(out_stride_ * slots); pin_out_.ensure(out_stride_ * slots);
    ext_.resize(slots);
    std::vector<SmallJob> jobs;
    for (int i = 0; i < slots; i++) {
        ext_[i].init(n_scan, horizon, s, out_dev_.p + out_stride_ * i);
        jobs.insert(jobs.end(), ext_[i].jobs_host().begin(), ext_[i].jobs_host().end());
    }
    jobs_.ensure(jobs.size());
    LLB_CUDA(cudaMemcpyAsync(jobs_.p, jobs.data(), sizeof(SmallJob) * jobs.size(), cudaMemcpyHostToDevice, s));
    LLB_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < 2; k++) LLB_CUDA(cudaEventCreateWithFlags(&in_ev_[k], cudaEventDisableTiming));
    n_scan_ = n_scan; horizon_ = horizon;
}

void FeatureBatch::release()
{
    for (auto &e : ext_) e.release();
    ext_.clear(); jobs_.release(); in_dev_.release(); out_dev_.release(); pin_out_.release();
    for (int k = 0; k < 2; k++) { pin_in_[k].release(); if (in_ev_[k]) cudaEventDestroy(in_ev_[k]); in_ev_[k] = nullptr; in_busy_[k] = false; }
    n_scan_ = 0;
}

int FeatureBatch::extract(const FeSweepHost *sweeps, cudaStream_t s)
{
    const int S = (int)ext_.size();
    // ---- one input block: the table of the sweeps at its head, then every slot's sweep; one H2D
    std::vector<size_t> off(S + 1);
    off[0] = (sizeof(FeView) * (size_t)S + 255) & ~(size_t)255;
    int n_max = 1;
    for (int i = 0; i < S; i++) {
        off[i + 1] = off[i] + ((FeatureExtractor::input_bytes(sweeps[i].n, n_scan_) + 255) & ~(size_t)255);
        n_max = std::max(n_max, sweeps[i].n);
    }
    const int rb = ring_; ring_ ^= 1;
    if (in_busy_[rb]) { LLB_CUDA(cudaEventSynchronize(in_ev_[rb])); in_busy_[rb] = false; }
    pin_in_[rb].ensure(off[S]); in_dev_.ensure(off[S]);
    unsigned char *hp = pin_in_[rb].p, *dp = in_dev_.p;
    FeView *table = reinterpret_cast<FeView *>(hp);
    size_t sent = off[0];                   // the table at the head goes last
    for (int i = 0; i < S; i++) {           // host work: the repacking of the sweep; the copies of 8 slots at a time overlap it
        const FeSweepHost &h = sweeps[i];
        table[i] = ext_[i].stage(h.cloud32, h.n, h.start_ring, h.end_ring, h.start_ori, h.end_ori, h.ori_diff, h.ground, h.col,
                                 h.range, hp + off[i], dp + off[i]);
        if ((i & 7) == 7 || i == S - 1) {
            LLB_CUDA(cudaMemcpyAsync(dp + sent, hp + sent, off[i + 1] - sent, cudaMemcpyHostToDevice, s));
            sent = off[i + 1];
        }
    }
    LLB_CUDA(cudaMemcpyAsync(dp, hp, off[0], cudaMemcpyHostToDevice, s));
    LLB_CUDA(cudaEventRecord(in_ev_[rb], s)); in_busy_[rb] = true;
    const int launches = FeatureExtractor::launch(reinterpret_cast<const FeView *>(dp), S, n_max, n_scan_, horizon_, jobs_.p, s);
    // ---- one 2-D D2H: the used part (header, three fixed clouds, n_max less-flat points) of every slot's result block
    const size_t used = ext_[0].out_bytes_used(n_max);
    LLB_CUDA(cudaMemcpy2DAsync(pin_out_.p, used, out_dev_.p, out_stride_, used, S, cudaMemcpyDeviceToHost, s));
    for (int i = 0; i < S; i++) ext_[i].set_host_out(pin_out_.p + used * i);
    return launches;
}

int FeatureExtractor::transform_to_end(const float T[6], float4 *corner_out, float4 *surf_out, cudaStream_t s, const FeEndImu *imu)
{
    FeEndJob jb;
    if (imu) jb.imu = *imu;
    else {                                                   // a node that never received an IMU message: cos(0), sin(0), zeros
        for (int k = 0; k < 6; k++) jb.imu.cs_start[k] = (k & 1) ? sinf(0.f) : cosf(0.f);
        for (int k = 0; k < 3; k++) { jb.imu.shift_from_start[k] = 0.f; jb.imu.last[k] = 0.f; }
    }
    jb.in[0] = dev_cloud(1); jb.in[1] = dev_cloud(3); jb.out[0] = corner_out; jb.out[1] = surf_out;
    jb.n[0] = counts()[1]; jb.n[1] = counts()[3];
    for (int i = 0; i < 6; i++) jb.T[i] = T[i];
    const int total = jb.n[0] + jb.n[1];
    if (total <= 0) return 0;
    fe_to_end_kernel<<<std::min(div_up(total, FE_TPB), 148 * 4), FE_TPB, 0, s>>>(jb);
    LLB_CUDA(cudaGetLastError());
    return 1;
}

void FeatureExtractor::get_state(float *curv, int *picked, int *label, int n, cudaStream_t s)
{
    LLB_CUDA(cudaMemcpyAsync(curv, curv_.p, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
    LLB_CUDA(cudaMemcpyAsync(picked, picked_.p, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    LLB_CUDA(cudaMemcpyAsync(label, label_.p, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    LLB_CUDA(cudaStreamSynchronize(s));
}

}  // namespace llb
