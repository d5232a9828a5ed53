// projection.cu — K8 kernels: imageProjection (IP = LeGO-LOAM/src/imageProjection.cpp) on the device, see projection.cuh.
// Compiled with -fmad=false: every float expression keeps the reference's operation order and its C++ promotions
// (float members against double M_PI expressions); atan2 on float arguments is glibc's atan2f restated.
#include "projection.cuh"
#include "glibc_atan2f.cuh"
#include <cstring>
#include <cmath>

namespace llb {

namespace {

constexpr int IP_TPB = 256;
constexpr double IP_PI = 3.14159265358979323846;

// angle in degrees as the reference forms it: atan2(float, float) * 180 / M_PI assigned to a float
__device__ __forceinline__ float atan2_deg(float y, float x)
{
    return (float)((double)(glibcm::atan2f_(y, x) * 180.0f) / IP_PI);
}

// projectPointCloud IP:213-257: the pixel of every point; the last point of the cloud in a pixel wins
__global__ void __launch_bounds__(IP_TPB)
ip_project_kernel(IpView v)
{
    const int i = blockIdx.x * IP_TPB + threadIdx.x;
    if (i >= v.n) return;
    const int N = v.prm.n_scan, H = v.prm.horizon;
    const float x = __ldg(&v.cloud32[8 * (size_t)i]), y = __ldg(&v.cloud32[8 * (size_t)i + 1]), z = __ldg(&v.cloud32[8 * (size_t)i + 2]);
    const int row = (int)__ldg(&v.ring[i]);                 // useCloudRing (UT:60)
    if (row >= N) return;
    const float ha = atan2_deg(x, y);
    const double cd = -round(((double)ha - 90.0) / (double)v.prm.ang_res_x) + (double)(H / 2);
    if (!(cd >= 0.0)) return;                                // (cannot happen for finite points: cd is in [H/4, 5H/4])
    long long col = (long long)cd;
    if (col >= H) col -= H;
    if (col >= H) return;
    const float range = sqrtf(x * x + y * y + z * z);
    if (range < v.prm.sensor_min_range) return;
    atomicMax(&v.winner[row * H + (int)col], i);
}

// rangeMat / fullCloud of every pixel from its winner (and the winner table goes back to "none" for the next sweep)
__global__ void __launch_bounds__(IP_TPB)
ip_image_kernel(IpView v)
{
    const int N = v.prm.n_scan, H = v.prm.horizon;
    const int p = blockIdx.x * IP_TPB + threadIdx.x;
    if (p >= N * H) return;
    const int w = v.winner[p];
    if (w < 0) {
        const float nan = __int_as_float(0x7fc00000);
        v.full[p] = make_float4(nan, nan, nan, -1.f);        // resetParameters IP:144-157
        v.range_mat[p] = FLT_MAX;
        return;
    }
    v.winner[p] = -1;
    const float x = __ldg(&v.cloud32[8 * (size_t)w]), y = __ldg(&v.cloud32[8 * (size_t)w + 1]), z = __ldg(&v.cloud32[8 * (size_t)w + 2]);
    const int row = p / H, col = p - row * H;
    v.range_mat[p] = sqrtf(x * x + y * y + z * z);
    v.full[p] = make_float4(x, y, z, (float)((double)(float)row + (double)(float)col / 10000.0));   // IP:250
}

// groundRemoval IP:259-310, one thread per column (the reference's sweep over the rows, statement for statement),
// then the initial labels IP:301-307 and the per-pixel state of the segmentation
__global__ void __launch_bounds__(IP_TPB)
ip_ground_kernel(IpView v)
{
    const int N = v.prm.n_scan, H = v.prm.horizon;
    const int j = blockIdx.x * IP_TPB + threadIdx.x;
    if (j >= H) return;
    for (int i = 0; i < N; i++) v.ground_mat[i * H + j] = 0;
    for (int i = 0; i < v.prm.ground_scan_ind; ++i) {
        const int lo = j + i * H, up = j + (i + 1) * H;
        const float4 a = v.full[lo], b = v.full[up];
        if (a.w == -1.f || b.w == -1.f) { v.ground_mat[lo] = -1; continue; }
        const float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
        const float angle = atan2_deg(dz, sqrtf(dx * dx + dy * dy));
        if (fabsf(angle - v.prm.sensor_mount_angle) <= 10) { v.ground_mat[lo] = 1; v.ground_mat[up] = 1; }
    }
    for (int i = 0; i < N; i++) {
        const int p = i * H + j;
        v.label_mat[p] = (v.ground_mat[p] == 1 || v.range_mat[p] == FLT_MAX) ? -1 : 0;
        v.parent[p] = p; v.cnt[p] = 0; v.number[p] = 0;
        reinterpret_cast<uint4 *>(v.rowmask)[p] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// the edge test of labelComponents IP:408-416 (symmetric in the two pixels)
__device__ __forceinline__ bool ip_connected(float a, float b, float s, float c, float theta)
{
    const float d1 = fmaxf(a, b), d2 = fminf(a, b);
    return glibcm::atan2f_(d2 * s, d1 - d2 * c) > theta;
}

// lock-free union-find, smaller index = root (ECL-CC style hooking with intermediate pointer jumping)
__device__ __forceinline__ int uf_find(int *parent_, int x)
{
    volatile int *parent = parent_;                          // other threads hook and compress concurrently
    int cur = parent[x];
    if (cur != x) {
        int next, prev = x;
        while (cur > (next = parent[cur])) { parent[prev] = next; prev = cur; cur = next; }
    }
    return cur;
}

__device__ __forceinline__ void uf_unite(int *parent, int a, int b)
{
    int ra = uf_find(parent, a), rb = uf_find(parent, b);
    bool repeat;
    do {
        repeat = false;
        if (ra != rb) {
            int ret;
            if (ra < rb) { if ((ret = atomicCAS(&parent[rb], rb, ra)) != rb) { rb = ret; repeat = true; } }
            else { if ((ret = atomicCAS(&parent[ra], ra, rb)) != ra) { ra = ret; repeat = true; } }
        }
    } while (repeat);
}

__global__ void __launch_bounds__(IP_TPB)
ip_union_kernel(IpView v)
{
    const int N = v.prm.n_scan, H = v.prm.horizon;
    const int p = blockIdx.x * IP_TPB + threadIdx.x;
    if (p >= N * H || v.label_mat[p] != 0) return;
    const int row = p / H, col = p - row * H;
    const float a = v.range_mat[p];
    const int q = row * H + (col + 1 == H ? 0 : col + 1);   // columns wrap (IP:397-400)
    if (q != p && v.label_mat[q] == 0 && ip_connected(a, v.range_mat[q], v.prm.sin_ax, v.prm.cos_ax, v.prm.segment_theta))
        uf_unite(v.parent, p, q);
    if (row + 1 < N) {
        const int d = p + H;
        if (v.label_mat[d] == 0 && ip_connected(a, v.range_mat[d], v.prm.sin_ay, v.prm.cos_ay, v.prm.segment_theta))
            uf_unite(v.parent, p, d);
    }
}

// every candidate pixel learns its root (= the BFS seed of its segment); per root: segment size and the rows of the
// points the BFS would have PUSHED (all but the seed itself, IP:420-425)
__global__ void __launch_bounds__(IP_TPB)
ip_flatten_kernel(IpView v)
{
    const int N = v.prm.n_scan, H = v.prm.horizon;
    const int p = blockIdx.x * IP_TPB + threadIdx.x;
    if (p >= N * H || v.label_mat[p] != 0) return;
    const int r = uf_find(v.parent, p);
    v.root[p] = r;                                           // (not parent[p]: another thread's path compression may still
    atomicAdd(&v.cnt[r], 1);                                 //  overwrite that with an ancestor that is not the root)
    if (p != r) { const int row = p / H; atomicOr(&v.rowmask[4 * (size_t)r + (row >> 5)], 1u << (row & 31)); }
}

__device__ __forceinline__ bool ip_feasible(const IpView &v, int root)
{
    const int c = v.cnt[root];                               // allPushedIndSize IP:429-441
    if (c >= 30) return true;
    if (c < v.prm.valid_point_num) return false;
    const uint4 m = reinterpret_cast<const uint4 *>(v.rowmask)[root];
    return __popc(m.x) + __popc(m.y) + __popc(m.z) + __popc(m.w) >= v.prm.valid_line_num;
}

// label numbers (labelCount IP:443-448), the final labels, then cloudSegmentation IP:312-368 as keep flags + scans +
// scatter in raster order, the ring bounds and findStartEndAngle IP:199-211.  ONE CTA: three scans over N*H pixels.
__global__ void __launch_bounds__(1024)
ip_finalize_kernel(IpView v)
{
    __shared__ int s_scan[33];
    const int N = v.prm.n_scan, H = v.prm.horizon, NP = N * H;
    const int tid = threadIdx.x;
    const int per = (NP + 1023) / 1024;
    const int lo = min(tid * per, NP), hi = min(lo + per, NP);
    // ---- 1: kept seeds in raster order -> label numbers
    int s = 0;
    for (int p = lo; p < hi; p++) s += (v.label_mat[p] == 0 && v.root[p] == p && ip_feasible(v, p)) ? 1 : 0;
    int total_labels;
    int run = block_excl_scan(s, s_scan, total_labels);
    for (int p = lo; p < hi; p++)
        if (v.label_mat[p] == 0 && v.root[p] == p && ip_feasible(v, p)) v.number[p] = ++run;
    __syncthreads();
    // ---- 2: final labels + keep flags
    int ks = 0, ko = 0;
    for (int p = lo; p < hi; p++) {
        int lab = v.label_mat[p];
        if (lab == 0) { const int r = v.root[p]; lab = v.number[r] > 0 ? v.number[r] : 999999; }
        v.label_mat[p] = lab;
        const bool gnd = v.ground_mat[p] == 1;
        const int row = p / H, col = p - row * H;
        if (lab > 0 || gnd) {
            if (lab == 999999) { if (row > v.prm.ground_scan_ind && col % 5 == 0) ko++; }
            else if (!(gnd && col % 5 != 0 && col > 5 && col < H - 5)) ks++;
        }
    }
    int tot_s, tot_o;
    int pos_s = block_excl_scan(ks, s_scan, tot_s);
    int pos_o = block_excl_scan(ko, s_scan, tot_o);
    // ---- 3: scatter
    for (int p = lo; p < hi; p++) {
        const int row = p / H, col = p - row * H;
        if (col == 0) {                                      // IP:318 / IP:358
            v.start_ring[row] = pos_s - 1 + 5;
            if (row > 0) v.end_ring[row - 1] = pos_s - 1 - 5;
        }
        const int lab = v.label_mat[p];
        const bool gnd = v.ground_mat[p] == 1;
        if (lab > 0 || gnd) {
            if (lab == 999999) {
                if (row > v.prm.ground_scan_ind && col % 5 == 0) v.outlier[pos_o++] = v.full[p];
                continue;
            }
            if (gnd && col % 5 != 0 && col > 5 && col < H - 5) continue;
            v.ground_flag[pos_s] = gnd ? 1 : 0;
            v.col_ind[pos_s] = (unsigned)col;
            v.seg_range[pos_s] = v.range_mat[p];
            v.seg[pos_s] = v.full[p];
            pos_s++;
        }
    }
    if (tid == 0) {
        v.end_ring[N - 1] = tot_s - 1 - 5;
        IpHeader h;
        h.n_seg = tot_s; h.n_outlier = tot_o; h.n_labels = total_labels; h.pad = 0; h.pad2 = 0.f;
        h.start_ori = h.end_ori = h.ori_diff = 0.f;
        if (v.n > 0) {                                       // findStartEndAngle IP:199-211
            const float x0 = v.cloud32[0], y0 = v.cloud32[1];
            const float xl = v.cloud32[8 * (size_t)(v.n - 1)], yl = v.cloud32[8 * (size_t)(v.n - 1) + 1];
            float so = -glibcm::atan2f_(y0, x0);
            float eo = (float)((double)(-glibcm::atan2f_(yl, xl)) + 2 * IP_PI);
            if ((double)(eo - so) > 3 * IP_PI) eo = (float)((double)eo - 2 * IP_PI);
            else if ((double)(eo - so) < IP_PI) eo = (float)((double)eo + 2 * IP_PI);
            h.start_ori = so; h.end_ori = eo; h.ori_diff = eo - so;
        }
        *v.hdr = h;
    }
}

__global__ void ip_fill_int_kernel(int *p, int n, int val)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = val;
}

}  // namespace

void ImageProjector::init(int n_scan, int horizon, float ang_res_x, float ang_res_y, int ground_scan_ind, cudaStream_t s)
{
    release();
    prm_.n_scan = n_scan; prm_.horizon = horizon; prm_.ground_scan_ind = ground_scan_ind;
    prm_.ang_res_x = ang_res_x; prm_.ang_res_y = ang_res_y;
    prm_.sensor_min_range = 1.0f; prm_.sensor_mount_angle = 0.0f;             // UT:111-112
    prm_.segment_theta = (float)(60.0 / 180.0 * M_PI);                        // UT:113
    prm_.valid_point_num = 5; prm_.valid_line_num = 3;                        // UT:114-115
    const float ax = (float)(ang_res_x / 180.0 * M_PI), ay = (float)(ang_res_y / 180.0 * M_PI);   // UT:116-117
    prm_.sin_ax = sinf(ax); prm_.cos_ax = cosf(ax); prm_.sin_ay = sinf(ay); prm_.cos_ay = cosf(ay);   // host libm = the reference's
    cap_ = n_scan * horizon;
    const size_t np = (size_t)cap_;
    winner_.ensure(np); label_mat_.ensure(np); parent_.ensure(np); root_.ensure(np); cnt_.ensure(np); number_.ensure(np);
    start_ring_.ensure(n_scan); end_ring_.ensure(n_scan);
    rowmask_.ensure(4 * np); col_ind_.ensure(np);
    full_.ensure(np); seg_.ensure(np); outlier_.ensure(np);
    range_mat_.ensure(np); seg_range_.ensure(np); ground_mat_.ensure(np); ground_flag_.ensure(np);
    hdr_.ensure(1); pin_hdr_.ensure(1); pin_rings_.ensure(2 * (size_t)n_scan);
    std::memset(pin_hdr_.p, 0, sizeof(IpHeader));
    ip_fill_int_kernel<<<148, 256, 0, s>>>(winner_.p, cap_, -1);
    LLB_CUDA(cudaGetLastError());
    for (int k = 0; k < 2; k++) LLB_CUDA(cudaEventCreateWithFlags(&in_ev_[k], cudaEventDisableTiming));
    LLB_CUDA(cudaStreamSynchronize(s));
}

void ImageProjector::release()
{
    winner_.release(); label_mat_.release(); parent_.release(); root_.release(); cnt_.release(); number_.release(); start_ring_.release();
    end_ring_.release(); rowmask_.release(); col_ind_.release(); full_.release(); seg_.release(); outlier_.release();
    range_mat_.release(); seg_range_.release(); ground_mat_.release(); ground_flag_.release(); hdr_.release(); pin_hdr_.release(); pin_rings_.release();
    in_dev_.release();
    for (int k = 0; k < 2; k++) { pin_in_[k].release(); if (in_ev_[k]) cudaEventDestroy(in_ev_[k]); in_ev_[k] = nullptr; in_busy_[k] = false; }
    prm_ = IpParams{};
    cap_ = 0;
}

int ImageProjector::process(const float *cloud32_host, const unsigned short *ring_host, int n, cudaStream_t s)
{
    // one pinned block: [cloud, 32 B per point][ring, 2 B per point]; one H2D
    const size_t o_ring = ((size_t)n * 32 + 255) & ~(size_t)255, total = o_ring + (((size_t)n * 2 + 255) & ~(size_t)255) + 256;
    const int rb = ring_pos_; ring_pos_ ^= 1;
    if (in_busy_[rb]) { LLB_CUDA(cudaEventSynchronize(in_ev_[rb])); in_busy_[rb] = false; }
    pin_in_[rb].ensure(total); in_dev_.ensure(total);
    if (n > 0) {
        std::memcpy(pin_in_[rb].p, cloud32_host, (size_t)n * 32);
        std::memcpy(pin_in_[rb].p + o_ring, ring_host, (size_t)n * 2);
        LLB_CUDA(cudaMemcpyAsync(in_dev_.p, pin_in_[rb].p, o_ring + (size_t)n * 2, cudaMemcpyHostToDevice, s));
        LLB_CUDA(cudaEventRecord(in_ev_[rb], s)); in_busy_[rb] = true;
    }
    IpView v{};
    v.prm = prm_;
    v.cloud32 = reinterpret_cast<const float *>(in_dev_.p); v.ring = reinterpret_cast<const unsigned short *>(in_dev_.p + o_ring);
    v.n = n;
    v.winner = winner_.p; v.full = full_.p; v.range_mat = range_mat_.p; v.ground_mat = ground_mat_.p; v.label_mat = label_mat_.p;
    v.parent = parent_.p; v.root = root_.p; v.cnt = cnt_.p; v.rowmask = rowmask_.p; v.number = number_.p;
    v.seg = seg_.p; v.outlier = outlier_.p; v.ground_flag = ground_flag_.p; v.col_ind = col_ind_.p; v.seg_range = seg_range_.p;
    v.start_ring = start_ring_.p; v.end_ring = end_ring_.p; v.hdr = hdr_.p;
    const int gp = div_up(cap_, IP_TPB);
    int launches = 0;
    if (n > 0) { ip_project_kernel<<<div_up(n, IP_TPB), IP_TPB, 0, s>>>(v); launches++; }
    ip_image_kernel<<<gp, IP_TPB, 0, s>>>(v);
    ip_ground_kernel<<<div_up(prm_.horizon, IP_TPB), IP_TPB, 0, s>>>(v);
    ip_union_kernel<<<gp, IP_TPB, 0, s>>>(v);
    ip_flatten_kernel<<<gp, IP_TPB, 0, s>>>(v);
    ip_finalize_kernel<<<1, 1024, 0, s>>>(v);
    launches += 5;
    LLB_CUDA(cudaGetLastError());
    LLB_CUDA(cudaMemcpyAsync(pin_hdr_.p, hdr_.p, sizeof(IpHeader), cudaMemcpyDeviceToHost, s));
    LLB_CUDA(cudaMemcpyAsync(pin_rings_.p, start_ring_.p, sizeof(int) * prm_.n_scan, cudaMemcpyDeviceToHost, s));
    LLB_CUDA(cudaMemcpyAsync(pin_rings_.p + prm_.n_scan, end_ring_.p, sizeof(int) * prm_.n_scan, cudaMemcpyDeviceToHost, s));
    return launches;
}

}  // namespace llb
