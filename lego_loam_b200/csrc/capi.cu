// capi.cu — the C ABI of include/llb200.h: context, host<->device staging and the
// orchestration of K1 (voxel), K2 (grid index), K3+K4 (scan-to-map iteration) and K5
// (odometry).  There is no CPU fallback anywhere in this file: without a CUDA device
// llb_create fails with LLB_ERR_NO_DEVICE.
#include "../../include/llb200.h"
#include "common.cuh"
#include "voxel.cuh"
#include "grid_index.cuh"
#include "s2m.cuh"
#include "odom.cuh"
#include "keyframes.cuh"
#include "features.cuh"
#include "projection.cuh"
#include "shard.cuh"
#include "loop.cuh"

#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>

using namespace llb;

namespace {

__global__ void unpack_points_kernel(const float *__restrict__ src32, int n, float4 *__restrict__ dst)
{
    // src: pcl::PointXYZI stride (8 floats): x y z w intensity c1 c2 c3
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(src32) + 2 * i);
        const float inten = __ldg(src32 + 8 * i + 4);
        dst[i] = make_float4(a.x, a.y, a.z, inten);
    }
}

struct Cloud {                    // device cloud with a device-resident length
    DevBuf<float4> pts;
    const float4 *view = nullptr; // where the points are: pts.p (uploaded from the host) or the caller's device memory
    int n_host = 0;               // exact length when known on the host, else upper bound
    bool exact = true;
};

}  // namespace

struct llb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    llb_params prm{};
    std::string err;
    long long launches = 0;

    // staging: the caller's 32 B-stride clouds are DMA'd as they are and compacted on the device
    PinnedBuf<float> pin_in[3];
    DevBuf<float> raw32[3];
    cudaEvent_t pin_ev[3] = { nullptr, nullptr, nullptr };
    bool pin_busy[3] = { false, false, false };
    struct Reg { const void *p; size_t bytes; bool ours; };
    std::vector<Reg> regs;        // caller buffers page-locked by this context (pin_host_clouds)
    bool async_pending = false;   // llb_s2m_optimize_async issued, llb_s2m_result not yet called
    DevBuf<float4> tmp_in;
    PinnedBuf<float4> pin_out;
    PinnedBuf<int> pin_counts;
    PinnedBuf<S2mState> pin_state;
    PinnedBuf<OdomState> pin_ostate;

    DevBuf<int> counts;           // device-resident lengths, see enum below
    enum { C_CORNER_DS = 0, C_SURF_DS, C_OUTLIER_DS, C_SURFTOTAL_DS, C_MAP_CORNER_DS, C_MAP_SURF_DS, C_VOX_TMP,
           C_LOOP_LATEST, C_LOOP_HIST_DS, C_GLOBAL_DS, C_N };

    VoxelFilter vox;
    VoxelFilter vox2;             // second set of voxel scratch: the two map filters MO:1057-1064 run concurrently
    cudaStream_t stream2 = nullptr;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    cudaEvent_t ds_fork_ev = nullptr, ds_ev = nullptr;   // downsampleCurrentScan on the forked stream (small sweeps)
    bool ds_pending = false;
    // scan side (MO:109-118)
    Cloud cornerLast, surfLast, outlierLast;
    DevBuf<float4> cornerLastDS, surfLastDS, outlierLastDS, surfTotalLastDS;
    bool scan_set = false, scan_ds_done = false;
    // map side (MO:123-126)
    Cloud mapCornerRaw, mapSurfRaw;
    DevBuf<float4> mapCornerDS, mapSurfDS;
    const float4 *mapCornerDS_view = nullptr, *mapSurfDS_view = nullptr;   // own buffer or caller's device memory
    int mapCornerDS_upper = 0, mapSurfDS_upper = 0;
    bool map_counts_on_dev = false;
    GridIndex gridCorner, gridSurf;
    bool map_set = false;

    S2mSolver s2m;
    DevBuf<float4> dbg_coeff; DevBuf<int> dbg_valid; DevBuf<int> dbg_knn; DevBuf<float> dbg_d2;
    int dbg_nc = 0, dbg_ns = 0;
    bool dbg_ready = false;
    DevBuf<float4> tmp_vox;

    OdomSolver odom;
    FeatureExtractor features;
    ImageProjector projection;
    bool projection_done = false;
    bool features_done = false;
    float features_ms = 0.f;
    int features_last_n[2] = { -1, -1 };   // sizes of laserCloudCornerLast / laserCloudSurfLast set by llb_features_publish_last

    // device-resident key-frame store + assembled raw local map
    KeyFrameStore kfs;
    DevBuf<float4> asmCorner, asmSurf;
    DevBuf<AsmSeg> asm_segs;
    PinnedBuf<AsmSeg> pin_segs;
    cudaEvent_t asm_ev = nullptr;
    bool asm_busy = false;
    int asm_rc = 0, asm_rs = 0;

    // loop closure + global map (loop.cuh)
    IcpSolver icp;
    DevBuf<float4> loopLatestRaw, loopLatest, loopHistRaw, loopHistDS, globalRaw, globalDS;
    VoxelFilter vox3;              // scratch of the history / global-map filters (kept apart from the map filters' scratch)
    PinnedBuf<IcpState> pin_icp;
    int loop_n_latest = -1, loop_n_hist_raw = 0, loop_n_hist = -1, global_n = -1, loop_n_latest_raw = 0;

    // sharded local map (BASELINE config 4, shard.cuh): plan, this rank's part of the two raw maps, scratch
    ShardPlan shard{};
    DevBuf<float4> shardCorner, shardSurf;
    DevBuf<int> shard_blk, shard_cnt;      // compaction scratch; {kept corner, kept surf, owned corner DS, owned surf DS}
    DevBuf<float> shard_samp;
    PinnedBuf<float> shard_samp_pin;
    int shard_kept[2] = { 0, 0 }, shard_local[2] = { 0, 0 }, shard_owned[2] = { 0, 0 }, shard_global[2] = { -1, -1 };

    // fused multi-GPU exchange (sharded registration): this rank's mailbox + the peers' mailboxes mapped through cudaIpc
    unsigned char *p2p_mem = nullptr;
    S2mPeers peers{};
    void *p2p_opened[S2M_MAX_PEERS] = {};
    bool p2p_ready = false;
};

namespace {

S2mParams s2m_params(const llb_params &p)
{
    S2mParams q;
    q.knn_max_sqdist = p.knn_max_sqdist; q.min_corr = p.s2m_min_correspondences;
    q.degeneracy_thresh = p.s2m_degeneracy_thresh; q.converge_deg = p.s2m_converge_deg;
    q.converge_cm = p.s2m_converge_cm; q.corner_map_min = p.corner_map_min; q.surf_map_min = p.surf_map_min;
    q.max_ctas = p.s2m_max_ctas;
    return q;
}

OdomParams odom_params(const llb_params &p)
{
    OdomParams q;
    q.nearest_sqdist = p.odom_nearest_sqdist; q.max_iter = p.odom_max_iterations;
    q.min_corr = p.odom_min_correspondences; q.degeneracy_thresh = p.odom_degeneracy_thresh;
    q.converge_deg = p.odom_converge_deg; q.converge_cm = p.odom_converge_cm;
    return q;
}

template <typename F>
int guarded(llb_ctx *ctx, F &&f)
{
    if (!ctx) return LLB_ERR_INVALID;
    try {
        cudaError_t e = cudaSetDevice(ctx->device);
        if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return LLB_ERR_CUDA; }
        return f();
    } catch (const CudaError &e) {
        ctx->err = e.what();
        return LLB_ERR_CUDA;
    } catch (const std::exception &e) {
        ctx->err = e.what();
        return LLB_ERR_INVALID;
    }
}

// page-lock the caller's buffer once (pin_host_clouds); false => fall back to the staging copy
bool ensure_registered(llb_ctx *c, const void *p, size_t bytes)
{
    for (auto &r : c->regs)
        if (r.p == p && r.bytes >= bytes) return true;
    for (size_t i = 0; i < c->regs.size(); i++)
        if (c->regs[i].p == p) {                              // same buffer grew: register the larger range
            if (c->regs[i].ours) { cudaStreamSynchronize(c->stream); cudaHostUnregister(const_cast<void *>(p)); }
            c->regs.erase(c->regs.begin() + i);
            break;
        }
    if (c->regs.size() >= 64) {                               // bounded cache: drop the oldest entry
        if (c->regs[0].ours) { cudaStreamSynchronize(c->stream); cudaHostUnregister(const_cast<void *>(c->regs[0].p)); }
        c->regs.erase(c->regs.begin());
    }
    cudaError_t e = cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {          // another context of this process pinned it
        cudaGetLastError();
        c->regs.push_back({ p, bytes, false });
        return true;
    }
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    c->regs.push_back({ p, bytes, true });
    return true;
}

// host (32 B stride) -> device float4.  slot selects the staging pair.
void upload_cloud(llb_ctx *c, int slot, const llb_point *src, int n, DevBuf<float4> &dst)
{
    dst.ensure(std::max(n, 1));
    if (n <= 0) return;
    static_assert(sizeof(llb_point) == 32, "pcl::PointXYZI layout");
    if (c->prm.pin_host_clouds && ensure_registered(c, src, (size_t)n * sizeof(llb_point))) {
        c->raw32[slot].ensure((size_t)n * 8);
        LLB_CUDA(cudaMemcpyAsync(c->raw32[slot].p, src, (size_t)n * sizeof(llb_point), cudaMemcpyHostToDevice, c->stream));
        unpack_points_kernel<<<std::min(div_up(n, 256), 148 * 8), 256, 0, c->stream>>>(c->raw32[slot].p, n, dst.p);
        LLB_CUDA(cudaGetLastError());
        c->launches++;
        return;
    }
    // the staging pair of this slot may still be the source of an earlier DMA
    if (c->pin_busy[slot]) { LLB_CUDA(cudaEventSynchronize(c->pin_ev[slot])); c->pin_busy[slot] = false; }
    c->pin_in[slot].ensure((size_t)n * 8);
    c->raw32[slot].ensure((size_t)n * 8);
    std::memcpy(c->pin_in[slot].p, src, (size_t)n * sizeof(llb_point));
    LLB_CUDA(cudaMemcpyAsync(c->raw32[slot].p, c->pin_in[slot].p, (size_t)n * sizeof(llb_point),
                             cudaMemcpyHostToDevice, c->stream));
    LLB_CUDA(cudaEventRecord(c->pin_ev[slot], c->stream));
    c->pin_busy[slot] = true;
    unpack_points_kernel<<<std::min(div_up(n, 256), 148 * 8), 256, 0, c->stream>>>(c->raw32[slot].p, n, dst.p);
    LLB_CUDA(cudaGetLastError());
    c->launches++;
}

void set_cloud(llb_ctx *c, int slot, Cloud &cl, const llb_point *src, int n)
{
    upload_cloud(c, slot, src, n, cl.pts);
    cl.view = cl.pts.p;
    cl.n_host = n; cl.exact = true;
}

// device float4 -> host llb_point (blocking)
void download_cloud(llb_ctx *c, const float4 *src, int n, llb_point *dst)
{
    if (n <= 0) return;
    c->pin_out.ensure(n);
    LLB_CUDA(cudaMemcpyAsync(c->pin_out.p, src, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    LLB_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; i++) {
        const float4 p = c->pin_out.p[i];
        dst[i].x = p.x; dst[i].y = p.y; dst[i].z = p.z; dst[i].w = 1.0f;
        dst[i].intensity = p.w; dst[i].c1 = dst[i].c2 = dst[i].c3 = 0.f;
    }
}

// downsampleCurrentScan of small sweeps runs on the forked stream beside whatever the caller enqueues next (normally
// the index build of the map); every consumer of the DS clouds / their counts and every producer that would overwrite
// the sweep joins here first
void join_scan_ds(llb_ctx *c)
{
    if (!c->ds_pending) return;
    LLB_CUDA(cudaStreamWaitEvent(c->stream, c->ds_ev, 0));
    c->ds_pending = false;
}

int read_count(llb_ctx *c, int which)
{
    join_scan_ds(c);
    LLB_CUDA(cudaMemcpyAsync(c->pin_counts.p, c->counts.p, sizeof(int) * llb_ctx::C_N, cudaMemcpyDeviceToHost, c->stream));
    LLB_CUDA(cudaStreamSynchronize(c->stream));
    return c->pin_counts.p[which];
}

void build_indices(llb_ctx *c)
{
    const float radius = std::sqrt(c->prm.knn_max_sqdist);
    const int *nc = c->map_counts_on_dev ? c->counts.p + llb_ctx::C_MAP_CORNER_DS : nullptr;
    const int *ns = c->map_counts_on_dev ? c->counts.p + llb_ctx::C_MAP_SURF_DS : nullptr;
    c->launches += GridIndex::build_pair(c->gridCorner, c->mapCornerDS_view, nc, c->mapCornerDS_upper,
                                         c->gridSurf, c->mapSurfDS_view, ns, c->mapSurfDS_upper, radius, c->stream);
    c->map_set = true;
    // every map setter ends here: the registration kernels learn whether this map is one rank's part of a sharded map
    c->s2m.set_shard(c->shard.axis, c->shard.lo, c->shard.hi, c->shard.axis >= 0 ? c->shard_global[0] : -1,
                     c->shard.axis >= 0 ? c->shard_global[1] : -1);
}

void voxel_map_raw(llb_ctx *c, const float4 *corner, int rc, const float4 *surf, int rs)
{
    c->shard = ShardPlan{};                                  // an unsharded map
    c->mapCornerDS.ensure(std::max(rc, 1)); c->mapSurfDS.ensure(std::max(rs, 1));
    VoxelInput a; a.a = corner; a.na = rc;
    VoxelInput b; b.a = surf; b.na = rs;
    // the two filters are independent (MO:1058-1060 / MO:1061-1063): the corner filter runs on a forked stream with its
    // own scratch while the (larger) surf filter runs on the context's stream; both are ~18 dependent small launches,
    // so running them side by side nearly halves the latency of the pair
    LLB_CUDA(cudaEventRecord(c->fork_ev, c->stream));
    LLB_CUDA(cudaStreamWaitEvent(c->stream2, c->fork_ev, 0));
    c->launches += c->vox2.run(a, c->prm.corner_leaf, c->mapCornerDS.p, c->counts.p + llb_ctx::C_MAP_CORNER_DS, c->stream2);
    LLB_CUDA(cudaEventRecord(c->join_ev, c->stream2));
    c->launches += c->vox.run(b, c->prm.surf_leaf, c->mapSurfDS.p, c->counts.p + llb_ctx::C_MAP_SURF_DS, c->stream);
    LLB_CUDA(cudaStreamWaitEvent(c->stream, c->join_ev, 0));
    c->mapCornerDS_view = c->mapCornerDS.p; c->mapSurfDS_view = c->mapSurfDS.p;
    c->mapCornerDS_upper = rc; c->mapSurfDS_upper = rs;
    c->map_counts_on_dev = true;
}

S2mQueries s2m_queries(llb_ctx *c)
{
    join_scan_ds(c);
    S2mQueries q;
    q.corner = c->cornerLastDS.p; q.nc_dev = c->counts.p + llb_ctx::C_CORNER_DS; q.nc_upper = c->cornerLast.n_host;
    q.surf = c->surfTotalLastDS.p; q.ns_dev = c->counts.p + llb_ctx::C_SURFTOTAL_DS;
    q.ns_upper = c->surfLast.n_host + c->outlierLast.n_host;
    return q;
}

void downsample_scan(llb_ctx *c)
{
    join_scan_ds(c);
    const int nc = c->cornerLast.n_host, ns = c->surfLast.n_host, no = c->outlierLast.n_host;
    c->cornerLastDS.ensure(std::max(nc, 1)); c->surfLastDS.ensure(std::max(ns, 1));
    c->outlierLastDS.ensure(std::max(no, 1)); c->surfTotalLastDS.ensure(std::max(ns + no, 1));
    VoxelInput in[3];
    in[0].a = c->cornerLast.view; in[0].na = nc;
    in[1].a = c->surfLast.view; in[1].na = ns;
    in[2].a = c->outlierLast.view; in[2].na = no;
    const float leaf[3] = { c->prm.corner_leaf, c->prm.surf_leaf, c->prm.outlier_leaf };
    float4 *out[3] = { c->cornerLastDS.p, c->surfLastDS.p, c->outlierLastDS.p };
    int *cnt[3] = { c->counts.p + llb_ctx::C_CORNER_DS, c->counts.p + llb_ctx::C_SURF_DS, c->counts.p + llb_ctx::C_OUTLIER_DS };
    // sweeps that fit the cluster kernel need no scratch: their four filters run on the forked stream and overlap the
    // map side of the registration (index build / map voxel filters) that the caller enqueues next
    const bool forked = std::max(nc, std::max(ns, no)) <= VoxelFilter::SMALL_MAX && ns + no <= VoxelFilter::SMALL_MAX;
    cudaStream_t vs = c->stream;
    if (forked) {
        LLB_CUDA(cudaEventRecord(c->ds_fork_ev, c->stream));
        LLB_CUDA(cudaStreamWaitEvent(c->stream2, c->ds_fork_ev, 0));
        vs = c->stream2;
    }
    c->launches += c->vox.run_batch(in, leaf, out, cnt, 3, vs);                        // MO:1069-1082
    VoxelInput tot;                                                                    // MO:1084-1090 (C12)
    tot.a = c->surfLastDS.p; tot.na_dev = cnt[1]; tot.na = ns;
    tot.b = c->outlierLastDS.p; tot.nb_dev = cnt[2]; tot.nb = no;
    c->launches += c->vox.run(tot, c->prm.surf_leaf, c->surfTotalLastDS.p, c->counts.p + llb_ctx::C_SURFTOTAL_DS, vs);
    if (forked) { LLB_CUDA(cudaEventRecord(c->ds_ev, c->stream2)); c->ds_pending = true; }
    c->scan_ds_done = true;
}

// a context that holds one rank's SLAB of a sharded map sees only its own queries: the single-rank entry points would
// silently return a partial solution, so they refuse; llb_s2m_optimize_sharded / llb_s2m_accumulate + llb_s2m_solve go on
bool holds_slab(llb_ctx *c)
{
    if (c->shard.axis >= 0 && c->shard.world > 1) {
        c->err = "this context holds one rank's slab of a sharded map: use llb_s2m_optimize_sharded (or llb_s2m_accumulate / "
                 "llb_s2m_solve with an all-reduce)";
        return true;
    }
    return false;
}

void fill_stats(llb_ctx *c, llb_stats *st, const S2mState &s, float ms)
{
    if (!st) return;
    st->iterations = s.iters; st->converged = s.converged; st->n_correspondences = s.n_corr;
    st->is_degenerate = s.is_degenerate; st->skipped = s.skipped; st->device_ms = ms;
    st->n_corner_ds = c->pin_counts.p[llb_ctx::C_CORNER_DS]; st->n_surf_ds = c->pin_counts.p[llb_ctx::C_SURFTOTAL_DS];
}

}  // namespace

extern "C" {

int llb_abi_version(void) { return LLB_ABI_VERSION; }

void llb_params_default(llb_params *p)
{
    if (!p) return;
    p->corner_leaf = 0.2f; p->surf_leaf = 0.4f; p->outlier_leaf = 0.4f;
    p->knn_max_sqdist = 1.0f;
    p->s2m_max_iterations = 10; p->s2m_min_correspondences = 50; p->s2m_degeneracy_thresh = 100.f;
    p->s2m_converge_deg = 0.05f; p->s2m_converge_cm = 0.05f;
    p->corner_map_min = 10; p->surf_map_min = 100;
    p->odom_nearest_sqdist = 25.f; p->odom_max_iterations = 25; p->odom_min_correspondences = 10;
    p->odom_degeneracy_thresh = 10.f; p->odom_converge_deg = 0.1f; p->odom_converge_cm = 0.1f;
    p->max_grid_cells = 1 << 23;
    p->pin_host_clouds = 0;
    p->s2m_max_ctas = 0;
}

int llb_create(const llb_params *p, int device, llb_ctx **out)
{
    if (!out) return LLB_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return LLB_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) return LLB_ERR_NO_DEVICE;
    llb_ctx *c = new llb_ctx();
    c->device = device;
    if (p) c->prm = *p; else llb_params_default(&c->prm);
    int rc = guarded(c, [&]() {
        LLB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        LLB_CUDA(cudaEventCreate(&c->ev0)); LLB_CUDA(cudaEventCreate(&c->ev1));
        for (int i = 0; i < 3; i++) LLB_CUDA(cudaEventCreateWithFlags(&c->pin_ev[i], cudaEventDisableTiming));
        LLB_CUDA(cudaEventCreateWithFlags(&c->asm_ev, cudaEventDisableTiming));
        c->counts.ensure(llb_ctx::C_N);
        LLB_CUDA(cudaMemset(c->counts.p, 0, sizeof(int) * llb_ctx::C_N));
        c->pin_counts.ensure(llb_ctx::C_N + 8);
        c->pin_state.ensure(1);
        c->pin_ostate.ensure(1);
        c->vox.init();
        c->vox2.init();
        c->vox3.init();
        LLB_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
        LLB_CUDA(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
        LLB_CUDA(cudaEventCreateWithFlags(&c->join_ev, cudaEventDisableTiming));
        LLB_CUDA(cudaEventCreateWithFlags(&c->ds_fork_ev, cudaEventDisableTiming));
        LLB_CUDA(cudaEventCreateWithFlags(&c->ds_ev, cudaEventDisableTiming));
        c->gridCorner.init(c->prm.max_grid_cells);
        c->gridSurf.init(c->prm.max_grid_cells);
        c->s2m.init(s2m_params(c->prm));
        c->odom.init(odom_params(c->prm));
        LLB_CUDA(cudaDeviceSynchronize());
        return (int)LLB_OK;
    });
    if (rc != LLB_OK) { delete c; return rc; }
    *out = c;
    return LLB_OK;
}

int llb_destroy(llb_ctx *c)
{
    if (!c) return LLB_ERR_INVALID;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 3; i++) { c->pin_in[i].release(); c->raw32[i].release(); if (c->pin_ev[i]) cudaEventDestroy(c->pin_ev[i]); }
    c->tmp_in.release();
    for (auto &r : c->regs) if (r.ours) cudaHostUnregister(const_cast<void *>(r.p));
    c->regs.clear();
    c->pin_out.release(); c->pin_counts.release(); c->pin_state.release(); c->pin_ostate.release();
    c->counts.release(); c->vox.release(); c->vox2.release();
    if (c->fork_ev) cudaEventDestroy(c->fork_ev);
    if (c->join_ev) cudaEventDestroy(c->join_ev);
    if (c->ds_fork_ev) cudaEventDestroy(c->ds_fork_ev);
    if (c->ds_ev) cudaEventDestroy(c->ds_ev);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    c->cornerLast.pts.release(); c->surfLast.pts.release(); c->outlierLast.pts.release();
    c->cornerLastDS.release(); c->surfLastDS.release(); c->outlierLastDS.release(); c->surfTotalLastDS.release();
    c->mapCornerRaw.pts.release(); c->mapSurfRaw.pts.release(); c->mapCornerDS.release(); c->mapSurfDS.release();
    c->gridCorner.release(); c->gridSurf.release(); c->s2m.release(); c->odom.release();
    c->dbg_coeff.release(); c->dbg_valid.release(); c->dbg_knn.release(); c->dbg_d2.release(); c->tmp_vox.release();
    for (int r = 0; r < S2M_MAX_PEERS; r++) if (c->p2p_opened[r]) cudaIpcCloseMemHandle(c->p2p_opened[r]);
    if (c->p2p_mem) cudaFree(c->p2p_mem);
    c->features.release();
    c->projection.release();
    c->icp.release(); c->vox3.release(); c->pin_icp.release();
    for (DevBuf<float4> *b : { &c->loopLatestRaw, &c->loopLatest, &c->loopHistRaw, &c->loopHistDS, &c->globalRaw, &c->globalDS, &c->shardCorner, &c->shardSurf }) b->release();
    c->shard_blk.release(); c->shard_cnt.release(); c->shard_samp.release(); c->shard_samp_pin.release();
    c->kfs.release(); c->asmCorner.release(); c->asmSurf.release(); c->asm_segs.release(); c->pin_segs.release();
    if (c->asm_ev) cudaEventDestroy(c->asm_ev);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return LLB_OK;
}

const char *llb_last_error(const llb_ctx *c) { return c ? c->err.c_str() : "null context"; }
void *llb_stream(llb_ctx *c) { return c ? (void *)c->stream : nullptr; }
long long llb_launch_count(const llb_ctx *c) { return c ? c->launches : 0; }

int llb_synchronize(llb_ctx *c)
{
    return guarded(c, [&]() { join_scan_ds(c); LLB_CUDA(cudaStreamSynchronize(c->stream)); return (int)LLB_OK; });
}

int llb_reserve(llb_ctx *c, int max_scan_points, int max_raw_map_points, int max_keyframes)
{
    return guarded(c, [&]() {
        if (max_scan_points < 0 || max_raw_map_points < 0 || max_keyframes < 0) return (int)LLB_ERR_INVALID;
        const size_t ns = (size_t)std::max(max_scan_points, 1), nm = (size_t)std::max(max_raw_map_points, 1);
        // scan side: the three sweeps clouds, their staging and the four DS clouds
        for (int i = 0; i < 3; i++) { c->pin_in[i].ensure(std::max(ns, nm) * 8); c->raw32[i].ensure(std::max(ns, nm) * 8); }
        c->cornerLast.pts.ensure(ns); c->surfLast.pts.ensure(ns); c->outlierLast.pts.ensure(ns);
        c->cornerLastDS.ensure(ns); c->surfLastDS.ensure(ns); c->outlierLastDS.ensure(ns); c->surfTotalLastDS.ensure(2 * ns);
        // map side: raw map (host hand-over and key-frame assembly), DS map, voxel scratch, both spatial indices
        c->mapCornerRaw.pts.ensure(nm); c->mapSurfRaw.pts.ensure(nm);
        c->asmCorner.ensure(nm); c->asmSurf.ensure(nm);
        c->mapCornerDS.ensure(nm); c->mapSurfDS.ensure(nm);
        c->vox.reserve((int)std::max(nm, 2 * ns)); c->vox2.reserve((int)nm);
        c->gridCorner.job(nullptr, nullptr, (int)nm); c->gridSurf.job(nullptr, nullptr, (int)nm);
        c->tmp_in.ensure(std::max(ns, nm)); c->tmp_vox.ensure(std::max(ns, nm)); c->pin_out.ensure(std::max(ns, nm));
        const int nq = 3 * (int)ns;
        c->dbg_coeff.ensure(nq); c->dbg_valid.ensure(nq); c->dbg_knn.ensure((size_t)nq * 5); c->dbg_d2.ensure((size_t)nq * 5);
        c->asm_segs.ensure(3 * (size_t)std::max(max_keyframes, 1)); c->pin_segs.ensure(3 * (size_t)std::max(max_keyframes, 1));
        if (max_keyframes > 0) c->kfs.reserve((size_t)max_keyframes * 3 * ns / 2);   // DS clouds are well below the raw sweep size
        LLB_CUDA(cudaDeviceSynchronize());
        return (int)LLB_OK;
    });
}

int llb_voxel_downsample(llb_ctx *c, const llb_point *in, int n, float leaf, llb_point *out, int cap, int *m)
{
    return guarded(c, [&]() {
        if (n < 0 || (n > 0 && !in) || !m || !(leaf > 0.f)) return (int)LLB_ERR_INVALID;
        if (n == 0) { *m = 0; return (int)LLB_OK; }
        upload_cloud(c, 0, in, n, c->tmp_in);
        c->tmp_vox.ensure(n);
        VoxelInput vi; vi.a = c->tmp_in.p; vi.na = n;
        c->launches += c->vox.run(vi, leaf, c->tmp_vox.p, c->counts.p + llb_ctx::C_VOX_TMP, c->stream);
        int cnt = read_count(c, llb_ctx::C_VOX_TMP);
        *m = cnt;
        if (cnt > cap) return (int)LLB_ERR_CAPACITY;
        if (out) download_cloud(c, c->tmp_vox.p, cnt, out);
        return (int)LLB_OK;
    });
}

int llb_map_set_ds(llb_ctx *c, const llb_point *corner, int mc, const llb_point *surf, int ms)
{
    return guarded(c, [&]() {
        if (mc < 0 || ms < 0 || (mc > 0 && !corner) || (ms > 0 && !surf)) return (int)LLB_ERR_INVALID;
        upload_cloud(c, 0, corner, mc, c->mapCornerDS);
        upload_cloud(c, 1, surf, ms, c->mapSurfDS);
        c->mapCornerDS_view = c->mapCornerDS.p; c->mapSurfDS_view = c->mapSurfDS.p;
        c->mapCornerDS_upper = mc; c->mapSurfDS_upper = ms;
        c->map_counts_on_dev = false;
        c->shard = ShardPlan{};
        build_indices(c);
        return (int)LLB_OK;
    });
}

int llb_map_set_ds_dev(llb_ctx *c, const void *corner, int mc, const void *surf, int ms)
{
    return guarded(c, [&]() {
        if (mc < 0 || ms < 0 || (mc > 0 && !corner) || (ms > 0 && !surf)) return (int)LLB_ERR_INVALID;
        c->mapCornerDS_view = (const float4 *)corner; c->mapSurfDS_view = (const float4 *)surf;
        c->mapCornerDS_upper = mc; c->mapSurfDS_upper = ms;
        c->map_counts_on_dev = false;
        c->shard = ShardPlan{};
        build_indices(c);
        return (int)LLB_OK;
    });
}

int llb_map_set_raw(llb_ctx *c, const llb_point *corner, int rc, const llb_point *surf, int rs)
{
    return guarded(c, [&]() {
        if (rc < 0 || rs < 0 || (rc > 0 && !corner) || (rs > 0 && !surf)) return (int)LLB_ERR_INVALID;
        set_cloud(c, 0, c->mapCornerRaw, corner, rc);
        set_cloud(c, 1, c->mapSurfRaw, surf, rs);
        voxel_map_raw(c, c->mapCornerRaw.pts.p, rc, c->mapSurfRaw.pts.p, rs);
        build_indices(c);
        return (int)LLB_OK;
    });
}

int llb_map_set_raw_dev(llb_ctx *c, const void *corner, int rc, const void *surf, int rs)
{
    return guarded(c, [&]() {
        if (rc < 0 || rs < 0 || (rc > 0 && !corner) || (rs > 0 && !surf)) return (int)LLB_ERR_INVALID;
        voxel_map_raw(c, (const float4 *)corner, rc, (const float4 *)surf, rs);
        build_indices(c);
        return (int)LLB_OK;
    });
}

int llb_map_get_ds(llb_ctx *c, int which, llb_point *out, int cap, int *n)
{
    return guarded(c, [&]() {
        if (!n || which < 0 || which > 1) return (int)LLB_ERR_INVALID;
        if (!c->map_set) return (int)LLB_ERR_STATE;
        int cnt = which == 0 ? c->mapCornerDS_upper : c->mapSurfDS_upper;
        if (c->map_counts_on_dev) cnt = read_count(c, which == 0 ? llb_ctx::C_MAP_CORNER_DS : llb_ctx::C_MAP_SURF_DS);
        *n = cnt;
        if (!out) return (int)LLB_OK;
        if (cnt > cap) return (int)LLB_ERR_CAPACITY;
        download_cloud(c, which == 0 ? c->mapCornerDS_view : c->mapSurfDS_view, cnt, out);
        return (int)LLB_OK;
    });
}

int llb_scan_set(llb_ctx *c, const llb_point *corner, int nc, const llb_point *surf, int ns,
                 const llb_point *outlier, int no)
{
    return guarded(c, [&]() {
        if (nc < 0 || ns < 0 || no < 0 || (nc > 0 && !corner) || (ns > 0 && !surf) || (no > 0 && !outlier))
            return (int)LLB_ERR_INVALID;
        join_scan_ds(c);
        set_cloud(c, 0, c->cornerLast, corner, nc);
        set_cloud(c, 1, c->surfLast, surf, ns);
        set_cloud(c, 2, c->outlierLast, outlier, no);
        c->scan_set = true; c->scan_ds_done = false;
        return (int)LLB_OK;
    });
}

int llb_scan_set_dev(llb_ctx *c, const void *corner, int nc, const void *surf, int ns, const void *outlier, int no)
{
    return guarded(c, [&]() {
        if (nc < 0 || ns < 0 || no < 0) return (int)LLB_ERR_INVALID;
        join_scan_ds(c);
        // device-resident sweeps are BORROWED (no copy): the pointers must stay valid and unchanged until the next
        // llb_scan_set* / the end of the registration that uses them
        auto cp = [&](Cloud &cl, const void *src, int n) {
            cl.view = (const float4 *)src;
            cl.n_host = n; cl.exact = true;
        };
        cp(c->cornerLast, corner, nc); cp(c->surfLast, surf, ns); cp(c->outlierLast, outlier, no);
        c->scan_set = true; c->scan_ds_done = false;
        return (int)LLB_OK;
    });
}

int llb_downsample_current_scan(llb_ctx *c, int counts[4])
{
    return guarded(c, [&]() {
        if (!c->scan_set) return (int)LLB_ERR_STATE;
        downsample_scan(c);
        if (counts) {
            read_count(c, 0);
            for (int i = 0; i < 4; i++) counts[i] = c->pin_counts.p[i];
        }
        return (int)LLB_OK;
    });
}

int llb_scan_get_ds(llb_ctx *c, int which, llb_point *out, int cap, int *n)
{
    return guarded(c, [&]() {
        if (!n || which < 0 || which > 3) return (int)LLB_ERR_INVALID;
        if (!c->scan_ds_done) return (int)LLB_ERR_STATE;
        int cnt = read_count(c, which);
        *n = cnt;
        if (!out) return (int)LLB_OK;
        if (cnt > cap) return (int)LLB_ERR_CAPACITY;
        const float4 *src = which == 0 ? c->cornerLastDS.p : which == 1 ? c->surfLastDS.p
                          : which == 2 ? c->outlierLastDS.p : c->surfTotalLastDS.p;
        download_cloud(c, src, cnt, out);
        return (int)LLB_OK;
    });
}

int llb_s2m_iterate(llb_ctx *c, float T[6], int iter, int *converged, int *n_corr)
{
    return guarded(c, [&]() {
        if (!T || iter < 0) return (int)LLB_ERR_INVALID;
        if (!c->map_set || !c->scan_ds_done) return (int)LLB_ERR_STATE;
        if (holds_slab(c)) return (int)LLB_ERR_STATE;
        S2mQueries q = s2m_queries(c);
        const int nq = q.nc_upper + q.ns_upper;
        c->dbg_coeff.ensure(std::max(nq, 1)); c->dbg_valid.ensure(std::max(nq, 1));
        c->dbg_knn.ensure((size_t)std::max(nq, 1) * 5); c->dbg_d2.ensure((size_t)std::max(nq, 1) * 5);
        S2mDebug dbg{ c->dbg_coeff.p, c->dbg_valid.p, c->dbg_knn.p, c->dbg_d2.p };
        c->launches += c->s2m.prepare(T, nullptr, c->gridCorner.desc_dev(), c->gridSurf.desc_dev(), c->stream);
        c->launches += c->s2m.run(iter, iter + 1, q, c->gridCorner.view(), c->gridSurf.view(), dbg, 0, 1, true, c->stream);
        LLB_CUDA(cudaMemcpyAsync(c->pin_state.p, c->s2m.state_dev(), sizeof(S2mState), cudaMemcpyDeviceToHost, c->stream));
        read_count(c, 0);                                           // also synchronises
        const S2mState &s = *c->pin_state.p;
        c->dbg_nc = c->pin_counts.p[llb_ctx::C_CORNER_DS]; c->dbg_ns = c->pin_counts.p[llb_ctx::C_SURFTOTAL_DS];
        c->dbg_ready = !s.skipped;
        for (int i = 0; i < 6; i++) T[i] = s.T[i];
        if (converged) *converged = s.converged;
        if (n_corr) *n_corr = s.n_corr;
        return (int)LLB_OK;
    });
}

int llb_s2m_optimize(llb_ctx *c, float T[6], llb_stats *stats)
{
    return guarded(c, [&]() {
        if (!T) return (int)LLB_ERR_INVALID;
        if (!c->map_set || !c->scan_ds_done) return (int)LLB_ERR_STATE;
        if (holds_slab(c)) return (int)LLB_ERR_STATE;
        S2mQueries q = s2m_queries(c);
        S2mDebug dbg{ nullptr, nullptr, nullptr, nullptr };
        LLB_CUDA(cudaEventRecord(c->ev0, c->stream));
        c->launches += c->s2m.prepare(T, nullptr, c->gridCorner.desc_dev(), c->gridSurf.desc_dev(), c->stream);
        c->launches += c->s2m.run(0, c->prm.s2m_max_iterations, q, c->gridCorner.view(), c->gridSurf.view(), dbg, 0, 1,
                                  true, c->stream);
        LLB_CUDA(cudaEventRecord(c->ev1, c->stream));
        LLB_CUDA(cudaMemcpyAsync(c->pin_state.p, c->s2m.state_dev(), sizeof(S2mState), cudaMemcpyDeviceToHost, c->stream));
        read_count(c, 0);
        float ms = 0.f;
        LLB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        const S2mState &s = *c->pin_state.p;
        if (!s.skipped) for (int i = 0; i < 6; i++) T[i] = s.T[i];
        fill_stats(c, stats, s, ms);
        c->dbg_ready = false;
        return (int)LLB_OK;
    });
}

int llb_s2m_optimize_async(llb_ctx *c, const float T[6])
{
    return guarded(c, [&]() {
        if (!T) return (int)LLB_ERR_INVALID;
        if (!c->map_set || !c->scan_ds_done) return (int)LLB_ERR_STATE;
        if (holds_slab(c)) return (int)LLB_ERR_STATE;
        S2mQueries q = s2m_queries(c);
        S2mDebug dbg{ nullptr, nullptr, nullptr, nullptr };
        LLB_CUDA(cudaEventRecord(c->ev0, c->stream));
        c->launches += c->s2m.prepare(T, nullptr, c->gridCorner.desc_dev(), c->gridSurf.desc_dev(), c->stream);
        c->launches += c->s2m.run(0, c->prm.s2m_max_iterations, q, c->gridCorner.view(), c->gridSurf.view(), dbg, 0, 1,
                                  true, c->stream);
        LLB_CUDA(cudaEventRecord(c->ev1, c->stream));
        LLB_CUDA(cudaMemcpyAsync(c->pin_state.p, c->s2m.state_dev(), sizeof(S2mState), cudaMemcpyDeviceToHost, c->stream));
        LLB_CUDA(cudaMemcpyAsync(c->pin_counts.p, c->counts.p, sizeof(int) * llb_ctx::C_N, cudaMemcpyDeviceToHost, c->stream));
        c->async_pending = true;
        c->dbg_ready = false;
        return (int)LLB_OK;
    });
}

int llb_s2m_result(llb_ctx *c, float T[6], llb_stats *stats)
{
    return guarded(c, [&]() {
        if (!T) return (int)LLB_ERR_INVALID;
        if (!c->async_pending) return (int)LLB_ERR_STATE;
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        c->async_pending = false;
        float ms = 0.f;
        LLB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        const S2mState &s = *c->pin_state.p;
        if (!s.skipped) for (int i = 0; i < 6; i++) T[i] = s.T[i];
        fill_stats(c, stats, s, ms);
        return (int)LLB_OK;
    });
}

int llb_s2m_optimize_dev(llb_ctx *c, float *T_dev)
{
    return guarded(c, [&]() {
        if (!T_dev) return (int)LLB_ERR_INVALID;
        if (!c->map_set || !c->scan_ds_done) return (int)LLB_ERR_STATE;
        if (holds_slab(c)) return (int)LLB_ERR_STATE;
        S2mQueries q = s2m_queries(c);
        S2mDebug dbg{ nullptr, nullptr, nullptr, nullptr };
        c->launches += c->s2m.prepare(nullptr, T_dev, c->gridCorner.desc_dev(), c->gridSurf.desc_dev(), c->stream);
        c->launches += c->s2m.run(0, c->prm.s2m_max_iterations, q, c->gridCorner.view(), c->gridSurf.view(), dbg, 0, 1,
                                  true, c->stream);
        // S2mState starts with T[6]
        LLB_CUDA(cudaMemcpyAsync(T_dev, c->s2m.state_dev(), 6 * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        c->dbg_ready = false;
        return (int)LLB_OK;
    });
}

int llb_s2m_pose_set(llb_ctx *c, const float T[6])
{
    return guarded(c, [&]() {
        if (!T) return (int)LLB_ERR_INVALID;
        if (!c->map_set) return (int)LLB_ERR_STATE;
        c->launches += c->s2m.prepare(T, nullptr, c->gridCorner.desc_dev(), c->gridSurf.desc_dev(), c->stream);
        return (int)LLB_OK;
    });
}

int llb_s2m_pose_get(llb_ctx *c, float T[6])
{
    return guarded(c, [&]() {
        if (!T) return (int)LLB_ERR_INVALID;
        LLB_CUDA(cudaMemcpyAsync(c->pin_state.p, c->s2m.state_dev(), sizeof(S2mState), cudaMemcpyDeviceToHost, c->stream));
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < 6; i++) T[i] = c->pin_state.p->T[i];
        return (int)LLB_OK;
    });
}

int llb_s2m_accumulate(llb_ctx *c, int iter, int rank, int world, double **acc)
{
    return guarded(c, [&]() {
        if (!acc || world < 1 || rank < 0 || rank >= world) return (int)LLB_ERR_INVALID;
        if (!c->map_set || !c->scan_ds_done) return (int)LLB_ERR_STATE;
        S2mQueries q = s2m_queries(c);
        S2mDebug dbg{ nullptr, nullptr, nullptr, nullptr };
        c->launches += c->s2m.run(iter, iter + 1, q, c->gridCorner.view(), c->gridSurf.view(), dbg, rank, world, false, c->stream);
        *acc = c->s2m.acc_dev();
        return (int)LLB_OK;
    });
}

int llb_s2m_time_iteration(llb_ctx *c, const float T[6], int reps, float *ms_per_launch, int *n_queries)
{
    return guarded(c, [&]() {
        if (!T || reps < 1 || !ms_per_launch) return (int)LLB_ERR_INVALID;
        if (!c->map_set || !c->scan_ds_done) return (int)LLB_ERR_STATE;
        S2mQueries q = s2m_queries(c);
        S2mDebug dbg{ nullptr, nullptr, nullptr, nullptr };
        c->launches += c->s2m.prepare(T, nullptr, c->gridCorner.desc_dev(), c->gridSurf.desc_dev(), c->stream);
        c->launches += c->s2m.run(1, 2, q, c->gridCorner.view(), c->gridSurf.view(), dbg, 0, 1, false, c->stream);
        LLB_CUDA(cudaEventRecord(c->ev0, c->stream));
        for (int r = 0; r < reps; r++)
            c->launches += c->s2m.run(1, 2, q, c->gridCorner.view(), c->gridSurf.view(), dbg, 0, 1, false, c->stream);
        LLB_CUDA(cudaEventRecord(c->ev1, c->stream));
        read_count(c, 0);
        float ms = 0.f;
        LLB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        *ms_per_launch = ms / reps;
        if (n_queries) *n_queries = c->pin_counts.p[llb_ctx::C_CORNER_DS] + c->pin_counts.p[llb_ctx::C_SURFTOTAL_DS];
        return (int)LLB_OK;
    });
}

int llb_s2m_solve(llb_ctx *c, int iter, int *converged)
{
    return guarded(c, [&]() {
        c->launches += c->s2m.solve(iter, c->stream);
        if (converged) {
            LLB_CUDA(cudaMemcpyAsync(c->pin_state.p, c->s2m.state_dev(), sizeof(S2mState), cudaMemcpyDeviceToHost, c->stream));
            LLB_CUDA(cudaStreamSynchronize(c->stream));
            *converged = c->pin_state.p->converged;
        }
        return (int)LLB_OK;
    });
}

int llb_get_correspondences(llb_ctx *c, llb_point *ori, llb_point *coeff, int cap, int *n)
{
    return guarded(c, [&]() {
        if (!n) return (int)LLB_ERR_INVALID;
        if (!c->dbg_ready) { *n = 0; return (int)LLB_ERR_STATE; }
        const int nq = c->dbg_nc + c->dbg_ns;
        std::vector<float4> co(std::max(nq, 1)), qc(std::max(c->dbg_nc, 1)), qs(std::max(c->dbg_ns, 1));
        std::vector<int> va(std::max(nq, 1));
        // rows live at [0,nc) for corner queries and [nc, nc+ns) for surf queries
        if (nq > 0) {
            LLB_CUDA(cudaMemcpyAsync(co.data(), c->dbg_coeff.p, sizeof(float4) * nq, cudaMemcpyDeviceToHost, c->stream));
            LLB_CUDA(cudaMemcpyAsync(va.data(), c->dbg_valid.p, sizeof(int) * nq, cudaMemcpyDeviceToHost, c->stream));
        }
        if (c->dbg_nc > 0) LLB_CUDA(cudaMemcpyAsync(qc.data(), c->cornerLastDS.p, sizeof(float4) * c->dbg_nc, cudaMemcpyDeviceToHost, c->stream));
        if (c->dbg_ns > 0) LLB_CUDA(cudaMemcpyAsync(qs.data(), c->surfTotalLastDS.p, sizeof(float4) * c->dbg_ns, cudaMemcpyDeviceToHost, c->stream));
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        int cnt = 0;
        for (int i = 0; i < nq; i++) cnt += va[i] != 0;
        *n = cnt;
        if (!ori || !coeff) return (int)LLB_OK;
        if (cnt > cap) return (int)LLB_ERR_CAPACITY;
        int k = 0;
        for (int i = 0; i < nq; i++) {
            if (!va[i]) continue;
            const float4 p = i < c->dbg_nc ? qc[i] : qs[i - c->dbg_nc];
            ori[k] = llb_point{ p.x, p.y, p.z, 1.0f, p.w, 0.f, 0.f, 0.f };
            coeff[k] = llb_point{ co[i].x, co[i].y, co[i].z, 1.0f, co[i].w, 0.f, 0.f, 0.f };
            k++;
        }
        return (int)LLB_OK;
    });
}

int llb_get_knn(llb_ctx *c, int which, int *idx5, float *d2, int cap, int *n)
{
    return guarded(c, [&]() {
        if (!n || which < 0 || which > 1) return (int)LLB_ERR_INVALID;
        if (!c->dbg_ready) { *n = 0; return (int)LLB_ERR_STATE; }
        const int cnt = which == 0 ? c->dbg_nc : c->dbg_ns;
        const int off = which == 0 ? 0 : c->dbg_nc;
        *n = cnt;
        if (!idx5 || !d2) return (int)LLB_OK;
        if (cnt > cap) return (int)LLB_ERR_CAPACITY;
        if (cnt > 0) {
            LLB_CUDA(cudaMemcpyAsync(idx5, c->dbg_knn.p + (size_t)off * 5, sizeof(int) * 5 * cnt, cudaMemcpyDeviceToHost, c->stream));
            LLB_CUDA(cudaMemcpyAsync(d2, c->dbg_d2.p + (size_t)off * 5, sizeof(float) * 5 * cnt, cudaMemcpyDeviceToHost, c->stream));
            LLB_CUDA(cudaStreamSynchronize(c->stream));
        }
        return (int)LLB_OK;
    });
}

int llb_get_normal_equations(llb_ctx *c, float AtA[36], float AtB[6], float X[6])
{
    return guarded(c, [&]() {
        LLB_CUDA(cudaMemcpyAsync(c->pin_state.p, c->s2m.state_dev(), sizeof(S2mState), cudaMemcpyDeviceToHost, c->stream));
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        if (AtA) std::memcpy(AtA, c->pin_state.p->AtA, sizeof(float) * 36);
        if (AtB) std::memcpy(AtB, c->pin_state.p->AtB, sizeof(float) * 6);
        if (X) std::memcpy(X, c->pin_state.p->X, sizeof(float) * 6);
        return (int)LLB_OK;
    });
}

int llb_s2m_get_cta_profile(llb_ctx *c, double *out, int cap, int *n_ctas)
{
    return guarded(c, [&]() {
        if (!n_ctas) return (int)LLB_ERR_INVALID;
        const int g = c->s2m.last_grid();
        *n_ctas = g;
        if (!out) return (int)LLB_OK;
        if (g > cap) return (int)LLB_ERR_CAPACITY;
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        if (g > 0) LLB_CUDA(cudaMemcpy(out, c->s2m.cta_profile_dev(), sizeof(double) * 4 * g, cudaMemcpyDeviceToHost));
        return (int)LLB_OK;
    });
}

int llb_s2m_get_profile(llb_ctx *c, int iter, long long stamps[8])
{
    return guarded(c, [&]() {
        if (!stamps || iter < 0) return (int)LLB_ERR_INVALID;
        LLB_CUDA(cudaMemcpyAsync(c->pin_state.p, c->s2m.state_dev(), sizeof(S2mState), cudaMemcpyDeviceToHost, c->stream));
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < 8; i++) stamps[i] = c->pin_state.p->prof[iter % 10][i];
        return (int)LLB_OK;
    });
}

int llb_get_degeneracy(llb_ctx *c, int *deg, float matP[36])
{
    return guarded(c, [&]() {
        if (matP) c->launches += c->s2m.ensure_matp(c->stream);
        LLB_CUDA(cudaMemcpyAsync(c->pin_state.p, c->s2m.state_dev(), sizeof(S2mState), cudaMemcpyDeviceToHost, c->stream));
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        if (deg) *deg = c->pin_state.p->is_degenerate;
        if (matP) std::memcpy(matP, c->pin_state.p->matP, sizeof(float) * 36);
        return (int)LLB_OK;
    });
}

int llb_set_degeneracy(llb_ctx *c, int deg, const float matP[36])
{
    return guarded(c, [&]() {
        if (!matP) return (int)LLB_ERR_INVALID;
        S2mState *d = c->s2m.state_dev();
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        LLB_CUDA(cudaMemcpy(&d->is_degenerate, &deg, sizeof(int), cudaMemcpyHostToDevice));
        LLB_CUDA(cudaMemcpy(d->matP, matP, sizeof(float) * 36, cudaMemcpyHostToDevice));
        const int one = 1;
        LLB_CUDA(cudaMemcpy(&d->matP_valid, &one, sizeof(int), cudaMemcpyHostToDevice));
        return (int)LLB_OK;
    });
}

// ------------------------------------------------------------------ fused multi-GPU exchange (BASELINE config 4)

namespace {
constexpr size_t P2P_DATA_BYTES = sizeof(double) * 2 * S2M_MAX_PEERS * 32;
constexpr size_t P2P_BYTES = P2P_DATA_BYTES + sizeof(unsigned long long) * 2 * S2M_MAX_PEERS;
}

int llb_p2p_export(llb_ctx *c, unsigned char handle[64])
{
    return guarded(c, [&]() {
        if (!handle) return (int)LLB_ERR_INVALID;
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        if (!c->p2p_mem) {
            LLB_CUDA(cudaMalloc(&c->p2p_mem, P2P_BYTES));
            LLB_CUDA(cudaMemset(c->p2p_mem, 0, P2P_BYTES));
            LLB_CUDA(cudaDeviceSynchronize());
        }
        cudaIpcMemHandle_t h;
        LLB_CUDA(cudaIpcGetMemHandle(&h, c->p2p_mem));
        std::memcpy(handle, &h, 64);
        return (int)LLB_OK;
    });
}

int llb_p2p_import(llb_ctx *c, int rank, int world, const unsigned char *handles)
{
    return guarded(c, [&]() {
        if (!handles || world < 1 || world > S2M_MAX_PEERS || rank < 0 || rank >= world) return (int)LLB_ERR_INVALID;
        if (!c->p2p_mem) return (int)LLB_ERR_STATE;
        for (int r = 0; r < world; r++) {
            unsigned char *base = c->p2p_mem;
            if (r != rank) {
                if (c->p2p_opened[r]) { cudaIpcCloseMemHandle(c->p2p_opened[r]); c->p2p_opened[r] = nullptr; }
                cudaIpcMemHandle_t h;
                std::memcpy(&h, handles + 64 * r, 64);
                void *p = nullptr;
                LLB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
                c->p2p_opened[r] = p;
                base = (unsigned char *)p;
            }
            c->peers.box[r] = (double *)base;
            c->peers.flag[r] = (unsigned long long *)(base + P2P_DATA_BYTES);
        }
        c->peers.rank = rank; c->peers.world = world;
        c->p2p_ready = true;
        return (int)LLB_OK;
    });
}

int llb_s2m_optimize_sharded(llb_ctx *c, float T[6], llb_stats *stats)
{
    return guarded(c, [&]() {
        if (!T) return (int)LLB_ERR_INVALID;
        if (!c->map_set || !c->scan_ds_done || !c->p2p_ready) return (int)LLB_ERR_STATE;
        S2mQueries q = s2m_queries(c);
        S2mDebug dbg{ nullptr, nullptr, nullptr, nullptr };
        LLB_CUDA(cudaEventRecord(c->ev0, c->stream));
        c->launches += c->s2m.prepare(T, nullptr, c->gridCorner.desc_dev(), c->gridSurf.desc_dev(), c->stream);
        // replicated map: the queries are dealt round-robin (qi = rank + world * j); sharded map: a rank walks all queries
        // and takes those inside its slab
        const bool slab = c->shard.axis >= 0 && c->shard.world > 1;
        if (slab && (c->shard.world != c->peers.world || c->shard.rank != c->peers.rank)) return (int)LLB_ERR_STATE;
        c->launches += c->s2m.run(0, c->prm.s2m_max_iterations, q, c->gridCorner.view(), c->gridSurf.view(), dbg,
                                  slab ? 0 : c->peers.rank, slab ? 1 : c->peers.world, true, c->stream, &c->peers);
        LLB_CUDA(cudaEventRecord(c->ev1, c->stream));
        LLB_CUDA(cudaMemcpyAsync(c->pin_state.p, c->s2m.state_dev(), sizeof(S2mState), cudaMemcpyDeviceToHost, c->stream));
        read_count(c, 0);
        float ms = 0.f;
        LLB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        const S2mState &s = *c->pin_state.p;
        if (s.peer_timeout) { c->err = "a peer rank did not deliver its normal equations (llb_s2m_optimize_sharded must be "
                                       "called by every rank with the same inputs)"; return (int)LLB_ERR_STATE; }
        if (!s.skipped) for (int i = 0; i < 6; i++) T[i] = s.T[i];
        fill_stats(c, stats, s, ms);
        c->dbg_ready = false;
        return (int)LLB_OK;
    });
}

// ------------------------------------------------------------------ sharded local map (BASELINE config 4, shard.cuh)

// pure host function: the slab of `rank` from a sample of map points (nsamp x {x, y, z}): axis = the longest extent of the
// sample, borders = its quantiles; rank 0 / world - 1 are open towards -inf / +inf
int llb_shard_plan(const float *sample_xyz, int nsamp, int rank, int world, int *axis_out, float *lo_out, float *hi_out)
{
    if (!sample_xyz || nsamp < 1 || world < 1 || rank < 0 || rank >= world || !axis_out || !lo_out || !hi_out) return (int)LLB_ERR_INVALID;
    float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int k = 0; k < nsamp; k++)
        for (int a = 0; a < 3; a++) { mn[a] = std::min(mn[a], sample_xyz[3 * k + a]); mx[a] = std::max(mx[a], sample_xyz[3 * k + a]); }
    int axis = 0;
    for (int a = 1; a < 3; a++) if (mx[a] - mn[a] > mx[axis] - mn[axis]) axis = a;
    std::vector<float> v((size_t)nsamp);
    for (int k = 0; k < nsamp; k++) v[k] = sample_xyz[3 * k + axis];
    // the two order statistics that border this rank's slab (the values a full sort would put there)
    auto order_stat = [&](size_t k) { std::nth_element(v.begin(), v.begin() + k, v.end()); return v[k]; };
    *axis_out = axis;
    *lo_out = rank == 0 ? -FLT_MAX : order_stat((size_t)nsamp * rank / world);
    *hi_out = rank == world - 1 ? FLT_MAX : order_stat((size_t)nsamp * (rank + 1) / world);
    return (int)LLB_OK;
}

namespace {

// the slab of this rank, from a deterministic sample of the raw surf map (every rank holds the same raw map and draws
// the same sample): axis = the longer horizontal extent of the sample, borders = its quantiles
void plan_shard(llb_ctx *c, const float4 *corner, int rc, const float4 *surf, int rs, int rank, int world)
{
    ShardPlan pl; pl.rank = rank; pl.world = world;
    const float4 *src = rs > 0 ? surf : corner; const int n = rs > 0 ? rs : rc;
    if (world <= 1 || n <= 0) { pl.axis = 0; c->shard = pl; return; }
    const int nsamp = std::min(n, 2048), stride = std::max(1, n / nsamp);
    c->shard_samp.ensure(3 * nsamp); c->shard_samp_pin.ensure(3 * nsamp);
    c->launches += launch_shard_sample(src, n, stride, nsamp, c->shard_samp.p, c->stream);
    LLB_CUDA(cudaMemcpyAsync(c->shard_samp_pin.p, c->shard_samp.p, sizeof(float) * 3 * nsamp, cudaMemcpyDeviceToHost, c->stream));
    LLB_CUDA(cudaStreamSynchronize(c->stream));
    int axis = 0; float lo = 0.f, hi = 0.f;
    llb_shard_plan(c->shard_samp_pin.p, nsamp, rank, world, &axis, &lo, &hi);
    pl.axis = axis; pl.lo = lo; pl.hi = hi;
    c->shard = pl;
}

// lattice range [ilo, ihi] on the plan's axis of the voxels (edge `leaf`) this rank must filter: every voxel that can
// hold a centroid within `reach` of a point of the slab, plus one voxel of margin on both sides
void slab_voxels(const ShardPlan &pl, float leaf, float reach, int &ilo, int &ihi)
{
    const float inv = 1.0f / leaf;
    ilo = pl.lo <= -FLT_MAX ? INT_MIN / 2 : (int)std::floor((pl.lo - reach) * inv) - 1;
    ihi = pl.hi >= FLT_MAX ? INT_MAX / 2 : (int)std::floor((pl.hi + reach) * inv) + 1;
}

int map_set_raw_sharded(llb_ctx *c, const float4 *corner, int rc, const float4 *surf, int rs, int rank, int world)
{
    if (world < 1 || world > S2M_MAX_PEERS || rank < 0 || rank >= world) return (int)LLB_ERR_INVALID;
    plan_shard(c, corner, rc, surf, rs, rank, world);
    const ShardPlan pl = c->shard;
    const float reach = std::sqrt(c->prm.knn_max_sqdist);
    c->shardCorner.ensure(std::max(rc, 1)); c->shardSurf.ensure(std::max(rs, 1));
    c->shard_blk.ensure((size_t)div_up(std::max(rc, rs), 1024) + 2);
    c->shard_cnt.ensure(4);
    c->mapCornerDS.ensure(std::max(rc, 1)); c->mapSurfDS.ensure(std::max(rs, 1));
    int ilo, ihi;
    // this rank's part of the two raw maps (stable, whole voxels) ...
    slab_voxels(pl, c->prm.corner_leaf, reach, ilo, ihi);
    c->launches += launch_shard_compact(corner, rc, pl.axis, 1.0f / c->prm.corner_leaf, ilo, ihi, c->shardCorner.p,
                                        c->shard_cnt.p + 0, c->shard_blk.p, c->stream);
    slab_voxels(pl, c->prm.surf_leaf, reach, ilo, ihi);
    c->launches += launch_shard_compact(surf, rs, pl.axis, 1.0f / c->prm.surf_leaf, ilo, ihi, c->shardSurf.p,
                                        c->shard_cnt.p + 1, c->shard_blk.p, c->stream);
    // ... voxel-filtered on the lattice of the WHOLE maps (MO:1057-1064 restricted to the slab) ...
    VoxelInput a; a.a = c->shardCorner.p; a.na_dev = c->shard_cnt.p + 0; a.na = rc;
    VoxelInput ab; ab.a = corner; ab.na = rc;
    VoxelInput b; b.a = c->shardSurf.p; b.na_dev = c->shard_cnt.p + 1; b.na = rs;
    VoxelInput bb; bb.a = surf; bb.na = rs;
    LLB_CUDA(cudaEventRecord(c->fork_ev, c->stream));
    LLB_CUDA(cudaStreamWaitEvent(c->stream2, c->fork_ev, 0));
    c->launches += c->vox2.run_with_bounds(a, ab, c->prm.corner_leaf, c->mapCornerDS.p, c->counts.p + llb_ctx::C_MAP_CORNER_DS, c->stream2);
    LLB_CUDA(cudaEventRecord(c->join_ev, c->stream2));
    c->launches += c->vox.run_with_bounds(b, bb, c->prm.surf_leaf, c->mapSurfDS.p, c->counts.p + llb_ctx::C_MAP_SURF_DS, c->stream);
    LLB_CUDA(cudaStreamWaitEvent(c->stream, c->join_ev, 0));
    c->mapCornerDS_view = c->mapCornerDS.p; c->mapSurfDS_view = c->mapSurfDS.p;
    c->mapCornerDS_upper = rc; c->mapSurfDS_upper = rs;
    c->map_counts_on_dev = true;
    // ... the centroids this rank owns (halo excluded): their sum over the ranks is the size of the unsharded map
    c->launches += launch_shard_count_owned(c->mapCornerDS.p, c->counts.p + llb_ctx::C_MAP_CORNER_DS, rc, pl.axis, pl.lo, pl.hi,
                                            c->shard_cnt.p + 2, c->stream);
    c->launches += launch_shard_count_owned(c->mapSurfDS.p, c->counts.p + llb_ctx::C_MAP_SURF_DS, rs, pl.axis, pl.lo, pl.hi,
                                            c->shard_cnt.p + 3, c->stream);
    c->shard_global[0] = c->shard_global[1] = -1;
    build_indices(c);                                        // MO:1333-1334 on this rank's part only
    return (int)LLB_OK;
}

}  // namespace

int llb_map_set_raw_sharded(llb_ctx *c, const llb_point *corner, int rc, const llb_point *surf, int rs, int rank, int world)
{
    return guarded(c, [&]() {
        if (rc < 0 || rs < 0 || (rc > 0 && !corner) || (rs > 0 && !surf)) return (int)LLB_ERR_INVALID;
        set_cloud(c, 0, c->mapCornerRaw, corner, rc);
        set_cloud(c, 1, c->mapSurfRaw, surf, rs);
        return map_set_raw_sharded(c, c->mapCornerRaw.pts.p, rc, c->mapSurfRaw.pts.p, rs, rank, world);
    });
}

int llb_map_set_raw_sharded_dev(llb_ctx *c, const void *corner, int rc, const void *surf, int rs, int rank, int world)
{
    return guarded(c, [&]() {
        if (rc < 0 || rs < 0 || (rc > 0 && !corner) || (rs > 0 && !surf)) return (int)LLB_ERR_INVALID;
        return map_set_raw_sharded(c, (const float4 *)corner, rc, (const float4 *)surf, rs, rank, world);
    });
}

int llb_map_shard_info(llb_ctx *c, llb_shard_info *out)
{
    return guarded(c, [&]() {
        if (!out) return (int)LLB_ERR_INVALID;
        if (c->shard.axis < 0) return (int)LLB_ERR_STATE;
        int *h = c->pin_counts.p + llb_ctx::C_N;
        LLB_CUDA(cudaMemcpyAsync(h, c->shard_cnt.p, sizeof(int) * 4, cudaMemcpyDeviceToHost, c->stream));
        const int dsn[2] = { read_count(c, llb_ctx::C_MAP_CORNER_DS), c->pin_counts.p[llb_ctx::C_MAP_SURF_DS] };   // (one sync)
        out->axis = c->shard.axis; out->lo = c->shard.lo; out->hi = c->shard.hi; out->rank = c->shard.rank; out->world = c->shard.world;
        for (int k = 0; k < 2; k++) { out->raw_kept[k] = h[k]; out->ds_local[k] = dsn[k]; out->ds_owned[k] = h[2 + k]; }
        return (int)LLB_OK;
    });
}

int llb_map_shard_set_global(llb_ctx *c, const int global_ds[2])
{
    return guarded(c, [&]() {
        if (!global_ds || c->shard.axis < 0) return (int)(global_ds ? LLB_ERR_STATE : LLB_ERR_INVALID);
        c->shard_global[0] = global_ds[0]; c->shard_global[1] = global_ds[1];
        c->s2m.set_shard(c->shard.axis, c->shard.lo, c->shard.hi, global_ds[0], global_ds[1]);
        return (int)LLB_OK;
    });
}

// ------------------------------------------------------------------ loop closure + global map

namespace {

AsmSeg seg_for_pose(const float *p)
{   // transformPointCloud(cloud, &pose) MO:577-606: cos / sin of the float members are the float overloads (utility.h:
    // using namespace std), i.e. the host libm's cosf / sinf as in updateTransformPointCloudSinCos
    AsmSeg sg{};
    sg.ctRoll = cosf(p[0]); sg.stRoll = sinf(p[0]);
    sg.ctPitch = cosf(p[1]); sg.stPitch = sinf(p[1]);
    sg.ctYaw = cosf(p[2]); sg.stYaw = sinf(p[2]);
    sg.tx = p[3]; sg.ty = p[4]; sg.tz = p[5];
    return sg;
}

// key-frame clouds `which` (bit k: cloud k of KeyFrameRec) of ids[] in order, transformed by poses[], back to back into dst
size_t assemble_clouds(llb_ctx *c, const int *ids, const float *poses, int n, unsigned which, DevBuf<float4> &dst)
{
    size_t total = 0;
    for (int k = 0; k < n; k++) {
        const KeyFrameRec &r = c->kfs.rec(ids[k]);
        for (int j = 0; j < 3; j++) if (which & (1u << j)) total += (size_t)r.n[j];
    }
    dst.ensure(std::max<size_t>(total, 1));
    c->asm_segs.ensure(std::max(3 * n, 1));
    if (c->asm_busy) { LLB_CUDA(cudaEventSynchronize(c->asm_ev)); c->asm_busy = false; }
    c->pin_segs.ensure(std::max(3 * n, 1));
    size_t off = 0; int nseg = 0, nmax = 1;
    for (int k = 0; k < n; k++) {
        const KeyFrameRec &r = c->kfs.rec(ids[k]);
        AsmSeg sg = seg_for_pose(poses + 6 * k);
        for (int j = 0; j < 3; j++) {
            if (!(which & (1u << j))) continue;
            sg.src = r.cloud[j]; sg.n = r.n[j]; sg.dst = dst.p + off; off += (size_t)r.n[j];
            if (sg.n > 0) { c->pin_segs.p[nseg++] = sg; nmax = std::max(nmax, sg.n); }
        }
    }
    if (nseg > 0) {
        LLB_CUDA(cudaMemcpyAsync(c->asm_segs.p, c->pin_segs.p, sizeof(AsmSeg) * nseg, cudaMemcpyHostToDevice, c->stream));
        LLB_CUDA(cudaEventRecord(c->asm_ev, c->stream));
        c->asm_busy = true;
        launch_kf_assemble(c->asm_segs.p, nseg, nmax, c->stream);
        c->launches++;
    }
    return total;
}

bool valid_ids(llb_ctx *c, const int *ids, int n)
{
    for (int k = 0; k < n; k++) if (ids[k] < 0 || ids[k] >= c->kfs.size()) return false;
    return true;
}

}  // namespace

void llb_loop_params_default(llb_loop_params *p)
{
    if (!p) return;
    p->max_iterations = 100; p->max_correspondence_distance = 100.0;          // MO:893-894
    p->transformation_epsilon = 1e-6; p->euclidean_fitness_epsilon = 1e-6;    // MO:895-896
}

int llb_loop_set_clouds(llb_ctx *c, int latest_id, const float latest_pose[6], const int *hist_ids, const float *hist_poses,
                        int n_hist, float history_leaf, int counts[2])
{
    return guarded(c, [&]() {
        if (!latest_pose || n_hist < 0 || (n_hist > 0 && (!hist_ids || !hist_poses)) || !(history_leaf > 0.f)) return (int)LLB_ERR_INVALID;
        if (!valid_ids(c, &latest_id, 1) || !valid_ids(c, hist_ids, n_hist)) return (int)LLB_ERR_INVALID;
        // MO:840-851: corner + surf clouds of the latest key-frame at its pose, points with (int)intensity >= 0
        const size_t nl = assemble_clouds(c, &latest_id, latest_pose, 1, 3u, c->loopLatestRaw);
        c->loopLatest.ensure(std::max<size_t>(nl, 1));
        c->launches += launch_loop_filter_intensity(c->loopLatestRaw.p, (int)nl, c->loopLatest.p, c->counts.p + llb_ctx::C_LOOP_LATEST, c->stream);
        c->loop_n_latest_raw = (int)nl;
        // MO:853-861: corner + surf clouds of the history frames around the closest one, VoxelGrid(history_leaf)
        const size_t nh = assemble_clouds(c, hist_ids, hist_poses, n_hist, 3u, c->loopHistRaw);
        if (nh > (size_t)INT_MAX) return (int)LLB_ERR_CAPACITY;
        c->loopHistDS.ensure(std::max<size_t>(nh, 1));
        VoxelInput in; in.a = c->loopHistRaw.p; in.na = (int)nh;
        c->launches += c->vox3.run(in, history_leaf, c->loopHistDS.p, c->counts.p + llb_ctx::C_LOOP_HIST_DS, c->stream);
        c->loop_n_hist_raw = (int)nh;
        c->loop_n_latest = read_count(c, llb_ctx::C_LOOP_LATEST);
        c->loop_n_hist = c->pin_counts.p[llb_ctx::C_LOOP_HIST_DS];
        if (counts) { counts[0] = c->loop_n_latest; counts[1] = c->loop_n_hist; }
        return (int)LLB_OK;
    });
}

int llb_loop_set_clouds_host(llb_ctx *c, const llb_point *latest, int n_latest, const llb_point *history_ds, int n_hist)
{
    return guarded(c, [&]() {
        if (n_latest < 0 || n_hist < 0 || (n_latest > 0 && !latest) || (n_hist > 0 && !history_ds)) return (int)LLB_ERR_INVALID;
        upload_cloud(c, 0, latest, n_latest, c->loopLatest);
        upload_cloud(c, 1, history_ds, n_hist, c->loopHistDS);
        c->loop_n_latest = n_latest; c->loop_n_hist = n_hist; c->loop_n_latest_raw = 0; c->loop_n_hist_raw = 0;
        return (int)LLB_OK;
    });
}

int llb_loop_icp(llb_ctx *c, const llb_loop_params *prm, llb_icp_result *out)
{
    return guarded(c, [&]() {
        if (!out) return (int)LLB_ERR_INVALID;
        if (c->loop_n_latest < 0 || c->loop_n_hist < 0) return (int)LLB_ERR_STATE;
        llb_loop_params p; llb_loop_params_default(&p);
        if (prm) p = *prm;
        if (p.max_iterations < 1 || !(p.max_correspondence_distance > 0)) return (int)LLB_ERR_INVALID;
        IcpParams ip{ p.max_iterations, p.max_correspondence_distance, p.transformation_epsilon, p.euclidean_fitness_epsilon };
        c->pin_icp.ensure(1);
        LLB_CUDA(cudaEventRecord(c->ev0, c->stream));
        c->launches += c->icp.run(ip, c->loopLatest.p, c->loop_n_latest, c->loopHistDS.p, c->loop_n_hist, 0, c->stream);
        LLB_CUDA(cudaEventRecord(c->ev1, c->stream));
        LLB_CUDA(cudaMemcpyAsync(c->pin_icp.p, c->icp.state_dev(), sizeof(IcpState), cudaMemcpyDeviceToHost, c->stream));
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        const IcpState &s = *c->pin_icp.p;
        for (int i = 0; i < 16; i++) out->T[i] = s.T[i];
        out->has_converged = s.converged; out->iterations = s.iterations; out->convergence_state = s.state;
        out->n_correspondences = s.n_corr; out->fitness_score = s.fitness;
        out->n_source = c->loop_n_latest; out->n_target = c->loop_n_hist;
        for (int i = 0; i < 17; i++) out->sums[i] = s.sums[i];
        LLB_CUDA(cudaEventElapsedTime(&out->device_ms, c->ev0, c->ev1));
        return (int)LLB_OK;
    });
}

int llb_loop_get_cloud(llb_ctx *c, int which, llb_point *out, int cap, int *n)
{
    return guarded(c, [&]() {
        if (!n || which < 0 || which > 3) return (int)LLB_ERR_INVALID;
        const int cnt = which == 0 ? c->loop_n_latest : which == 1 ? c->loop_n_hist_raw : which == 2 ? c->loop_n_hist : c->global_n;
        if (cnt < 0) return (int)LLB_ERR_STATE;
        *n = cnt;
        if (!out) return (int)LLB_OK;
        if (cnt > cap) return (int)LLB_ERR_CAPACITY;
        download_cloud(c, which == 0 ? c->loopLatest.p : which == 1 ? c->loopHistRaw.p : which == 2 ? c->loopHistDS.p : c->globalDS.p, cnt, out);
        return (int)LLB_OK;
    });
}

int llb_loop_get_nn(llb_ctx *c, int *idx, float *sqdist, int cap, int *n)
{
    return guarded(c, [&]() {
        if (!n) return (int)LLB_ERR_INVALID;
        if (c->loop_n_latest < 0) return (int)LLB_ERR_STATE;
        *n = c->loop_n_latest;
        if (!idx && !sqdist) return (int)LLB_OK;
        if (*n > cap) return (int)LLB_ERR_CAPACITY;
        std::vector<unsigned long long> h((size_t)std::max(*n, 1));
        LLB_CUDA(cudaMemcpyAsync(h.data(), c->icp.nn_dev(), sizeof(unsigned long long) * (size_t)*n, cudaMemcpyDeviceToHost, c->stream));
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < *n; i++) {
            const unsigned hi = (unsigned)(h[i] >> 32);
            if (idx) idx[i] = (int)(h[i] & 0xffffffffu);
            if (sqdist) std::memcpy(&sqdist[i], &hi, 4);
        }
        return (int)LLB_OK;
    });
}

int llb_global_map_assemble(llb_ctx *c, const int *ids, const float *poses, int n, float leaf, int *n_out)
{
    return guarded(c, [&]() {
        if (n < 0 || (n > 0 && (!ids || !poses)) || !(leaf > 0.f)) return (int)LLB_ERR_INVALID;
        if (!valid_ids(c, ids, n)) return (int)LLB_ERR_INVALID;
        // MO:780-788: corner, surf and outlier clouds of every selected key-frame at its pose, then one VoxelGrid(leaf)
        const size_t tot = assemble_clouds(c, ids, poses, n, 7u, c->globalRaw);
        if (tot > (size_t)INT_MAX) return (int)LLB_ERR_CAPACITY;
        c->globalDS.ensure(std::max<size_t>(tot, 1));
        VoxelInput in; in.a = c->globalRaw.p; in.na = (int)tot;
        c->launches += c->vox3.run(in, leaf, c->globalDS.p, c->counts.p + llb_ctx::C_GLOBAL_DS, c->stream);
        c->global_n = read_count(c, llb_ctx::C_GLOBAL_DS);
        if (n_out) *n_out = c->global_n;
        return (int)LLB_OK;
    });
}

// ------------------------------------------------------------------ key-frame store

int llb_keyframe_add(llb_ctx *c, int *id)
{
    return guarded(c, [&]() {
        if (!c->scan_ds_done) return (int)LLB_ERR_STATE;
        read_count(c, 0);                                     // sizes of the DS clouds (synchronises the stream)
        const int n[3] = { c->pin_counts.p[llb_ctx::C_CORNER_DS], c->pin_counts.p[llb_ctx::C_SURF_DS],
                           c->pin_counts.p[llb_ctx::C_OUTLIER_DS] };
        const float4 *src[3] = { c->cornerLastDS.p, c->surfLastDS.p, c->outlierLastDS.p };
        float4 *dst[3];
        const int k = c->kfs.add(n, dst);
        for (int j = 0; j < 3; j++)
            if (n[j] > 0) LLB_CUDA(cudaMemcpyAsync(dst[j], src[j], sizeof(float4) * n[j], cudaMemcpyDeviceToDevice, c->stream));
        if (id) *id = k;
        return (int)LLB_OK;
    });
}

int llb_keyframe_add_clouds(llb_ctx *c, const llb_point *corner, int nc, const llb_point *surf, int ns,
                            const llb_point *outlier, int no, int *id)
{
    return guarded(c, [&]() {
        if (nc < 0 || ns < 0 || no < 0 || (nc > 0 && !corner) || (ns > 0 && !surf) || (no > 0 && !outlier))
            return (int)LLB_ERR_INVALID;
        const int n[3] = { nc, ns, no };
        const llb_point *src[3] = { corner, surf, outlier };
        float4 *dst[3];
        const int k = c->kfs.add(n, dst);
        for (int j = 0; j < 3; j++) {
            if (n[j] <= 0) continue;
            upload_cloud(c, j, src[j], n[j], c->tmp_in);       // staged + unpacked, then placed into the arena
            LLB_CUDA(cudaMemcpyAsync(dst[j], c->tmp_in.p, sizeof(float4) * n[j], cudaMemcpyDeviceToDevice, c->stream));
        }
        if (id) *id = k;
        return (int)LLB_OK;
    });
}

int llb_keyframe_count(llb_ctx *c, int *n)
{
    if (!c || !n) return LLB_ERR_INVALID;
    *n = c->kfs.size();
    return LLB_OK;
}

int llb_keyframe_clear(llb_ctx *c)
{
    return guarded(c, [&]() {
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        c->kfs.clear();
        return (int)LLB_OK;
    });
}

int llb_map_assemble(llb_ctx *c, const int *ids, const float *poses, int n)
{
    return guarded(c, [&]() {
        if (n < 0 || (n > 0 && (!ids || !poses))) return (int)LLB_ERR_INVALID;
        size_t rc = 0, rs = 0;
        for (int k = 0; k < n; k++) {
            if (ids[k] < 0 || ids[k] >= c->kfs.size()) return (int)LLB_ERR_INVALID;
            const KeyFrameRec &r = c->kfs.rec(ids[k]);
            rc += r.n[0]; rs += (size_t)r.n[1] + r.n[2];
        }
        if (rc > (size_t)INT_MAX || rs > (size_t)INT_MAX) return (int)LLB_ERR_CAPACITY;
        c->asmCorner.ensure(std::max<size_t>(rc, 1)); c->asmSurf.ensure(std::max<size_t>(rs, 1));
        c->asm_segs.ensure(std::max(3 * n, 1));
        if (c->asm_busy) { LLB_CUDA(cudaEventSynchronize(c->asm_ev)); c->asm_busy = false; }
        c->pin_segs.ensure(std::max(3 * n, 1));
        size_t oc = 0, os = 0;
        int nseg = 0, nmax = 1;
        for (int k = 0; k < n; k++) {
            const KeyFrameRec &r = c->kfs.rec(ids[k]);
            const float *p = poses + 6 * k;                  // PointTypePose: roll, pitch, yaw, x, y, z
            AsmSeg sg{};
            // updateTransformPointCloudSinCos MO:529-543: float overloads of cos / sin of the host libm, as the
            // reference (std::cos(float)); the device only multiplies and adds
            sg.ctRoll = cosf(p[0]); sg.stRoll = sinf(p[0]);
            sg.ctPitch = cosf(p[1]); sg.stPitch = sinf(p[1]);
            sg.ctYaw = cosf(p[2]); sg.stYaw = sinf(p[2]);
            sg.tx = p[3]; sg.ty = p[4]; sg.tz = p[5];
            // MO:1050-1054: corner map += corner_k; surf map += surf_k; surf map += outlier_k
            sg.src = r.cloud[0]; sg.n = r.n[0]; sg.dst = c->asmCorner.p + oc; oc += r.n[0];
            if (sg.n > 0) { c->pin_segs.p[nseg++] = sg; nmax = std::max(nmax, sg.n); }
            sg.src = r.cloud[1]; sg.n = r.n[1]; sg.dst = c->asmSurf.p + os; os += r.n[1];
            if (sg.n > 0) { c->pin_segs.p[nseg++] = sg; nmax = std::max(nmax, sg.n); }
            sg.src = r.cloud[2]; sg.n = r.n[2]; sg.dst = c->asmSurf.p + os; os += r.n[2];
            if (sg.n > 0) { c->pin_segs.p[nseg++] = sg; nmax = std::max(nmax, sg.n); }
        }
        if (nseg > 0) {
            LLB_CUDA(cudaMemcpyAsync(c->asm_segs.p, c->pin_segs.p, sizeof(AsmSeg) * nseg, cudaMemcpyHostToDevice, c->stream));
            LLB_CUDA(cudaEventRecord(c->asm_ev, c->stream));
            c->asm_busy = true;
            launch_kf_assemble(c->asm_segs.p, nseg, nmax, c->stream);
            c->launches++;
        }
        c->asm_rc = (int)rc; c->asm_rs = (int)rs;
        voxel_map_raw(c, c->asmCorner.p, (int)rc, c->asmSurf.p, (int)rs);      // MO:1057-1064
        build_indices(c);                                                      // MO:1333-1334
        return (int)LLB_OK;
    });
}

int llb_map_get_raw(llb_ctx *c, int which, llb_point *out, int cap, int *n)
{
    return guarded(c, [&]() {
        if (!n || which < 0 || which > 1) return (int)LLB_ERR_INVALID;
        const int cnt = which == 0 ? c->asm_rc : c->asm_rs;
        *n = cnt;
        if (!out) return (int)LLB_OK;
        if (cnt > cap) return (int)LLB_ERR_CAPACITY;
        download_cloud(c, which == 0 ? c->asmCorner.p : c->asmSurf.p, cnt, out);
        return (int)LLB_OK;
    });
}

// ------------------------------------------------------------------ featureAssociation

int llb_odom_set_last(llb_ctx *c, const llb_point *corner, int ncl, const llb_point *surf, int nsl)
{
    return guarded(c, [&]() {
        if (ncl < 0 || nsl < 0 || (ncl > 0 && !corner) || (nsl > 0 && !surf)) return (int)LLB_ERR_INVALID;
        upload_cloud(c, 0, corner, ncl, c->odom.cornerLast());
        upload_cloud(c, 1, surf, nsl, c->odom.surfLast());
        c->launches += c->odom.set_last(ncl, nsl, c->stream);
        return (int)LLB_OK;
    });
}

int llb_odom_set_features(llb_ctx *c, const llb_point *sharp, int nsharp, const llb_point *flat, int nflat)
{
    return guarded(c, [&]() {
        if (nsharp < 0 || nflat < 0 || (nsharp > 0 && !sharp) || (nflat > 0 && !flat)) return (int)LLB_ERR_INVALID;
        upload_cloud(c, 0, sharp, nsharp, c->odom.sharp());
        upload_cloud(c, 1, flat, nflat, c->odom.flat());
        c->odom.set_features(nsharp, nflat);
        return (int)LLB_OK;
    });
}

// ---------------------------------------------------------------- imageProjection
int llb_projection_init(llb_ctx *c, int n_scan, int horizon_scan, float ang_res_x, float ang_res_y, int ground_scan_ind)
{
    return guarded(c, [&]() {
        if (n_scan <= 0 || n_scan > FE_MAX_RINGS || horizon_scan < 16 || horizon_scan > 4096 || !(ang_res_x > 0.f) || !(ang_res_y > 0.f) ||
            ground_scan_ind < 0 || ground_scan_ind >= n_scan) return (int)LLB_ERR_INVALID;
        c->projection.init(n_scan, horizon_scan, ang_res_x, ang_res_y, ground_scan_ind, c->stream);
        c->projection_done = false;
        return (int)LLB_OK;
    });
}

int llb_projection_process(llb_ctx *c, const llb_point *cloud, const uint16_t *ring, int n, int *n_seg, int *n_out, float *device_ms)
{
    return guarded(c, [&]() {
        if (n < 0 || (n > 0 && (!cloud || !ring))) return (int)LLB_ERR_INVALID;
        if (!c->projection.ready()) return (int)LLB_ERR_STATE;
        LLB_CUDA(cudaEventRecord(c->ev0, c->stream));
        c->launches += c->projection.process(reinterpret_cast<const float *>(cloud), ring, n, c->stream);
        LLB_CUDA(cudaEventRecord(c->ev1, c->stream));
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        c->projection_done = true;
        if (n_seg) *n_seg = c->projection.header().n_seg;
        if (n_out) *n_out = c->projection.header().n_outlier;
        if (device_ms) LLB_CUDA(cudaEventElapsedTime(device_ms, c->ev0, c->ev1));
        return (int)LLB_OK;
    });
}

int llb_projection_get_cloud(llb_ctx *c, int which, llb_point *out, int cap, int *n)
{
    return guarded(c, [&]() {
        if (!n || which < 0 || which > 1) return (int)LLB_ERR_INVALID;
        if (!c->projection_done) return (int)LLB_ERR_STATE;
        const int cnt = which == 0 ? c->projection.header().n_seg : c->projection.header().n_outlier;
        *n = cnt;
        if (!out) return (int)LLB_OK;
        if (cnt > cap) return (int)LLB_ERR_CAPACITY;
        if (cnt > 0) download_cloud(c, which == 0 ? c->projection.seg_dev() : c->projection.outlier_dev(), cnt, out);
        return (int)LLB_OK;
    });
}

int llb_projection_get_info(llb_ctx *c, int32_t *start_ring, int32_t *end_ring, float ori[3], uint8_t *ground, uint32_t *col,
                            float *range, int cap)
{
    return guarded(c, [&]() {
        if (!c->projection_done) return (int)LLB_ERR_STATE;
        const ImageProjector &ip = c->projection;
        const int ns = ip.params().n_scan, n = ip.header().n_seg;
        if ((ground || col || range) && n > cap) return (int)LLB_ERR_CAPACITY;
        if (start_ring) std::memcpy(start_ring, ip.start_ring_host(), sizeof(int) * ns);
        if (end_ring) std::memcpy(end_ring, ip.end_ring_host(), sizeof(int) * ns);
        if (ori) { ori[0] = ip.header().start_ori; ori[1] = ip.header().end_ori; ori[2] = ip.header().ori_diff; }
        if (n > 0) {
            if (ground) LLB_CUDA(cudaMemcpyAsync(ground, ip.ground_flag_dev(), (size_t)n, cudaMemcpyDeviceToHost, c->stream));
            if (col) LLB_CUDA(cudaMemcpyAsync(col, ip.col_ind_dev(), sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream));
            if (range) LLB_CUDA(cudaMemcpyAsync(range, ip.seg_range_dev(), sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
            LLB_CUDA(cudaStreamSynchronize(c->stream));
        }
        return (int)LLB_OK;
    });
}

int llb_projection_get_images(llb_ctx *c, float *range_mat, int8_t *ground_mat, int32_t *label_mat)
{
    return guarded(c, [&]() {
        if (!c->projection_done) return (int)LLB_ERR_STATE;
        const ImageProjector &ip = c->projection;
        const size_t np = (size_t)ip.params().n_scan * ip.params().horizon;
        if (range_mat) LLB_CUDA(cudaMemcpyAsync(range_mat, ip.range_mat_dev(), sizeof(float) * np, cudaMemcpyDeviceToHost, c->stream));
        if (ground_mat) LLB_CUDA(cudaMemcpyAsync(ground_mat, ip.ground_mat_dev(), np, cudaMemcpyDeviceToHost, c->stream));
        if (label_mat) LLB_CUDA(cudaMemcpyAsync(label_mat, ip.label_mat_dev(), sizeof(int) * np, cudaMemcpyDeviceToHost, c->stream));
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        return (int)LLB_OK;
    });
}

int llb_projection_to_features(llb_ctx *c, int counts[4], float *device_ms)
{
    return guarded(c, [&]() {
        if (!c->projection_done || !c->features.ready()) return (int)LLB_ERR_STATE;
        const ImageProjector &ip = c->projection;
        if (ip.params().n_scan != c->features.n_scan() || ip.params().horizon != c->features.horizon()) return (int)LLB_ERR_STATE;
        const IpHeader &h = ip.header();
        // ring bounds must stay inside the cloud (IP:318, IP:358): the feature kernels index with them
        for (int r = 0; r < ip.params().n_scan; r++)
            if (ip.start_ring_host()[r] < 4 || ip.end_ring_host()[r] > h.n_seg - 6 ||
                ip.end_ring_host()[r] - ip.start_ring_host()[r] > ip.params().horizon) return (int)LLB_ERR_INVALID;
        LLB_CUDA(cudaEventRecord(c->ev0, c->stream));
        c->launches += c->features.extract_dev(ip.seg_dev(), h.n_seg, ip.start_ring_dev(), ip.end_ring_dev(), h.start_ori, h.end_ori,
                                               h.ori_diff, ip.ground_flag_dev(), ip.col_ind_dev(), ip.seg_range_dev(), c->stream);
        LLB_CUDA(cudaEventRecord(c->ev1, c->stream));
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        LLB_CUDA(cudaEventElapsedTime(&c->features_ms, c->ev0, c->ev1));
        c->features_done = true;
        if (counts) for (int k = 0; k < 4; k++) counts[k] = c->features.counts()[k];
        if (device_ms) *device_ms = c->features_ms;
        return (int)LLB_OK;
    });
}

// ---------------------------------------------------------------- feature extraction
int llb_features_init(llb_ctx *c, int n_scan, int horizon_scan)
{
    return guarded(c, [&]() {
        if (n_scan <= 0 || n_scan > FE_MAX_RINGS || horizon_scan < 16 || horizon_scan > 4096) return (int)LLB_ERR_INVALID;
        c->features.init(n_scan, horizon_scan, c->stream);
        c->features_done = false;
        return (int)LLB_OK;
    });
}

int llb_features_extract(llb_ctx *c, const llb_segmented_cloud *seg, int counts[4], float *device_ms)
{
    return guarded(c, [&]() {
        if (!seg || seg->n < 0) return (int)LLB_ERR_INVALID;
        if (!c->features.ready()) return (int)LLB_ERR_STATE;
        if (seg->n > c->features.n_scan() * c->features.horizon()) return (int)LLB_ERR_CAPACITY;
        if (!seg->start_ring || !seg->end_ring) return (int)LLB_ERR_INVALID;
        if (seg->n > 0 && (!seg->cloud || !seg->ground_flag || !seg->col_ind || !seg->range)) return (int)LLB_ERR_INVALID;
        // ring bounds must stay inside the cloud (IP:318, IP:358): the kernels index with them
        for (int r = 0; r < c->features.n_scan(); r++)
            if (seg->start_ring[r] < 4 || seg->end_ring[r] > seg->n - 6 || seg->end_ring[r] - seg->start_ring[r] > c->features.horizon())
                return (int)LLB_ERR_INVALID;
        LLB_CUDA(cudaEventRecord(c->ev0, c->stream));
        c->launches += c->features.extract(reinterpret_cast<const float *>(seg->cloud), seg->n, seg->start_ring, seg->end_ring,
                                           seg->start_orientation, seg->end_orientation, seg->orientation_diff,
                                           seg->ground_flag, seg->col_ind, seg->range, c->stream);
        LLB_CUDA(cudaEventRecord(c->ev1, c->stream));
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        LLB_CUDA(cudaEventElapsedTime(&c->features_ms, c->ev0, c->ev1));
        c->features_done = true;
        if (counts) for (int k = 0; k < 4; k++) counts[k] = c->features.counts()[k];
        if (device_ms) *device_ms = c->features_ms;
        return (int)LLB_OK;
    });
}

int llb_features_get(llb_ctx *c, int which, llb_point *out, int cap, int *n)
{
    return guarded(c, [&]() {
        if (!n || which < 0 || which > 6) return (int)LLB_ERR_INVALID;
        if (!c->features_done) return (int)LLB_ERR_STATE;
        if (which >= 5) {
            if (c->features_last_n[which - 5] < 0) return (int)LLB_ERR_STATE;
            *n = c->features_last_n[which - 5];
            if (!out) return (int)LLB_OK;
            if (*n > cap) return (int)LLB_ERR_CAPACITY;
            download_cloud(c, which == 5 ? c->odom.cornerLast().p : c->odom.surfLast().p, *n, out);
            return (int)LLB_OK;
        }
        const int cnt = which == 4 ? c->features.n_points() : c->features.counts()[which];
        *n = cnt;
        if (!out) return (int)LLB_OK;
        if (cnt > cap) return (int)LLB_ERR_CAPACITY;
        if (which == 4) { download_cloud(c, c->features.dev_cloud(4), cnt, out); return (int)LLB_OK; }
        const float4 *src = c->features.host_cloud(which);
        for (int i = 0; i < cnt; i++) {
            out[i].x = src[i].x; out[i].y = src[i].y; out[i].z = src[i].z; out[i].w = 1.0f;
            out[i].intensity = src[i].w; out[i].c1 = out[i].c2 = out[i].c3 = 0.f;
        }
        return (int)LLB_OK;
    });
}

int llb_features_get_state(llb_ctx *c, float *curvature, int *neighbor_picked, int *label, int cap)
{
    return guarded(c, [&]() {
        if (!curvature || !neighbor_picked || !label) return (int)LLB_ERR_INVALID;
        if (!c->features_done) return (int)LLB_ERR_STATE;
        if (c->features.n_points() > cap) return (int)LLB_ERR_CAPACITY;
        c->features.get_state(curvature, neighbor_picked, label, c->features.n_points(), c->stream);
        return (int)LLB_OK;
    });
}

static_assert(sizeof(llb_imu_queue) == sizeof(FeImu) && LLB_IMU_QUEUE == FE_IMU_QUEUE, "llb_imu_queue mirrors FeImu");
static_assert(sizeof(llb_imu_sweep) == sizeof(FeImuOut) && sizeof(llb_imu_end) == sizeof(FeEndImu), "IMU structs mirror the device ones");

int llb_features_set_imu(llb_ctx *c, const llb_imu_queue *queue)
{
    return guarded(c, [&]() {
        if (!c->features.ready()) return (int)LLB_ERR_STATE;
        if (queue && (queue->pointer_last >= LLB_IMU_QUEUE || queue->pointer_last_iteration < 0 ||
                      queue->pointer_last_iteration >= LLB_IMU_QUEUE)) return (int)LLB_ERR_INVALID;
        c->features.set_imu(reinterpret_cast<const FeImu *>(queue));
        return (int)LLB_OK;
    });
}

int llb_features_get_imu(llb_ctx *c, llb_imu_sweep *out)
{
    return guarded(c, [&]() {
        if (!out) return (int)LLB_ERR_INVALID;
        if (!c->features_done) return (int)LLB_ERR_STATE;
        if (c->features.has_imu_out()) std::memcpy(out, &c->features.imu_out(), sizeof(*out));
        else std::memset(out, 0, sizeof(*out));
        return (int)LLB_OK;
    });
}

int llb_features_publish_last(llb_ctx *c, const float transformCur[6])
{
    return llb_features_publish_last_imu(c, transformCur, nullptr);
}

int llb_features_publish_last_imu(llb_ctx *c, const float transformCur[6], const llb_imu_end *imu)
{
    return guarded(c, [&]() {
        if (!transformCur) return (int)LLB_ERR_INVALID;
        if (!c->features_done) return (int)LLB_ERR_STATE;
        const int ncl = c->features.counts()[1], nsl = c->features.counts()[3];
        c->odom.cornerLast().ensure(std::max(ncl, 1)); c->odom.surfLast().ensure(std::max(nsl, 1));
        c->launches += c->features.transform_to_end(transformCur, c->odom.cornerLast().p, c->odom.surfLast().p, c->stream,
                                                    reinterpret_cast<const FeEndImu *>(imu));
        c->launches += c->odom.set_last(ncl, nsl, c->stream);
        c->features_last_n[0] = ncl; c->features_last_n[1] = nsl;
        return (int)LLB_OK;
    });
}

int llb_features_get_profile(llb_ctx *c, int cycles[10])
{
    return guarded(c, [&]() {
        if (!cycles) return (int)LLB_ERR_INVALID;
        if (!c->features_done) return (int)LLB_ERR_STATE;
        for (int k = 0; k < 10; k++) cycles[k] = c->features.phase_cycles()[k];
        return (int)LLB_OK;
    });
}

int llb_features_to_odometry(llb_ctx *c)
{
    return guarded(c, [&]() {
        if (!c->features_done) return (int)LLB_ERR_STATE;
        const int ns = c->features.counts()[0], nf = c->features.counts()[2];
        c->odom.sharp().ensure(std::max(ns, 1)); c->odom.flat().ensure(std::max(nf, 1));
        if (ns > 0) LLB_CUDA(cudaMemcpyAsync(c->odom.sharp().p, c->features.dev_cloud(0), sizeof(float4) * ns, cudaMemcpyDeviceToDevice, c->stream));
        if (nf > 0) LLB_CUDA(cudaMemcpyAsync(c->odom.flat().p, c->features.dev_cloud(2), sizeof(float4) * nf, cudaMemcpyDeviceToDevice, c->stream));
        c->odom.set_features(ns, nf);
        return (int)LLB_OK;
    });
}

static void fill_odom_stats(llb_stats *st, const OdomState &s, int which, float ms)
{
    if (!st) return;
    std::memset(st, 0, sizeof(*st));
    st->iterations = s.iters[which]; st->converged = s.converged[which]; st->n_correspondences = s.n_corr;
    st->is_degenerate = s.is_degenerate; st->skipped = s.skipped; st->device_ms = ms;
}

int llb_odom_optimize(llb_ctx *c, float T[6], llb_stats *st_surf, llb_stats *st_corner)
{
    return guarded(c, [&]() {
        if (!T) return (int)LLB_ERR_INVALID;
        if (!c->odom.ready()) return (int)LLB_ERR_STATE;
        LLB_CUDA(cudaEventRecord(c->ev0, c->stream));
        c->launches += c->odom.optimize(T, c->stream);
        LLB_CUDA(cudaEventRecord(c->ev1, c->stream));
        LLB_CUDA(cudaMemcpyAsync(c->pin_ostate.p, c->odom.state_dev(), sizeof(OdomState), cudaMemcpyDeviceToHost, c->stream));
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        float ms = 0.f;
        LLB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        const OdomState &s = *c->pin_ostate.p;
        for (int i = 0; i < 6; i++) T[i] = s.T[i];
        fill_odom_stats(st_surf, s, 0, ms); fill_odom_stats(st_corner, s, 1, ms);
        return (int)LLB_OK;
    });
}

int llb_odom_iterate(llb_ctx *c, int which, float T[6], int iter, int *more, int *n_corr)
{
    return guarded(c, [&]() {
        if (!T || which < 0 || which > 1 || iter < 0) return (int)LLB_ERR_INVALID;
        if (!c->odom.ready()) return (int)LLB_ERR_STATE;
        c->launches += c->odom.iterate(which, T, iter, c->stream);
        LLB_CUDA(cudaMemcpyAsync(c->pin_ostate.p, c->odom.state_dev(), sizeof(OdomState), cudaMemcpyDeviceToHost, c->stream));
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        const OdomState &s = *c->pin_ostate.p;
        for (int i = 0; i < 6; i++) T[i] = s.T[i];
        if (more) *more = s.more;
        if (n_corr) *n_corr = s.n_corr;
        return (int)LLB_OK;
    });
}

int llb_odom_get_correspondences(llb_ctx *c, llb_point *ori, llb_point *coeff, int cap, int *n)
{
    return guarded(c, [&]() {
        if (!n) return (int)LLB_ERR_INVALID;
        std::vector<float4> o, k;
        c->odom.download_correspondences(o, k, c->stream);
        *n = (int)o.size();
        if (!ori || !coeff) return (int)LLB_OK;
        if ((int)o.size() > cap) return (int)LLB_ERR_CAPACITY;
        for (size_t i = 0; i < o.size(); i++) {
            ori[i] = llb_point{ o[i].x, o[i].y, o[i].z, 1.0f, o[i].w, 0.f, 0.f, 0.f };
            coeff[i] = llb_point{ k[i].x, k[i].y, k[i].z, 1.0f, k[i].w, 0.f, 0.f, 0.f };
        }
        return (int)LLB_OK;
    });
}

int llb_odom_get_search_ind(llb_ctx *c, int which, float *i1, float *i2, float *i3, int cap, int *n)
{
    return guarded(c, [&]() {
        if (!n || which < 0 || which > 1) return (int)LLB_ERR_INVALID;
        std::vector<float> a, b, d;
        c->odom.download_search_ind(which, a, b, d, c->stream);
        *n = (int)a.size();
        if (!i1 || !i2) return (int)LLB_OK;
        if ((int)a.size() > cap) return (int)LLB_ERR_CAPACITY;
        std::copy(a.begin(), a.end(), i1); std::copy(b.begin(), b.end(), i2);
        if (i3 && which == 0) std::copy(d.begin(), d.end(), i3);
        return (int)LLB_OK;
    });
}

int llb_odom_get_degeneracy(llb_ctx *c, int *deg, float matP[9])
{
    return guarded(c, [&]() {
        LLB_CUDA(cudaMemcpyAsync(c->pin_ostate.p, c->odom.state_dev(), sizeof(OdomState), cudaMemcpyDeviceToHost, c->stream));
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        if (deg) *deg = c->pin_ostate.p->is_degenerate;
        if (matP) std::memcpy(matP, c->pin_ostate.p->matP, sizeof(float) * 9);
        return (int)LLB_OK;
    });
}

}  // extern "C"
