// batch_capi.cu — C ABI of the batched multi-registration engine (llb_batch_* in include/llb200.h):
// B independent sequences on one GPU, one stream, a fixed number of launches per step (batch.cuh).
// No CPU fallback: llb_batch_create fails with LLB_ERR_NO_DEVICE without an sm_100 device.
#include "../../include/llb200.h"
#include "batch.cuh"
#include "features.cuh"
#include "odom.cuh"

#include <cstring>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>

using namespace llb;

namespace {

// everything the kernels of one step read that the host decides: uploaded with ONE H2D copy per step
struct StepLayout {
    size_t off_scan_n, off_map_n, off_poses, off_small1, off_small2, off_grid, off_regs, off_unpack, off_vox, off_copy, bytes;
    explicit StepLayout(int B)
    {
        size_t o = 0;
        auto take = [&](size_t n) { size_t r = o; o = (o + n + 255) & ~(size_t)255; return r; };
        off_scan_n = take(sizeof(int) * 3 * B);
        off_map_n = take(sizeof(int) * 2 * B);
        off_poses = take(sizeof(float) * 6 * B);
        off_small1 = take(sizeof(SmallJob) * 3 * B);
        off_small2 = take(sizeof(SmallJob) * B);
        off_grid = take(sizeof(GridJob) * 2 * B);
        off_regs = take(sizeof(BatchReg) * B);
        off_unpack = take(sizeof(BatchUnpack) * 5 * B);
        off_vox = take(sizeof(LargeVoxelJob) * 2 * B);
        off_copy = take(sizeof(BatchCopy) * 3 * B);
        bytes = o;
    }
};

struct StepSig { int nunp, unp_max, ncopy, copy_max, nseg, seg_max, nvox, raw_max, ngrid, map_n_max, vox_cap1, vox_cap2; };

constexpr int RING = 3;
constexpr int PROF_N = 6;      // unpack+copies, downsample, index, knn, fit, solve(+prepare/collect)

}  // namespace

struct llb_batch {
    int device = 0;
    cudaStream_t stream = nullptr;
    llb_params prm{};
    S2mParams sprm{};
    std::string err;
    long long launches = 0;
    FeatureBatch features;               // feature extraction of the slots
    bool features_done = false;
    int vox_cap1 = 1024, vox_cap2 = 1024;
    int B = 0, cap_scan = 0, cap_map = 0, qcap = 0, grid_ctas = 0;
    StepLayout lay{ 1 };

    // per-slot device buffers, sliced out of big allocations
    DevBuf<float4> scan_in;      // [B][3][cap_scan]  own copies of host uploads
    DevBuf<float> scan_raw;      // [B][3][cap_scan*8]
    DevBuf<float4> scan_ds;      // [B][5*cap_scan]: cornerDS | surfDS | outlierDS | surfTotalDS (2*cap_scan)
    DevBuf<float4> map_in;       // [B][2][cap_map]
    DevBuf<float> map_raw;       // [B][2][cap_map*8]
    DevBuf<int> ds_n;            // [B][4]
    std::vector<GridIndex> grids;   // 2B
    DevBuf<S2mState> states;
    DevBuf<int> qperm; DevBuf<double> partials; DevBuf<float4> qprev; DevBuf<BatchQueue> queue; DevBuf<unsigned> ctl;
    int lm_grid = 0, part_stride = 0;
    DevBuf<long long> iter_prof;         // debug: LLB_ITER_PROF
    DevBuf<BatchSlotInfo> slot_info;
    DevBuf<BatchResult> results;
    PinnedBuf<BatchResult> pin_results;
    DevBuf<unsigned char> step_dev;
    PinnedBuf<unsigned char> step_pin[RING];
    cudaEvent_t step_ev[RING] = {};
    bool step_busy[RING] = {};
    int ring_pos = 0;

    // host-side slot state
    struct Slot {
        const float4 *scan[3] = { nullptr, nullptr, nullptr }; int scan_n[3] = { 0, 0, 0 };
        const float4 *map[2] = { nullptr, nullptr }; int map_n[2] = { 0, 0 };
        bool scan_set = false, map_set = false, map_dirty = false;
    };
    std::vector<Slot> slots;
    std::vector<BatchUnpack> pending_unpack;
    int pending_unpack_max = 0;
    struct Reg { const void *p; size_t bytes; bool ours; };
    std::vector<Reg> regs;
    std::vector<PinnedBuf<float>> stage;     // [B*5] lazily allocated staging (pin_host_clouds == 0: maps, odometry clouds)
    // host sweeps with pin_host_clouds == 0: packed back to back into ONE pinned block per step (two alternate) and sent
    // with ONE copy on its own stream - ~100 cudaMemcpyAsync calls per step cost 2.4 ms of host time (bench.py e2e)
    PinnedBuf<unsigned char> up_pin[2]; DevBuf<unsigned char> up_dev[2];
    size_t up_cap = 0, up_off = 0, up_sent = 0; int up_cur = 0;
    cudaEvent_t up_ev[2] = {}; bool up_busy[2] = {};
    cudaStream_t up_stream = nullptr;
    std::vector<cudaEvent_t> stage_ev; std::vector<char> stage_busy;

    // device-resident key-frame stores (llb_batch_enable_keyframes): per slot an arena of key-frame clouds, the assembled
    // raw local map, the voxel scratch of its two map filters and the DS maps they produce
    bool kf_enabled = false;
    int cap_raw = 0, max_kf = 0;
    std::vector<KeyFrameStore> kfs;
    std::vector<VoxelFilter> vox;            // 2B (corner, surf)
    DevBuf<float4> raw_map;                  // [B][2][cap_raw]
    DevBuf<float4> ds_map;                   // [B][2][cap_raw]
    DevBuf<int> ds_map_n;                    // [B][2]
    DevBuf<AsmSeg> seg_dev;
    PinnedBuf<AsmSeg> seg_pin[RING];
    struct AsmReq {
        std::vector<int> ids; std::vector<float> poses; bool pending = false; int rc = 0, rs = 0;
        // the segment table of the last assembly: reused while the same key-frames are asked for with the same poses (the
        // sin / cos of 6 angles per key-frame cloud are most of the host cost of a step with 100 key-frames per slot)
        std::vector<int> c_ids; std::vector<float> c_poses; std::vector<AsmSeg> c_seg; size_t c_oc = 0, c_os = 0;
    };
    std::vector<AsmReq> asm_req;
    std::vector<char> map_from_kf;           // the slot's current map is the assembled one (counts live on the device)
    std::vector<BatchCopy> pending_copy;
    int pending_copy_max = 0;
    bool have_results = false;

    // featureAssociation of the slots (llb_batch_odom_*): one persistent CTA per slot in one launch
    bool od_ready = false;
    OdomParams oprm{};
    DevBuf<float4> od_clouds;                // [B][4][cap_scan]: cornerLast, surfLast, cornerPointsSharp, surfPointsFlat
    DevBuf<float> od_raw;                    // [B][4][cap_scan * 8]
    DevBuf<float> od_ind;                    // [B][5 * cap_scan]
    DevBuf<OdomState> od_states;
    PinnedBuf<OdomState> od_pin_states;
    DevBuf<OdomBatchJob> od_jobs_dev; PinnedBuf<OdomBatchJob> od_jobs_pin;
    DevBuf<BatchUnpack> od_unp_dev; PinnedBuf<BatchUnpack> od_unp_pin;
    DevBuf<float> od_poses_dev; PinnedBuf<float> od_poses_pin;
    std::vector<int> od_n;                   // [B][4] lengths; -1 = not set
    std::vector<BatchUnpack> od_pending; int od_pending_max = 0;

    struct GraphEntry { StepSig sig; cudaGraphExec_t exec; long long launches; };
    std::vector<GraphEntry> graphs;          // captured steps, one per launch geometry
    bool use_graph = true;
    cudaStream_t stream2 = nullptr;          // forked stream of the step (downsampleCurrentScan beside the map side)
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool pending = false;
    bool profile = false;
    cudaEvent_t pev[64] = {};
    int n_pev = 0;
    int pev_kind[64] = {};
    float prof_ms[PROF_N] = {};
};

namespace {

template <typename F>
int guarded(llb_batch *b, F &&f)
{
    if (!b) return LLB_ERR_INVALID;
    try {
        cudaError_t e = cudaSetDevice(b->device);
        if (e != cudaSuccess) { b->err = cudaGetErrorString(e); return LLB_ERR_CUDA; }
        return f();
    } catch (const CudaError &e) {
        b->err = e.what();
        return LLB_ERR_CUDA;
    } catch (const std::exception &e) {
        b->err = e.what();
        return LLB_ERR_INVALID;
    }
}

bool ensure_registered(llb_batch *c, const void *p, size_t bytes)
{
    for (auto &r : c->regs)
        if (r.p == p && r.bytes >= bytes) return true;
    for (size_t i = 0; i < c->regs.size(); i++)
        if (c->regs[i].p == p) {
            if (c->regs[i].ours) { cudaStreamSynchronize(c->stream); cudaHostUnregister(const_cast<void *>(p)); }
            c->regs.erase(c->regs.begin() + i);
            break;
        }
    if (c->regs.size() >= 1024) {
        if (c->regs[0].ours) { cudaStreamSynchronize(c->stream); cudaHostUnregister(const_cast<void *>(c->regs[0].p)); }
        c->regs.erase(c->regs.begin());
    }
    cudaError_t e = cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); c->regs.push_back({ p, bytes, false }); return true; }
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    c->regs.push_back({ p, bytes, true });
    return true;
}

// host cloud (32 B stride) -> raw device slice (async DMA) + a pending unpack job into dst
void upload(llb_batch *c, int stage_id, const llb_point *src, int n, float *raw_dev, float4 *dst,
            std::vector<BatchUnpack> *list = nullptr, int *list_max = nullptr)
{
    if (n <= 0) return;
    const size_t bytes = (size_t)n * sizeof(llb_point);
    if (c->prm.pin_host_clouds && ensure_registered(c, src, bytes)) {
        LLB_CUDA(cudaMemcpyAsync(raw_dev, src, bytes, cudaMemcpyHostToDevice, c->stream));
    } else {
        if (c->stage_busy[stage_id]) { LLB_CUDA(cudaEventSynchronize(c->stage_ev[stage_id])); c->stage_busy[stage_id] = 0; }
        c->stage[stage_id].ensure((size_t)n * 8);
        std::memcpy(c->stage[stage_id].p, src, bytes);
        LLB_CUDA(cudaMemcpyAsync(raw_dev, c->stage[stage_id].p, bytes, cudaMemcpyHostToDevice, c->stream));
        LLB_CUDA(cudaEventRecord(c->stage_ev[stage_id], c->stream));
        c->stage_busy[stage_id] = 1;
    }
    if (list) { list->push_back(BatchUnpack{ raw_dev, dst, n }); *list_max = std::max(*list_max, n); return; }
    c->pending_unpack.push_back(BatchUnpack{ raw_dev, dst, n });
    c->pending_unpack_max = std::max(c->pending_unpack_max, n);
}

// packed staging of a host sweep cloud (pin_host_clouds == 0); false = no room, take the per-cloud path
bool upload_packed(llb_batch *c, const llb_point *src, int n, float4 *dst)
{
    if (n <= 0) return true;
    if (!c->up_stream) {
        c->up_cap = (size_t)c->B * 3 * c->cap_scan * sizeof(llb_point) + 256 * 3 * (size_t)c->B;
        for (int k = 0; k < 2; k++) {
            c->up_pin[k].ensure(c->up_cap); c->up_dev[k].ensure(c->up_cap);
            LLB_CUDA(cudaEventCreateWithFlags(&c->up_ev[k], cudaEventDisableTiming));
        }
        LLB_CUDA(cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
    }
    const size_t bytes = (size_t)n * sizeof(llb_point);
    if (c->up_off + bytes > c->up_cap) return false;
    if (c->up_off == 0 && c->up_busy[c->up_cur]) {           // the copy that last read this pinned block (two steps ago)
        LLB_CUDA(cudaEventSynchronize(c->up_ev[c->up_cur])); c->up_busy[c->up_cur] = false;
    }
    std::memcpy(c->up_pin[c->up_cur].p + c->up_off, src, bytes);
    c->pending_unpack.push_back(BatchUnpack{ reinterpret_cast<const float *>(c->up_dev[c->up_cur].p + c->up_off), dst, n });
    c->pending_unpack_max = std::max(c->pending_unpack_max, n);
    c->up_off = (c->up_off + bytes + 255) & ~(size_t)255;
    return true;
}

// sends what has been packed since the last call (one copy, own stream: it overlaps the kernels of a step in flight)
void upload_flush(llb_batch *c)
{
    if (!c->up_stream || c->up_off <= c->up_sent) return;
    const int k = c->up_cur;
    LLB_CUDA(cudaMemcpyAsync(c->up_dev[k].p + c->up_sent, c->up_pin[k].p + c->up_sent, c->up_off - c->up_sent,
                             cudaMemcpyHostToDevice, c->up_stream));
    LLB_CUDA(cudaEventRecord(c->up_ev[k], c->up_stream));
    c->up_busy[k] = true;
    c->up_sent = c->up_off;
}

void prof_mark(llb_batch *c, int kind)
{
    if (!c->profile || c->n_pev >= 64) return;
    LLB_CUDA(cudaEventRecord(c->pev[c->n_pev], c->stream));
    c->pev_kind[c->n_pev] = kind;
    c->n_pev++;
}

int enqueue_step(llb_batch *c, const float *T)
{
    const int B = c->B;
    for (int s = 0; s < B; s++)
        if (!c->slots[s].scan_set || !(c->slots[s].map_set || (c->kf_enabled && c->asm_req[s].pending))) return LLB_ERR_STATE;
    // capacity of the step's job tables, checked BEFORE anything is consumed: a refused step leaves the batch as it was
    if ((int)c->pending_unpack.size() > 5 * B || (int)c->pending_copy.size() > 3 * B) return LLB_ERR_STATE;
    // ---- build this step's tables in the next pinned block
    const int rp = c->ring_pos;
    c->ring_pos = (rp + 1) % RING;
    if (c->step_busy[rp]) { LLB_CUDA(cudaEventSynchronize(c->step_ev[rp])); c->step_busy[rp] = false; }
    unsigned char *hp = c->step_pin[rp].p, *dp = c->step_dev.p;
    const StepLayout &L = c->lay;
    int *h_scan_n = (int *)(hp + L.off_scan_n), *h_map_n = (int *)(hp + L.off_map_n);
    float *h_poses = (float *)(hp + L.off_poses);
    SmallJob *h_s1 = (SmallJob *)(hp + L.off_small1), *h_s2 = (SmallJob *)(hp + L.off_small2);
    GridJob *h_grid = (GridJob *)(hp + L.off_grid);
    BatchReg *h_regs = (BatchReg *)(hp + L.off_regs);
    BatchUnpack *h_unp = (BatchUnpack *)(hp + L.off_unpack);
    const int *d_scan_n = (const int *)(dp + L.off_scan_n), *d_map_n = (const int *)(dp + L.off_map_n);
    const float leaf[3] = { c->prm.corner_leaf, c->prm.surf_leaf, c->prm.outlier_leaf };
    int ngrid = 0, map_n_max = 1;
    int vmax1 = 1, vmax2 = 1;
    LargeVoxelJob *h_vox = (LargeVoxelJob *)(hp + L.off_vox);
    BatchCopy *h_copy = (BatchCopy *)(hp + L.off_copy);
    int nvox = 0, nseg = 0, seg_max = 1, raw_max = 1;
    AsmSeg *h_seg = c->kf_enabled ? c->seg_pin[rp].p : nullptr;
    for (int s = 0; s < B; s++) {
        llb_batch::Slot &sl = c->slots[s];
        float4 *ds = c->scan_ds.p + (size_t)s * 5 * c->cap_scan;
        float4 *ds_out[4] = { ds, ds + c->cap_scan, ds + 2 * (size_t)c->cap_scan, ds + 3 * (size_t)c->cap_scan };
        int *dsn = c->ds_n.p + 4 * s;
        for (int k = 0; k < 3; k++) {
            h_scan_n[3 * s + k] = sl.scan_n[k];
            SmallJob &j = h_s1[3 * s + k];                   // MO:1069-1082
            j.in.a = sl.scan[k]; j.in.na_dev = d_scan_n + 3 * s + k; j.in.na = sl.scan_n[k];
            j.in.b = nullptr; j.in.nb_dev = nullptr; j.in.nb = 0;
            j.leaf = leaf[k]; j.out = ds_out[k]; j.n_out = dsn + k;
        }
        vmax1 = std::max(vmax1, std::max(sl.scan_n[0], std::max(sl.scan_n[1], sl.scan_n[2])));
        vmax2 = std::max(vmax2, sl.scan_n[1] + sl.scan_n[2]);
        SmallJob &t = h_s2[s];                               // MO:1084-1090 (C12): surfDS + outlierDS -> DS again
        t.in.a = ds_out[1]; t.in.na_dev = dsn + 1; t.in.na = sl.scan_n[1];
        t.in.b = ds_out[2]; t.in.nb_dev = dsn + 2; t.in.nb = sl.scan_n[2];
        t.leaf = c->prm.surf_leaf; t.out = ds_out[3]; t.n_out = dsn + 3;
        if (c->kf_enabled && c->asm_req[s].pending) {
            // cloud part of extractSurroundingKeyFrames for this slot: segments of the fused transform + concatenation
            // launch, then its two map voxel filters (MO:1057-1064) as jobs of the batched multi-kernel path
            llb_batch::AsmReq &rq = c->asm_req[s];
            float4 *rawc = c->raw_map.p + (size_t)(2 * s) * c->cap_raw, *raws = rawc + c->cap_raw;
            float4 *dsc = c->ds_map.p + (size_t)(2 * s) * c->cap_raw, *dss = dsc + c->cap_raw;
            size_t oc = 0, os = 0;
            if (rq.c_ids == rq.ids && rq.c_poses == rq.poses && !rq.c_seg.empty()) {
                for (const AsmSeg &sg : rq.c_seg) { h_seg[nseg++] = sg; seg_max = std::max(seg_max, sg.n); }
                oc = rq.c_oc; os = rq.c_os;
            } else {
                rq.c_seg.clear();
                auto put = [&](const AsmSeg &sg) { h_seg[nseg++] = sg; seg_max = std::max(seg_max, sg.n); rq.c_seg.push_back(sg); };
                for (size_t k = 0; k < rq.ids.size(); k++) {
                    const KeyFrameRec &kr = c->kfs[s].rec(rq.ids[k]);
                    const float *p = rq.poses.data() + 6 * k;
                    AsmSeg sg{};
                    sg.ctRoll = cosf(p[0]); sg.stRoll = sinf(p[0]); sg.ctPitch = cosf(p[1]); sg.stPitch = sinf(p[1]);
                    sg.ctYaw = cosf(p[2]); sg.stYaw = sinf(p[2]); sg.tx = p[3]; sg.ty = p[4]; sg.tz = p[5];
                    sg.src = kr.cloud[0]; sg.n = kr.n[0]; sg.dst = rawc + oc; oc += kr.n[0];
                    if (sg.n > 0) put(sg);
                    sg.src = kr.cloud[1]; sg.n = kr.n[1]; sg.dst = raws + os; os += kr.n[1];
                    if (sg.n > 0) put(sg);
                    sg.src = kr.cloud[2]; sg.n = kr.n[2]; sg.dst = raws + os; os += kr.n[2];
                    if (sg.n > 0) put(sg);
                }
                rq.c_ids = rq.ids; rq.c_poses = rq.poses; rq.c_oc = oc; rq.c_os = os;
            }
            VoxelInput vc; vc.a = rawc; vc.na = (int)oc;
            VoxelInput vs; vs.a = raws; vs.na = (int)os;
            int *dn = c->ds_map_n.p + 2 * s;
            h_vox[nvox++] = c->vox[2 * s].large_job(vc, c->prm.corner_leaf, dsc, dn);
            h_vox[nvox++] = c->vox[2 * s + 1].large_job(vs, c->prm.surf_leaf, dss, dn + 1);
            raw_max = std::max(raw_max, (int)std::max(oc, os));
            sl.map[0] = dsc; sl.map[1] = dss; sl.map_n[0] = (int)oc; sl.map_n[1] = (int)os;   // upper bounds
            sl.map_set = true; sl.map_dirty = true;
            c->map_from_kf[s] = 1;
            rq.pending = false;
        }
        for (int k = 0; k < 2; k++) {
            h_map_n[2 * s + k] = sl.map_n[k];
            if (sl.map_dirty) {
                const int *n_dev = (c->kf_enabled && c->map_from_kf[s]) ? c->ds_map_n.p + 2 * s + k : d_map_n + 2 * s + k;
                h_grid[ngrid++] = c->grids[2 * s + k].job(sl.map[k], n_dev, sl.map_n[k]);
                map_n_max = std::max(map_n_max, sl.map_n[k]);
            }
        }
        sl.map_dirty = false;
        BatchReg &r = h_regs[s];
        r.corner = ds_out[0]; r.surf = ds_out[3]; r.nc_dev = dsn + 0; r.ns_dev = dsn + 3;
        r.cmap = c->grids[2 * s].view(); r.smap = c->grids[2 * s + 1].view();
        r.st = c->states.p + s;
        r.qperm = c->qperm.p + (size_t)s * c->qcap;
        r.qprev = c->qprev.p + (size_t)s * c->qcap;
        r.partials = c->partials.p + (size_t)s * c->part_stride * S2M_ACC;
        r.cap = c->qcap;
        for (int i = 0; i < 6; i++) h_poses[6 * s + i] = T[6 * s + i];
    }
    // shared-memory capacity of the two voxel launches: this step's largest filter, rounded to 1024 points
    c->vox_cap1 = (vmax1 + 1023) & ~1023; c->vox_cap2 = (vmax2 + 1023) & ~1023;
    const int nunp = (int)c->pending_unpack.size();
    for (int i = 0; i < nunp; i++) h_unp[i] = c->pending_unpack[i];
    // launch-geometry upper bounds are rounded up so that steps of similar size share one captured graph
    auto round_up = [](int v, int g) { return (v + g - 1) / g * g; };
    const int unp_max = round_up(c->pending_unpack_max, 2048);
    c->pending_unpack.clear(); c->pending_unpack_max = 0;
    const int ncopy = (int)c->pending_copy.size();
    for (int i = 0; i < ncopy; i++) h_copy[i] = c->pending_copy[i];
    const int copy_max = round_up(c->pending_copy_max, 2048);
    c->pending_copy.clear(); c->pending_copy_max = 0;
    seg_max = round_up(seg_max, 2048);
    map_n_max = round_up(map_n_max, 4096);
    raw_max = std::min(round_up(raw_max, 16384), std::max(c->cap_raw, raw_max));

    // ---- enqueue
    c->n_pev = 0;
    if (c->up_stream && c->up_off > 0) {                     // packed host sweeps of this step: sent (if not yet), then awaited
        upload_flush(c);
        LLB_CUDA(cudaStreamWaitEvent(c->stream, c->up_ev[c->up_cur], 0));
        c->up_cur ^= 1; c->up_off = 0; c->up_sent = 0;
    }
    LLB_CUDA(cudaEventRecord(c->ev0, c->stream));
    prof_mark(c, -1);
    LLB_CUDA(cudaMemcpyAsync(dp, hp, L.bytes, cudaMemcpyHostToDevice, c->stream));
    if (nseg > 0)
        LLB_CUDA(cudaMemcpyAsync(c->seg_dev.p, h_seg, sizeof(AsmSeg) * nseg, cudaMemcpyHostToDevice, c->stream));
    LLB_CUDA(cudaEventRecord(c->step_ev[rp], c->stream));    // after the LAST copy that reads ring slot rp (step block + segments)
    c->step_busy[rp] = true;
    // every kernel launch of the step; all arguments are device-resident tables or the counts of `sig`
    auto enqueue_kernels = [&]() -> long long {
        long long nl = 0;
        if (nunp > 0) {
            launch_batch_unpack((const BatchUnpack *)(dp + L.off_unpack), nunp, unp_max, c->stream);
            nl++;
        }
        if (ncopy > 0) {                                         // key-frames saved since the last step: DS clouds -> arenas,
            launch_batch_copy((const BatchCopy *)(dp + L.off_copy), ncopy, copy_max, c->stream);   // before they are overwritten
            nl++;
        }
        // downsampleCurrentScan only depends on the sweeps: it runs on a forked stream beside the map side of the step
        // (key-frame assembly, map voxel filters, index build) and is joined before the registrations start
        LLB_CUDA(cudaEventRecord(c->fork_ev, c->stream));
        LLB_CUDA(cudaStreamWaitEvent(c->stream2, c->fork_ev, 0));
        launch_voxel_cta_jobs((const SmallJob *)(dp + L.off_small1), 3 * B, c->vox_cap1, c->stream2);
        launch_voxel_cta_jobs((const SmallJob *)(dp + L.off_small2), B, c->vox_cap2, c->stream2);
        nl += 2;
        LLB_CUDA(cudaEventRecord(c->join_ev, c->stream2));
        if (nseg > 0) {
            launch_kf_assemble(c->seg_dev.p, nseg, seg_max, c->stream);
            nl++;
        }
        if (nvox > 0)
            nl += VoxelFilter::launch_large((const LargeVoxelJob *)(dp + L.off_vox), nvox,
                                                     raw_max, c->stream);      // scratch is sized for cap_raw >= raw_max
        prof_mark(c, 0);
        if (ngrid > 0)
            nl += GridIndex::build_table((const GridJob *)(dp + L.off_grid), ngrid, map_n_max,
                                                  std::sqrt(c->prm.knn_max_sqdist), c->grids[0].max_cells(), c->grid_ctas, c->stream);
        prof_mark(c, 2);
        LLB_CUDA(cudaStreamWaitEvent(c->stream, c->join_ev, 0));
        prof_mark(c, 1);                                         // what is left of downsampleCurrentScan after the overlap
        const BatchReg *regs = (const BatchReg *)(dp + L.off_regs);
        launch_batch_prepare(regs, (const float *)(dp + L.off_poses), B, c->sprm, c->prm.s2m_max_iterations, c->queue.p, c->stream);
        nl++;
        launch_batch_qsort(regs, B, std::max(c->vox_cap1, c->vox_cap2), c->stream);   // queries ordered by search cost
        nl++;
        prof_mark(c, 5);
        // scan2MapOptimization of every slot: ONE persistent launch for all LM iterations (batch.cu)
        launch_batch_lm(regs, B, c->lm_grid, c->sprm, c->prm.s2m_max_iterations, c->queue.p, c->stream);
        nl++;
        prof_mark(c, 4);
        launch_batch_collect(regs, B, c->results.p, c->queue.p, c->stream);
        nl++;
        prof_mark(c, 5);
        return nl;
    };
    // The ~30 dependent launches of a step are replayed as ONE CUDA graph (forked stream included): a step in steady
    // state has the same launch geometry every time - its tables live in the device block refreshed above - so the
    // graph is captured once per geometry and the host pays one launch instead of thirty (LLB_BATCH_GRAPH=0 disables).
    const StepSig sig{ nunp, unp_max, ncopy, copy_max, nseg, seg_max, nvox, raw_max, ngrid, map_n_max, c->vox_cap1, c->vox_cap2 };
    if (c->use_graph && !c->profile) {
        llb_batch::GraphEntry *ge = nullptr;
        for (auto &g : c->graphs) if (std::memcmp(&g.sig, &sig, sizeof(sig)) == 0) { ge = &g; break; }
        if (!ge) {
            if (c->graphs.size() >= 16) { cudaGraphExecDestroy(c->graphs[0].exec); c->graphs.erase(c->graphs.begin()); }
            cudaGraph_t graph = nullptr;
            LLB_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
            long long nl = 0;
            try { nl = enqueue_kernels(); } catch (...) { cudaStreamEndCapture(c->stream, &graph); if (graph) cudaGraphDestroy(graph); throw; }
            LLB_CUDA(cudaStreamEndCapture(c->stream, &graph));
            llb_batch::GraphEntry e{};
            e.sig = sig; e.launches = nl;
            LLB_CUDA(cudaGraphInstantiate(&e.exec, graph, 0));
            cudaGraphDestroy(graph);
            c->graphs.push_back(e);
            ge = &c->graphs.back();
        }
        LLB_CUDA(cudaGraphLaunch(ge->exec, c->stream));
        c->launches += ge->launches;
    } else {
        c->launches += enqueue_kernels();
    }
    LLB_CUDA(cudaEventRecord(c->ev1, c->stream));
    LLB_CUDA(cudaMemcpyAsync(c->pin_results.p, c->results.p, sizeof(BatchResult) * B, cudaMemcpyDeviceToHost, c->stream));
    c->pending = true;
    return LLB_OK;
}

int fetch_result(llb_batch *c, float *T, llb_stats *stats)
{
    if (!c->pending) return LLB_ERR_STATE;
    LLB_CUDA(cudaStreamSynchronize(c->stream));
    c->pending = false;
    c->have_results = true;
    float ms = 0.f;
    LLB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    for (int s = 0; s < c->B; s++) {
        const BatchResult &r = c->pin_results.p[s];
        if (T && !r.skipped) for (int i = 0; i < 6; i++) T[6 * s + i] = r.T[i];
        if (stats) {
            llb_stats &st = stats[s];
            st.iterations = r.iters; st.converged = r.converged; st.n_correspondences = r.n_corr;
            st.is_degenerate = r.is_degenerate; st.skipped = r.skipped; st.n_corner_ds = r.nc; st.n_surf_ds = r.ns;
            st.device_ms = ms;
        }
    }
    if (c->profile) {
        for (int k = 0; k < PROF_N; k++) c->prof_ms[k] = 0.f;
        for (int i = 1; i < c->n_pev; i++) {
            float d = 0.f;
            LLB_CUDA(cudaEventElapsedTime(&d, c->pev[i - 1], c->pev[i]));
            if (c->pev_kind[i] >= 0 && c->pev_kind[i] < PROF_N) c->prof_ms[c->pev_kind[i]] += d;
        }
    }
    return LLB_OK;
}

}  // namespace

extern "C" {

int llb_batch_create(const llb_params *p, int device, int n_slots, int max_scan_points, int max_map_points, llb_batch **out)
{
    if (!out) return LLB_ERR_INVALID;
    *out = nullptr;
    if (n_slots < 1 || n_slots > 4096 || max_scan_points < 1 || max_map_points < 1) return LLB_ERR_INVALID;
    // every filter of downsampleCurrentScan must fit the one-CTA voxel kernel (shared-memory sort of <= 16384 points);
    // surf + outlier are concatenated for the fourth filter: checked per sweep in llb_batch_scan_set
    if (max_scan_points > VoxelFilter::SMALL_MAX) return LLB_ERR_CAPACITY;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return LLB_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) return LLB_ERR_NO_DEVICE;
    llb_batch *c = new llb_batch();
    c->device = device;
    if (p) c->prm = *p; else llb_params_default(&c->prm);
    int rc = guarded(c, [&]() {
        const int B = n_slots;
        c->B = B; c->cap_scan = max_scan_points; c->cap_map = max_map_points;
        c->qcap = 3 * max_scan_points;
        c->lay = StepLayout(B);
        S2mParams &q = c->sprm;
        q.knn_max_sqdist = c->prm.knn_max_sqdist; q.min_corr = c->prm.s2m_min_correspondences;
        q.degeneracy_thresh = c->prm.s2m_degeneracy_thresh; q.converge_deg = c->prm.s2m_converge_deg;
        q.converge_cm = c->prm.s2m_converge_cm; q.corner_map_min = c->prm.corner_map_min;
        q.surf_map_min = c->prm.surf_map_min; q.max_ctas = 0;
        LLB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        LLB_CUDA(cudaEventCreate(&c->ev0)); LLB_CUDA(cudaEventCreate(&c->ev1));
        LLB_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
        c->use_graph = !(getenv("LLB_BATCH_GRAPH") && atoi(getenv("LLB_BATCH_GRAPH")) == 0);
        LLB_CUDA(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
        LLB_CUDA(cudaEventCreateWithFlags(&c->join_ev, cudaEventDisableTiming));
        for (int i = 0; i < 64; i++) LLB_CUDA(cudaEventCreate(&c->pev[i]));
        for (int i = 0; i < RING; i++) {
            LLB_CUDA(cudaEventCreateWithFlags(&c->step_ev[i], cudaEventDisableTiming));
            c->step_pin[i].ensure(c->lay.bytes);
        }
        c->step_dev.ensure(c->lay.bytes);
        c->scan_in.ensure((size_t)B * 3 * c->cap_scan);
        c->scan_raw.ensure((size_t)B * 3 * c->cap_scan * 8);
        c->scan_ds.ensure((size_t)B * 5 * c->cap_scan);
        c->map_in.ensure((size_t)B * 2 * c->cap_map);
        c->map_raw.ensure((size_t)B * 2 * c->cap_map * 8);
        c->ds_n.ensure((size_t)B * 4);
        LLB_CUDA(cudaMemset(c->ds_n.p, 0, sizeof(int) * 4 * B));
        c->states.ensure(B);
        c->qperm.ensure((size_t)B * c->qcap); c->qprev.ensure((size_t)B * c->qcap);
        // launch geometry: enough CTAs to fill 148 SMs several times over, independent of B
        c->grid_ctas = std::max(2, std::min(148 * 4, 148 * 8 / (2 * B)));
        c->part_stride = c->qcap / 32 + 2;                               // one partial per chunk of 32 queries
        c->partials.ensure((size_t)B * c->part_stride * S2M_ACC);
        c->queue.ensure(1);
        c->slot_info.ensure(B); c->ctl.ensure(B);
        LLB_CUDA(cudaMemset(c->slot_info.p, 0, sizeof(BatchSlotInfo) * B));
        { std::vector<unsigned> done((size_t)B, 0xffff8000u); LLB_CUDA(cudaMemcpy(c->ctl.p, done.data(), sizeof(unsigned) * B, cudaMemcpyHostToDevice)); }
        c->lm_grid = batch_lm_grid();
        BatchQueue bq{};
        bq.slot = c->slot_info.p; bq.ctl = c->ctl.p;
        if (getenv("LLB_ITER_PROF")) {                           // debug: per-warp cycles by phase of the registration kernel
            const size_t n = (size_t)c->lm_grid * (BATCH_ITER_THREADS / 32) * 12;
            c->iter_prof.ensure(n);
            LLB_CUDA(cudaMemset(c->iter_prof.p, 0, n * sizeof(long long)));
            bq.prof = c->iter_prof.p;
        }
        LLB_CUDA(cudaMemcpy(c->queue.p, &bq, sizeof(bq), cudaMemcpyHostToDevice));
        c->results.ensure(B); c->pin_results.ensure(B);
        const int cells = std::min(c->prm.max_grid_cells, 1 << 22);
        c->grids.resize(2 * (size_t)B);
        for (auto &g : c->grids) { g.init(cells); g.job(nullptr, nullptr, c->cap_map); }
        c->slots.resize(B);
        c->stage.resize((size_t)B * 9); c->stage_ev.assign((size_t)B * 9, nullptr); c->stage_busy.assign((size_t)B * 9, 0);
        for (auto &e : c->stage_ev) LLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        launch_batch_state_init(c->states.p, B, c->stream);
        LLB_CUDA(cudaDeviceSynchronize());
        return (int)LLB_OK;
    });
    if (rc != LLB_OK) { llb_batch_destroy(c); return rc; }
    *out = c;
    return LLB_OK;
}

int llb_batch_destroy(llb_batch *c)
{
    if (!c) return LLB_ERR_INVALID;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->iter_prof.p) {                                        // debug: phase durations (cycles) of the last profiled launch
        const int nw = c->lm_grid * (BATCH_ITER_THREADS / 32);
        std::vector<long long> h((size_t)nw * 12);
        cudaMemcpy(h.data(), c->iter_prof.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        static const char *names[10] = { "fetch item + slot record", "query + transform + bound", "A own row", "B neighbour index",
                                         "C rank runs", "gather 5", "fit + Jacobian", "products + partial", "fence + ticket", "LM step" };
        std::vector<long long> col(nw);
        auto stat = [&](int k, const char *name) {
            int n = 0;
            for (int w = 0; w < nw; w++) if (h[(size_t)w * 12 + 11] > 0 || k >= 10) col[n++] = h[(size_t)w * 12 + k];
            if (n == 0) return;
            std::sort(col.begin(), col.begin() + n);
            double sum = 0; for (int i = 0; i < n; i++) sum += (double)col[i];
            fprintf(stderr, "  %-28s mean %9.0f  p50 %9lld  p90 %9lld  max %9lld  (%d warps)\n", name, sum / n, col[n / 2], col[n * 9 / 10], col[n - 1], n);
        };
        fprintf(stderr, "LLB_ITER_PROF: cycles per warp summed over its items, by phase\n");
        for (int k = 0; k < 10; k++) stat(k, names[k]);
        stat(10, "warp lifetime"); stat(11, "items per warp");
        c->iter_prof.release();
    }
    for (auto &r : c->regs) if (r.ours) cudaHostUnregister(const_cast<void *>(r.p));
    for (int k = 0; k < 2; k++) { c->up_pin[k].release(); c->up_dev[k].release(); if (c->up_ev[k]) cudaEventDestroy(c->up_ev[k]); }
    if (c->up_stream) cudaStreamDestroy(c->up_stream);
    c->features.release();
    c->scan_in.release(); c->scan_raw.release(); c->scan_ds.release(); c->map_in.release(); c->map_raw.release();
    c->ds_n.release(); c->states.release(); c->qperm.release(); c->partials.release(); c->qprev.release(); c->queue.release(); c->slot_info.release(); c->ctl.release();
    c->results.release(); c->pin_results.release(); c->step_dev.release();
    for (int i = 0; i < RING; i++) { c->step_pin[i].release(); if (c->step_ev[i]) cudaEventDestroy(c->step_ev[i]); }
    for (auto &g : c->grids) g.release();
    c->od_clouds.release(); c->od_raw.release(); c->od_ind.release(); c->od_states.release(); c->od_pin_states.release();
    c->od_jobs_dev.release(); c->od_jobs_pin.release(); c->od_unp_dev.release(); c->od_unp_pin.release();
    c->od_poses_dev.release(); c->od_poses_pin.release();
    for (auto &k : c->kfs) k.release();
    for (auto &v : c->vox) v.release();
    c->raw_map.release(); c->ds_map.release(); c->ds_map_n.release(); c->seg_dev.release();
    for (int i = 0; i < RING; i++) c->seg_pin[i].release();
    for (auto &s : c->stage) s.release();
    for (auto &e : c->stage_ev) if (e) cudaEventDestroy(e);
    for (int i = 0; i < 64; i++) if (c->pev[i]) cudaEventDestroy(c->pev[i]);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (auto &g : c->graphs) cudaGraphExecDestroy(g.exec);
    c->graphs.clear();
    if (c->fork_ev) cudaEventDestroy(c->fork_ev);
    if (c->join_ev) cudaEventDestroy(c->join_ev);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return LLB_OK;
}

const char *llb_batch_last_error(const llb_batch *c) { return c ? c->err.c_str() : "null batch"; }
void *llb_batch_stream(llb_batch *c) { return c ? (void *)c->stream : nullptr; }
long long llb_batch_launch_count(const llb_batch *c) { return c ? c->launches : 0; }
int llb_batch_slots(const llb_batch *c) { return c ? c->B : 0; }

int llb_batch_scan_set(llb_batch *c, int slot, const llb_point *corner, int nc, const llb_point *surf, int ns,
                       const llb_point *outlier, int no)
{
    return guarded(c, [&]() {
        if (slot < 0 || slot >= c->B || nc < 0 || ns < 0 || no < 0 || (nc > 0 && !corner) || (ns > 0 && !surf) ||
            (no > 0 && !outlier)) return (int)LLB_ERR_INVALID;
        if (nc > c->cap_scan || ns > c->cap_scan || no > c->cap_scan || ns + no > VoxelFilter::SMALL_MAX)
            return (int)LLB_ERR_CAPACITY;
        const llb_point *src[3] = { corner, surf, outlier };
        const int n[3] = { nc, ns, no };
        llb_batch::Slot &sl = c->slots[slot];
        for (int k = 0; k < 3; k++) {
            float4 *dst = c->scan_in.p + ((size_t)slot * 3 + k) * c->cap_scan;
            if (c->prm.pin_host_clouds || !upload_packed(c, src[k], n[k], dst))
                upload(c, slot * 5 + k, src[k], n[k], c->scan_raw.p + ((size_t)slot * 3 + k) * c->cap_scan * 8, dst);
            sl.scan[k] = dst; sl.scan_n[k] = n[k];
        }
        sl.scan_set = true;
        return (int)LLB_OK;
    });
}

int llb_batch_scan_set_dev(llb_batch *c, int slot, const void *corner, int nc, const void *surf, int ns,
                           const void *outlier, int no)
{
    return guarded(c, [&]() {
        if (slot < 0 || slot >= c->B || nc < 0 || ns < 0 || no < 0) return (int)LLB_ERR_INVALID;
        if (nc > c->cap_scan || ns > c->cap_scan || no > c->cap_scan || ns + no > VoxelFilter::SMALL_MAX)
            return (int)LLB_ERR_CAPACITY;
        llb_batch::Slot &sl = c->slots[slot];
        sl.scan[0] = (const float4 *)corner; sl.scan[1] = (const float4 *)surf; sl.scan[2] = (const float4 *)outlier;
        sl.scan_n[0] = nc; sl.scan_n[1] = ns; sl.scan_n[2] = no;
        sl.scan_set = true;
        return (int)LLB_OK;
    });
}

int llb_batch_map_set_ds(llb_batch *c, int slot, const llb_point *corner, int mc, const llb_point *surf, int ms)
{
    return guarded(c, [&]() {
        if (slot < 0 || slot >= c->B || mc < 0 || ms < 0 || (mc > 0 && !corner) || (ms > 0 && !surf)) return (int)LLB_ERR_INVALID;
        if (mc > c->cap_map || ms > c->cap_map) return (int)LLB_ERR_CAPACITY;
        const llb_point *src[2] = { corner, surf };
        const int n[2] = { mc, ms };
        llb_batch::Slot &sl = c->slots[slot];
        for (int k = 0; k < 2; k++) {
            float4 *dst = c->map_in.p + ((size_t)slot * 2 + k) * c->cap_map;
            upload(c, slot * 5 + 3 + k, src[k], n[k], c->map_raw.p + ((size_t)slot * 2 + k) * c->cap_map * 8, dst);
            sl.map[k] = dst; sl.map_n[k] = n[k];
        }
        sl.map_set = true; sl.map_dirty = true;
        if (c->kf_enabled) c->map_from_kf[slot] = 0;
        return (int)LLB_OK;
    });
}

int llb_batch_map_set_ds_dev(llb_batch *c, int slot, const void *corner, int mc, const void *surf, int ms)
{
    return guarded(c, [&]() {
        if (slot < 0 || slot >= c->B || mc < 0 || ms < 0) return (int)LLB_ERR_INVALID;
        if (mc > c->cap_map || ms > c->cap_map) return (int)LLB_ERR_CAPACITY;
        llb_batch::Slot &sl = c->slots[slot];
        sl.map[0] = (const float4 *)corner; sl.map[1] = (const float4 *)surf;
        sl.map_n[0] = mc; sl.map_n[1] = ms;
        sl.map_set = true; sl.map_dirty = true;
        if (c->kf_enabled) c->map_from_kf[slot] = 0;
        return (int)LLB_OK;
    });
}

// bulk setters: one call for all slots (arrays of n_slots pointers / counts); a NULL pointer array entry with a zero
// count is an empty cloud
int llb_batch_scan_set_all(llb_batch *c, const llb_point *const *corner, const int *nc, const llb_point *const *surf,
                           const int *ns, const llb_point *const *outlier, const int *no)
{
    if (!c || !corner || !nc || !surf || !ns || !outlier || !no) return LLB_ERR_INVALID;
    for (int s = 0; s < c->B; s++) {
        const int rc = llb_batch_scan_set(c, s, corner[s], nc[s], surf[s], ns[s], outlier[s], no[s]);
        if (rc != LLB_OK) return rc;
    }
    return guarded(c, [&]() { upload_flush(c); return (int)LLB_OK; });   // the sweeps start crossing PCIe now
}

int llb_batch_map_set_ds_all(llb_batch *c, const llb_point *const *corner_ds, const int *mc, const llb_point *const *surf_ds,
                             const int *ms)
{
    if (!c || !corner_ds || !mc || !surf_ds || !ms) return LLB_ERR_INVALID;
    for (int s = 0; s < c->B; s++) {
        const int rc = llb_batch_map_set_ds(c, s, corner_ds[s], mc[s], surf_ds[s], ms[s]);
        if (rc != LLB_OK) return rc;
    }
    return LLB_OK;
}

int llb_batch_scan_set_dev_all(llb_batch *c, const void *const *corner, const int *nc, const void *const *surf,
                               const int *ns, const void *const *outlier, const int *no)
{
    if (!c || !corner || !nc || !surf || !ns || !outlier || !no) return LLB_ERR_INVALID;
    for (int s = 0; s < c->B; s++) {
        const int rc = llb_batch_scan_set_dev(c, s, corner[s], nc[s], surf[s], ns[s], outlier[s], no[s]);
        if (rc != LLB_OK) return rc;
    }
    return LLB_OK;
}

int llb_batch_map_set_ds_dev_all(llb_batch *c, const void *const *corner_ds, const int *mc, const void *const *surf_ds,
                                 const int *ms)
{
    if (!c || !corner_ds || !mc || !surf_ds || !ms) return LLB_ERR_INVALID;
    for (int s = 0; s < c->B; s++) {
        const int rc = llb_batch_map_set_ds_dev(c, s, corner_ds[s], mc[s], surf_ds[s], ms[s]);
        if (rc != LLB_OK) return rc;
    }
    return LLB_OK;
}

int llb_batch_register_async(llb_batch *c, const float *T)
{
    return guarded(c, [&]() {
        if (!T) return (int)LLB_ERR_INVALID;
        if (c->pending) return (int)LLB_ERR_STATE;
        return enqueue_step(c, T);
    });
}

int llb_batch_result(llb_batch *c, float *T, llb_stats *stats)
{
    return guarded(c, [&]() { return fetch_result(c, T, stats); });
}

int llb_batch_register(llb_batch *c, float *T, llb_stats *stats)
{
    return guarded(c, [&]() {
        if (!T) return (int)LLB_ERR_INVALID;
        if (c->pending) return (int)LLB_ERR_STATE;
        int rc = enqueue_step(c, T);
        if (rc != LLB_OK) return rc;
        return fetch_result(c, T, stats);
    });
}

int llb_batch_scan_get_ds(llb_batch *c, int slot, int which, llb_point *out, int cap, int *n)
{
    return guarded(c, [&]() {
        if (!n || slot < 0 || slot >= c->B || which < 0 || which > 3) return (int)LLB_ERR_INVALID;
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        int cnt[4];
        LLB_CUDA(cudaMemcpy(cnt, c->ds_n.p + 4 * slot, sizeof(cnt), cudaMemcpyDeviceToHost));
        *n = cnt[which];
        if (!out) return (int)LLB_OK;
        if (cnt[which] > cap) return (int)LLB_ERR_CAPACITY;
        std::vector<float4> tmp(std::max(cnt[which], 1));
        const float4 *src = c->scan_ds.p + (size_t)slot * 5 * c->cap_scan + (size_t)which * c->cap_scan;
        if (cnt[which] > 0) LLB_CUDA(cudaMemcpy(tmp.data(), src, sizeof(float4) * cnt[which], cudaMemcpyDeviceToHost));
        for (int i = 0; i < cnt[which]; i++)
            out[i] = llb_point{ tmp[i].x, tmp[i].y, tmp[i].z, 1.0f, tmp[i].w, 0.f, 0.f, 0.f };
        return (int)LLB_OK;
    });
}

int llb_batch_get_degeneracy(llb_batch *c, int slot, int *deg)
{
    return guarded(c, [&]() {
        if (slot < 0 || slot >= c->B || !deg) return (int)LLB_ERR_INVALID;
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        LLB_CUDA(cudaMemcpy(deg, &c->states.p[slot].is_degenerate, sizeof(int), cudaMemcpyDeviceToHost));
        return (int)LLB_OK;
    });
}

// ------------------------------------------------------------------ key-frame stores of the slots

int llb_batch_enable_keyframes(llb_batch *c, int max_raw_map_points, int max_keyframes)
{
    return guarded(c, [&]() {
        if (max_raw_map_points < 1 || max_keyframes < 1) return (int)LLB_ERR_INVALID;
        if (c->kf_enabled) return (int)LLB_ERR_STATE;
        const int B = c->B;
        c->cap_raw = std::max(max_raw_map_points, VoxelFilter::SMALL_MAX + 1);
        c->max_kf = max_keyframes;
        c->kfs.resize(B); c->vox.resize(2 * (size_t)B);
        for (auto &v : c->vox) { v.init(); v.reserve(c->cap_raw); }
        for (auto &k : c->kfs) k.reserve((size_t)max_keyframes * 3 * c->cap_scan / 2);
        c->raw_map.ensure((size_t)B * 2 * c->cap_raw); c->ds_map.ensure((size_t)B * 2 * c->cap_raw);
        c->ds_map_n.ensure((size_t)B * 2);
        LLB_CUDA(cudaMemset(c->ds_map_n.p, 0, sizeof(int) * 2 * B));
        c->seg_dev.ensure((size_t)B * 3 * max_keyframes);
        for (int i = 0; i < RING; i++) c->seg_pin[i].ensure((size_t)B * 3 * max_keyframes);
        for (auto &g : c->grids) g.job(nullptr, nullptr, c->cap_raw);       // an assembled DS map is bounded by its raw size
        c->asm_req.resize(B); c->map_from_kf.assign(B, 0);
        c->kf_enabled = true;
        LLB_CUDA(cudaDeviceSynchronize());
        return (int)LLB_OK;
    });
}

int llb_batch_keyframe_add(llb_batch *c, int slot, int *id)
{
    return guarded(c, [&]() {
        if (slot < 0 || slot >= c->B) return (int)LLB_ERR_INVALID;
        if (!c->kf_enabled || c->pending || !c->have_results) return (int)LLB_ERR_STATE;
        const BatchResult &r = c->pin_results.p[slot];       // DS cloud sizes of the slot's last step
        const int n[3] = { r.ds[0], r.ds[1], r.ds[2] };
        float4 *dst[3];
        const int k = c->kfs[slot].add(n, dst);
        const float4 *ds = c->scan_ds.p + (size_t)slot * 5 * c->cap_scan;
        const float4 *src[3] = { ds, ds + c->cap_scan, ds + 2 * (size_t)c->cap_scan };
        for (int j = 0; j < 3; j++)
            if (n[j] > 0) { c->pending_copy.push_back(BatchCopy{ src[j], dst[j], n[j] }); c->pending_copy_max = std::max(c->pending_copy_max, n[j]); }
        if (id) *id = k;
        return (int)LLB_OK;
    });
}

int llb_batch_keyframe_count(llb_batch *c, int slot, int *n)
{
    if (!c || !n || slot < 0 || slot >= c->B || !c->kf_enabled) return LLB_ERR_INVALID;
    *n = c->kfs[slot].size();
    return LLB_OK;
}

int llb_batch_map_assemble(llb_batch *c, int slot, const int *ids, const float *poses, int n)
{
    return guarded(c, [&]() {
        if (slot < 0 || slot >= c->B || n < 0 || (n > 0 && (!ids || !poses))) return (int)LLB_ERR_INVALID;
        if (!c->kf_enabled) return (int)LLB_ERR_STATE;
        if (n > c->max_kf) return (int)LLB_ERR_CAPACITY;
        size_t rc = 0, rs = 0;
        for (int k = 0; k < n; k++) {
            if (ids[k] < 0 || ids[k] >= c->kfs[slot].size()) return (int)LLB_ERR_INVALID;
            const KeyFrameRec &kr = c->kfs[slot].rec(ids[k]);
            rc += kr.n[0]; rs += (size_t)kr.n[1] + kr.n[2];
        }
        if (rc > (size_t)c->cap_raw || rs > (size_t)c->cap_raw) return (int)LLB_ERR_CAPACITY;
        llb_batch::AsmReq &rq = c->asm_req[slot];
        rq.ids.assign(ids, ids + n); rq.poses.assign(poses, poses + 6 * (size_t)n);
        rq.rc = (int)rc; rq.rs = (int)rs; rq.pending = true;
        return (int)LLB_OK;
    });
}

int llb_batch_map_assemble_all(llb_batch *c, const int *ids, const float *poses, const int *offset)
{
    if (!c || !ids || !poses || !offset) return LLB_ERR_INVALID;
    for (int s = 0; s < c->B; s++) {
        const int n = offset[s + 1] - offset[s];
        if (n < 0) return LLB_ERR_INVALID;
        if (n == 0) continue;
        const int rc = llb_batch_map_assemble(c, s, ids + offset[s], poses + 6 * (size_t)offset[s], n);
        if (rc != LLB_OK) return rc;
    }
    return LLB_OK;
}

// which: 0 raw corner map, 1 raw surf map, 2 DS corner map, 3 DS surf map of the slot's last assembled map
int llb_batch_map_get(llb_batch *c, int slot, int which, llb_point *out, int cap, int *n)
{
    return guarded(c, [&]() {
        if (!n || slot < 0 || slot >= c->B || which < 0 || which > 3) return (int)LLB_ERR_INVALID;
        if (!c->kf_enabled || !c->map_from_kf[slot]) return (int)LLB_ERR_STATE;
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        int cnt = 0;
        if (which < 2) cnt = which == 0 ? c->asm_req[slot].rc : c->asm_req[slot].rs;
        else LLB_CUDA(cudaMemcpy(&cnt, c->ds_map_n.p + 2 * slot + (which - 2), sizeof(int), cudaMemcpyDeviceToHost));
        *n = cnt;
        if (!out) return (int)LLB_OK;
        if (cnt > cap) return (int)LLB_ERR_CAPACITY;
        const float4 *src = (which < 2 ? c->raw_map.p : c->ds_map.p) + (size_t)(2 * slot + (which & 1)) * c->cap_raw;
        std::vector<float4> tmp(std::max(cnt, 1));
        if (cnt > 0) LLB_CUDA(cudaMemcpy(tmp.data(), src, sizeof(float4) * cnt, cudaMemcpyDeviceToHost));
        for (int i = 0; i < cnt; i++) out[i] = llb_point{ tmp[i].x, tmp[i].y, tmp[i].z, 1.0f, tmp[i].w, 0.f, 0.f, 0.f };
        return (int)LLB_OK;
    });
}

// ------------------------------------------------------------------ featureAssociation of the slots

namespace {
void odom_ensure(llb_batch *c)
{
    if (c->od_ready) return;
    const int B = c->B;
    const llb_params &p = c->prm;
    c->oprm.nearest_sqdist = p.odom_nearest_sqdist; c->oprm.max_iter = p.odom_max_iterations;
    c->oprm.min_corr = p.odom_min_correspondences; c->oprm.degeneracy_thresh = p.odom_degeneracy_thresh;
    c->oprm.converge_deg = p.odom_converge_deg; c->oprm.converge_cm = p.odom_converge_cm;
    c->od_clouds.ensure((size_t)B * 4 * c->cap_scan); c->od_raw.ensure((size_t)B * 4 * c->cap_scan * 8);
    c->od_ind.ensure((size_t)B * 5 * c->cap_scan);
    c->od_states.ensure(B); c->od_pin_states.ensure(B);
    c->od_jobs_dev.ensure(B); c->od_jobs_pin.ensure(B);
    c->od_unp_dev.ensure((size_t)B * 4); c->od_unp_pin.ensure((size_t)B * 4);
    c->od_poses_dev.ensure((size_t)B * 6); c->od_poses_pin.ensure((size_t)B * 6);
    c->od_n.assign((size_t)B * 4, -1);
    launch_odom_state_init(c->od_states.p, B, c->stream);
    launch_odom_fill(c->od_ind.p, B * 5 * c->cap_scan, -1.f, c->stream);
    LLB_CUDA(cudaStreamSynchronize(c->stream));
    c->od_ready = true;
}
}  // namespace

int llb_batch_odom_set(llb_batch *c, int slot, const llb_point *corner_last, int ncl, const llb_point *surf_last, int nsl,
                       const llb_point *corner_sharp, int nsharp, const llb_point *surf_flat, int nflat)
{
    return guarded(c, [&]() {
        const llb_point *src[4] = { corner_last, surf_last, corner_sharp, surf_flat };
        const int n[4] = { ncl, nsl, nsharp, nflat };
        if (slot < 0 || slot >= c->B) return (int)LLB_ERR_INVALID;
        for (int k = 0; k < 4; k++) {
            if (n[k] < 0 || (n[k] > 0 && !src[k])) return (int)LLB_ERR_INVALID;
            if (n[k] > c->cap_scan) return (int)LLB_ERR_CAPACITY;
        }
        odom_ensure(c);
        for (int k = 0; k < 4; k++) {
            const size_t o = ((size_t)slot * 4 + k) * c->cap_scan;
            upload(c, c->B * 5 + slot * 4 + k, src[k], n[k], c->od_raw.p + o * 8, c->od_clouds.p + o, &c->od_pending, &c->od_pending_max);
            c->od_n[(size_t)slot * 4 + k] = n[k];
        }
        return (int)LLB_OK;
    });
}

int llb_batch_odom_optimize(llb_batch *c, float *T, llb_stats *stats_surf, llb_stats *stats_corner)
{
    return guarded(c, [&]() {
        if (!T) return (int)LLB_ERR_INVALID;
        if (!c->od_ready || c->pending) return (int)LLB_ERR_STATE;
        const int B = c->B;
        for (int s = 0; s < B; s++)
            for (int k = 0; k < 4; k++) if (c->od_n[(size_t)s * 4 + k] < 0) return (int)LLB_ERR_STATE;
        for (int s = 0; s < B; s++) {
            OdomBatchJob &j = c->od_jobs_pin.p[s];
            const float4 *base = c->od_clouds.p + (size_t)s * 4 * c->cap_scan;
            j.cornerLast = base; j.surfLast = base + c->cap_scan; j.sharp = base + 2 * (size_t)c->cap_scan;
            j.flat = base + 3 * (size_t)c->cap_scan;
            j.ncl = c->od_n[4 * s]; j.nsl = c->od_n[4 * s + 1]; j.nsharp = c->od_n[4 * s + 2]; j.nflat = c->od_n[4 * s + 3];
            j.ind = c->od_ind.p + (size_t)s * 5 * c->cap_scan; j.cap = c->cap_scan;
            j.st = c->od_states.p + s;
            for (int i = 0; i < 6; i++) c->od_poses_pin.p[6 * s + i] = T[6 * s + i];
        }
        const int nunp = (int)c->od_pending.size();
        for (int i = 0; i < nunp; i++) c->od_unp_pin.p[i] = c->od_pending[i];
        const int unp_max = c->od_pending_max;
        c->od_pending.clear(); c->od_pending_max = 0;
        LLB_CUDA(cudaEventRecord(c->ev0, c->stream));
        LLB_CUDA(cudaMemcpyAsync(c->od_jobs_dev.p, c->od_jobs_pin.p, sizeof(OdomBatchJob) * B, cudaMemcpyHostToDevice, c->stream));
        LLB_CUDA(cudaMemcpyAsync(c->od_poses_dev.p, c->od_poses_pin.p, sizeof(float) * 6 * B, cudaMemcpyHostToDevice, c->stream));
        if (nunp > 0) {
            LLB_CUDA(cudaMemcpyAsync(c->od_unp_dev.p, c->od_unp_pin.p, sizeof(BatchUnpack) * nunp, cudaMemcpyHostToDevice, c->stream));
            launch_batch_unpack(c->od_unp_dev.p, nunp, unp_max, c->stream);
            c->launches++;
        }
        launch_odom_batch_set_pose(c->od_jobs_dev.p, c->od_poses_dev.p, B, c->stream);
        launch_odom_batch(c->oprm, c->od_jobs_dev.p, B, c->stream);      // updateTransformation FA:1666-1695 for every slot
        c->launches += 2;
        LLB_CUDA(cudaEventRecord(c->ev1, c->stream));
        LLB_CUDA(cudaMemcpyAsync(c->od_pin_states.p, c->od_states.p, sizeof(OdomState) * B, cudaMemcpyDeviceToHost, c->stream));
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        float ms = 0.f;
        LLB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        for (int s = 0; s < B; s++) {
            const OdomState &st = c->od_pin_states.p[s];
            for (int i = 0; i < 6; i++) T[6 * s + i] = st.T[i];
            for (int w = 0; w < 2; w++) {
                llb_stats *o = w == 0 ? stats_surf : stats_corner;
                if (!o) continue;
                std::memset(&o[s], 0, sizeof(llb_stats));
                o[s].iterations = st.iters[w]; o[s].converged = st.converged[w]; o[s].n_correspondences = st.n_corr;
                o[s].is_degenerate = st.is_degenerate; o[s].skipped = st.skipped; o[s].device_ms = ms;
            }
        }
        return (int)LLB_OK;
    });
}

int llb_batch_set_profile(llb_batch *c, int on)
{
    if (!c) return LLB_ERR_INVALID;
    c->profile = on != 0;
    return LLB_OK;
}

int llb_batch_get_profile(llb_batch *c, float ms[6], int geometry[4])
{
    if (!c) return LLB_ERR_INVALID;
    if (ms) for (int k = 0; k < PROF_N; k++) ms[k] = c->prof_ms[k];
    if (geometry) { geometry[0] = c->lm_grid; geometry[1] = BATCH_ITER_THREADS; geometry[2] = c->grid_ctas; geometry[3] = c->qcap; }
    return LLB_OK;
}

// ---------------------------------------------------------------- feature extraction of the slots
int llb_batch_features_init(llb_batch *c, int n_scan, int horizon_scan)
{
    return guarded(c, [&]() {
        if (n_scan <= 0 || n_scan > FE_MAX_RINGS || horizon_scan < 16 || horizon_scan > 4096) return (int)LLB_ERR_INVALID;
        c->features.init(c->B, n_scan, horizon_scan, c->stream);
        c->features_done = false;
        return (int)LLB_OK;
    });
}

int llb_batch_features_extract(llb_batch *c, const llb_segmented_cloud *segs, int *counts, float *device_ms)
{
    return guarded(c, [&]() {
        if (!segs) return (int)LLB_ERR_INVALID;
        if (!c->features.ready() || c->pending) return (int)LLB_ERR_STATE;
        const int B = c->B;
        std::vector<FeSweepHost> sw(B);
        for (int s = 0; s < B; s++) {
            const llb_segmented_cloud &g = segs[s];
            FeatureExtractor &e = c->features.slot(s);
            if (g.n < 0 || !g.start_ring || !g.end_ring) return (int)LLB_ERR_INVALID;
            if (g.n > e.n_scan() * e.horizon()) return (int)LLB_ERR_CAPACITY;
            if (g.n > 0 && (!g.cloud || !g.ground_flag || !g.col_ind || !g.range)) return (int)LLB_ERR_INVALID;
            for (int r = 0; r < e.n_scan(); r++)
                if (g.start_ring[r] < 4 || g.end_ring[r] > g.n - 6 || g.end_ring[r] - g.start_ring[r] > e.horizon())
                    return (int)LLB_ERR_INVALID;
            sw[s] = FeSweepHost{ reinterpret_cast<const float *>(g.cloud), g.n, g.start_ring, g.end_ring, g.start_orientation,
                                 g.end_orientation, g.orientation_diff, g.ground_flag, g.col_ind, g.range };
        }
        LLB_CUDA(cudaEventRecord(c->ev0, c->stream));
        c->launches += c->features.extract(sw.data(), c->stream);
        LLB_CUDA(cudaEventRecord(c->ev1, c->stream));
        LLB_CUDA(cudaStreamSynchronize(c->stream));
        c->features_done = true;
        if (counts) for (int s = 0; s < B; s++) for (int k = 0; k < 4; k++) counts[4 * s + k] = c->features.slot(s).counts()[k];
        if (device_ms) LLB_CUDA(cudaEventElapsedTime(device_ms, c->ev0, c->ev1));
        return (int)LLB_OK;
    });
}

int llb_batch_features_get(llb_batch *c, int slot, int which, llb_point *out, int cap, int *n)
{
    return guarded(c, [&]() {
        if (!n || slot < 0 || slot >= c->B || which < 0 || which > 3) return (int)LLB_ERR_INVALID;
        if (!c->features_done) return (int)LLB_ERR_STATE;
        FeatureExtractor &e = c->features.slot(slot);
        const int cnt = e.counts()[which];
        *n = cnt;
        if (!out) return (int)LLB_OK;
        if (cnt > cap) return (int)LLB_ERR_CAPACITY;
        const float4 *src = e.host_cloud(which);
        for (int i = 0; i < cnt; i++) {
            out[i].x = src[i].x; out[i].y = src[i].y; out[i].z = src[i].z; out[i].w = 1.0f;
            out[i].intensity = src[i].w; out[i].c1 = out[i].c2 = out[i].c3 = 0.f;
        }
        return (int)LLB_OK;
    });
}

}  // extern "C"
