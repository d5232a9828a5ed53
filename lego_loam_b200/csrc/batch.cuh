// batch.cuh — batched multi-registration engine (BASELINE config 5, SURVEY 8(e) row 1: "independent
// sequences: per-GPU batching of many registrations into one grid").
//
// A batch owns B slots; every slot is one independent sequence with its own scan clouds, DS local map,
// spatial index and persistent LM state (isDegenerate / matP).  ONE step registers all B
// slots with a FIXED number of launches that does not depend on B:
//     1   unpack of every host cloud uploaded this step (PCL 32 B stride -> float4)
//     2   downsampleCurrentScan, MO:1067-1091: 3B filters (one CTA each, radix sort in shared memory), then the B
//         "total" filters - on a forked stream beside the map side of the step
//   1+19  slots with a key-frame request: assembly of their raw local maps (one launch) + their 2 map voxel filters
//         each (one set of 19 launches, job table); key-frames saved since the last step are copied first (1)
//     5   spatial-index build of every map that changed (replaces 2B kdtree->setInputCloud, MO:1333-1334)
//     2   prepare (pose, sin/cos, guard MO:1331) + ordering of the queries by kNN cost (speed only)
//   2xI   per LM iteration (MO:1336-1346): kNN (thread per query, cost-balanced warps), then fit + Jacobian rows + fp64
//         products (thread per query) whose last CTA per slot also performs the LM step; converged slots drop out
//     1   collect (pose + stats of all slots -> one D2H)
// whereas the single-registration path (s2m.cu) is ONE persistent kernel tuned for latency.  Throughput
// mode trades the persistent kernel's 16 warps/SM (128 registers for the fits) for kernels whose occupancy
// matches their phase: the kNN phase is latency-bound and wants many warps in flight.
#pragma once
#include "common.cuh"
#include "grid_index.cuh"
#include "s2m.cuh"
#include "voxel_dev.cuh"
#include "keyframes.cuh"
#include <vector>

namespace llb {

struct BatchReg {                    // device-resident record of one slot, read by the batched kernels
    const float4 *corner, *surf;     // laserCloudCornerLastDS, laserCloudSurfTotalLastDS
    const int *nc_dev, *ns_dev;
    MapIndexView cmap, smap;
    S2mState *st;
    int *qperm;                      // [cap] cell-ordered query permutation (corner part, then surf part)
    double *partials;                // [chunks of 32 queries][S2M_ACC]
    float4 *qprev;                   // [cap] per query: its map-frame position and 5th squared distance (-1: none) of the
                                     // previous LM iteration (seeds the bound of the next search)
    int cap;                         // query capacity of qperm / qprev
};

struct BatchSlotInfo { int nc, ns, done, n_chunks; };   // done: skipped by the guard MO:1331 (or no iterations asked for)
struct BatchQueue {                  // device-resident work hand-out of the registration kernel of one step
    int live;                        // slots that still have iterations to run
    unsigned *ctl;                   // [B] per slot: (LM iteration << 16) | next chunk to hand out; 0xffffffff = done
    BatchSlotInfo *slot;             // [B] what a warp needs to know about a slot, one 16-byte load
    long long *prof;                 // debug (LLB_ITER_PROF): per-warp cycles by phase, [warps][12]; nullptr = off
};

struct BatchUnpack {                 // one host cloud to compact: raw PCL points -> float4
    const float *src32; float4 *dst; int n;
};

struct BatchResult {                 // per slot, gathered into one contiguous D2H block
    float T[6];
    int iters, converged, n_corr, is_degenerate, skipped, nc, ns, pad;
    int ds[4];                       // sizes of cornerLastDS, surfLastDS, outlierLastDS, surfTotalLastDS
};

struct BatchCopy {                   // float4 cloud copy (DS clouds of a sweep -> key-frame arena)
    const float4 *src; float4 *dst; int n;
};

constexpr int BATCH_ITER_THREADS = 128;     // 4 autonomous warps; shared memory (candidate staging) sets the CTAs per SM
constexpr int BATCH_ITER_CTAS_PER_SM = 4;

// kernels (batch.cu)
void launch_batch_copy(const BatchCopy *jobs_dev, int count, int n_max, cudaStream_t s);
void launch_batch_unpack(const BatchUnpack *jobs_dev, int count, int n_max, cudaStream_t s);
void launch_batch_prepare(const BatchReg *regs, const float *poses_dev, int B, const S2mParams &prm, int max_iter, BatchQueue *queue,
                          cudaStream_t s);
void launch_batch_qsort(const BatchReg *regs, int B, int cap, cudaStream_t s);
int batch_lm_grid();       // CTAs of the persistent registration kernel (all co-resident)
void launch_batch_lm(const BatchReg *regs, int B, int grid, const S2mParams &prm, int max_iter, BatchQueue *queue, cudaStream_t s);
void launch_batch_collect(const BatchReg *regs, int B, BatchResult *out, BatchQueue *queue, cudaStream_t s);
void launch_batch_state_init(S2mState *st, int B, cudaStream_t s);

}  // namespace llb
