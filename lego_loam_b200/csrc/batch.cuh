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
    int *nn;                         // [5][cap] positions of the 5 neighbours in the map's sorted array
    int *qperm;                      // [cap] cell-ordered query permutation (corner part, then surf part)
    float *d5;                       // [cap] 5th squared distance, -1 when fewer than 5 candidates inside the gate
    double *partials;                // [fit blocks][S2M_ACC]
    int cap;                         // query slots of nn / d5
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

constexpr int BATCH_FIT_THREADS = 256;
constexpr int BATCH_KNN_THREADS = 256;

// kernels (batch.cu)
void launch_batch_copy(const BatchCopy *jobs_dev, int count, int n_max, cudaStream_t s);
void launch_batch_unpack(const BatchUnpack *jobs_dev, int count, int n_max, cudaStream_t s);
void launch_batch_prepare(const BatchReg *regs, const float *poses_dev, int B, const S2mParams &prm, cudaStream_t s);
int batch_knn_variant();   // 3 (default): cost-ordered queries + flattened walk; 1: row-by-row walk in scan order; 2: two-phase list
void launch_batch_qsort(const BatchReg *regs, int B, int cap, cudaStream_t s);
void launch_batch_knn(const BatchReg *regs, int B, int ctas_per_slot, const S2mParams &prm, cudaStream_t s);
void launch_batch_fit(const BatchReg *regs, int B, int fit_blocks, int iter, const S2mParams &prm, cudaStream_t s);
void launch_batch_collect(const BatchReg *regs, int B, BatchResult *out, cudaStream_t s);
void launch_batch_state_init(S2mState *st, int B, cudaStream_t s);

}  // namespace llb
