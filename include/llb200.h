/*
 * llb200.h — C ABI of lego_loam_b200: the B200-native (sm_100a) drop-in for the
 * scan-matching hot path of LeGO-LOAM.
 *
 * The reference (priseup/LeGO-LOAM) has no plugin / FFI surface: the boundary is
 * a set of C++ member functions of two monolithic classes.
 * Each entry point below names the reference member function(s) it replaces
 * (MO = LeGO-LOAM/src/mapOptmization.cpp, FA = LeGO-LOAM/src/featureAssociation.cpp,
 * UT = LeGO-LOAM/include/utility.h).  The C++ adapter classes that keep the
 * reference's own names and signatures on top of this ABI are in
 * lego_loam_b200/host/ (see INTEGRATION.md).
 *
 * Conventions
 *  - plain C, no exceptions cross the boundary; every call returns llb_status;
 *    on any error the pose passed in is left untouched (the reference's own
 *    behaviour when a guard fails, MO:1238, MO:1331, FA:1668).
 *  - host clouds use pcl::PointXYZI's layout (llb_point, 32 B);
 *    the *_dev family takes device pointers to compact float4 {x,y,z,intensity}.
 *  - a context owns one CUDA stream and all workspaces; it is not re-entrant
 *    (the reference runs the path under one mutex, MO:1497).
 *  - pose T[6] = {rx, ry, rz, tx, ty, tz}, p_map = Ry Rx Rz p + t (MO:513-527).
 */
#ifndef LLB200_H_
#define LLB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LLB_ABI_VERSION 1

typedef enum {
    LLB_OK = 0,
    LLB_ERR_INVALID = 1,      /* bad argument */
    LLB_ERR_CUDA = 2,         /* CUDA runtime error, see llb_last_error */
    LLB_ERR_NO_DEVICE = 3,    /* no usable sm_100 device: there is NO CPU fallback */
    LLB_ERR_CAPACITY = 4,     /* caller buffer too small */
    LLB_ERR_STATE = 5         /* call order violated (e.g. optimise before map/scan set) */
} llb_status;

/* pcl::PointXYZI: union{float data[4]; {x,y,z}} + union{{intensity}; float data_c[4]} */
typedef struct {
    float x, y, z, w;         /* w = data[3], 1.0f in PCL */
    float intensity, c1, c2, c3;
} llb_point;

/* every tunable the hot path reads; defaults = the reference's constants */
typedef struct {
    float corner_leaf;            /* 0.2  MO:249 */
    float surf_leaf;              /* 0.4  MO:250 */
    float outlier_leaf;           /* 0.4  MO:251 */
    float knn_max_sqdist;         /* 1.0  MO:1101, MO:1183 */
    int   s2m_max_iterations;     /* 10   MO:1336 */
    int   s2m_min_correspondences;/* 50   MO:1238 */
    float s2m_degeneracy_thresh;  /* 100  MO:1287 */
    float s2m_converge_deg;       /* 0.05 MO:1323 */
    float s2m_converge_cm;        /* 0.05 MO:1323 */
    int   corner_map_min;         /* 10   MO:1331 (strict >) */
    int   surf_map_min;           /* 100  MO:1331 (strict >) */
    float odom_nearest_sqdist;    /* 25   UT:125 */
    int   odom_max_iterations;    /* 25   FA:1671, FA:1683 */
    int   odom_min_correspondences;/* 10  FA:1677, FA:1690 */
    float odom_degeneracy_thresh; /* 10   FA:1338, FA:1439 */
    float odom_converge_deg;      /* 0.1  FA:1373 */
    float odom_converge_cm;       /* 0.1  FA:1373 */
    int   max_grid_cells;         /* spatial-index cell budget per map (default 1<<23) */
    int   pin_host_clouds;        /* 0 (default): host clouds are copied into the context's pinned staging during the
                                     call and may be reused as soon as it returns.  1: the caller's buffers are
                                     page-locked once (cudaHostRegister, cached per pointer) and DMA'd in place:
                                     no host copy, but the caller must keep a cloud unchanged until the next
                                     blocking call on this context (llb_s2m_optimize / llb_s2m_result / llb_synchronize
                                     / any getter).  The reference's clouds are long-lived class members, so the
                                     adapters can enable it. */
    int   s2m_max_ctas;           /* cap on the CTAs of the persistent scan-to-map kernel; 0 (default) = one per SM:
                                     lowest latency for ONE registration.  When several contexts (sequences) share a
                                     GPU a cap of ~1/4 of the SMs lets their registrations overlap instead of queueing
                                     behind each other's cooperative launches (throughput mode). */
} llb_params;

typedef struct {
    int   iterations;             /* LM iterations executed */
    int   converged;              /* 1 if the convergence test ended the loop */
    int   n_correspondences;      /* rows of the last iteration (laserCloudOri size) */
    int   is_degenerate;          /* isDegenerate after the call */
    int   skipped;                /* 1 if the map-size guard MO:1331 / FA:1668 skipped the work */
    int   n_corner_ds, n_surf_ds; /* query counts used */
    float device_ms;              /* CUDA-event time of the device work of this call */
} llb_stats;

typedef struct llb_ctx llb_ctx;

/* ---- lifetime ---- */
int  llb_abi_version(void);
void llb_params_default(llb_params *p);
int  llb_create(const llb_params *p /* NULL = defaults */, int device, llb_ctx **out);
int  llb_destroy(llb_ctx *ctx);
const char *llb_last_error(const llb_ctx *ctx);
/* the CUDA stream (cudaStream_t) every call of this context is ordered on */
void *llb_stream(llb_ctx *ctx);
int  llb_synchronize(llb_ctx *ctx);

/* Optional: pre-size every workspace (sweep clouds of up to max_scan_points each, raw local maps of up to
 * max_raw_map_points each, a surrounding set of up to max_keyframes key-frames) so that no call below those bounds
 * allocates: workspaces otherwise grow on demand (cudaMalloc / cudaFree, a device synchronisation and a latency
 * spike of milliseconds whenever a cloud outgrows its buffer - the reference's own vectors reallocate the same way). */
int  llb_reserve(llb_ctx *ctx, int max_scan_points, int max_raw_map_points, int max_keyframes);

/* ---- pcl::VoxelGrid<PointXYZI>::filter (setLeafSize(leaf,leaf,leaf), defaults)
 *      call sites MO:1058-1063, MO:1070-1089, FA:779-780 ---- */
int llb_voxel_downsample(llb_ctx *ctx, const llb_point *in, int n, float leaf,
                         llb_point *out, int out_capacity, int *m);

/* ---- mapOptimization ---- */
/* laserCloudCornerFromMapDS / laserCloudSurfFromMapDS hand-over + the two
 * kdtree->setInputCloud calls of MO:1333-1334 (device spatial index build) */
int llb_map_set_ds(llb_ctx *ctx, const llb_point *corner_ds, int mc, const llb_point *surf_ds, int ms);
/* raw local map: tail of extractSurroundingKeyFrames MO:1057-1064 (two voxel
 * filters) followed by the index build */
int llb_map_set_raw(llb_ctx *ctx, const llb_point *corner, int rc, const llb_point *surf, int rs);
/* which: 0 corner, 1 surf.  n receives the size even when out == NULL */
int llb_map_get_ds(llb_ctx *ctx, int which, llb_point *out, int capacity, int *n);
/* laserCloudCornerLast / SurfLast / OutlierLast (handlers MO:608-627) */
int llb_scan_set(llb_ctx *ctx, const llb_point *corner_last, int nc, const llb_point *surf_last, int ns,
                 const llb_point *outlier_last, int no);
/* downsampleCurrentScan MO:1067-1091; counts = {cornerDS, surfDS, outlierDS, surfTotalDS}
 * (counts may be NULL: then no host synchronisation happens) */
int llb_downsample_current_scan(llb_ctx *ctx, int counts[4]);
/* which: 0 cornerLastDS, 1 surfLastDS, 2 outlierLastDS, 3 surfTotalLastDS */
int llb_scan_get_ds(llb_ctx *ctx, int which, llb_point *out, int capacity, int *n);
/* one iteration of the loop body MO:1338-1345: cornerOptimization + surfOptimization
 * + LMOptimization(iter).  T is transformTobeMapped in/out. */
int llb_s2m_iterate(llb_ctx *ctx, float T[6], int iter, int *converged, int *n_correspondences);
/* scan2MapOptimization MO:1329-1350 without transformUpdate's bookkeeping: guard,
 * index (already built by llb_map_set_*), <= 10 fused iterations with the
 * convergence test on device, no host round trip inside. */
int llb_s2m_optimize(llb_ctx *ctx, float T[6], llb_stats *stats);
/* the same, split so that several contexts (independent sequences on one GPU) can be kept in flight
 * from one host thread: _async enqueues the whole registration and returns without synchronising,
 * _result waits for it and returns pose + stats */
int llb_s2m_optimize_async(llb_ctx *ctx, const float T[6]);
int llb_s2m_result(llb_ctx *ctx, float T[6], llb_stats *stats);
/* laserCloudOri / coeffSel of the last iteration run by llb_s2m_iterate, in the
 * reference's order (corner rows then surf rows, each in query order) */
int llb_get_correspondences(llb_ctx *ctx, llb_point *ori, llb_point *coeff, int capacity, int *n);
/* diagnostics: 5-NN original map indices (-1 where < 5 candidates in range) and
 * squared distances per query of the last llb_s2m_iterate; which: 0 corner, 1 surf */
int llb_get_knn(llb_ctx *ctx, int which, int *idx5, float *sqdist5, int capacity, int *n);
/* AtA (36), AtB (6), X (6) of the last LM step */
int llb_get_normal_equations(llb_ctx *ctx, float AtA[36], float AtB[6], float X[6]);
/* isDegenerate / matP persist across registrations (MO:202-203) */
int llb_get_degeneracy(llb_ctx *ctx, int *is_degenerate, float matP[36]);
int llb_set_degeneracy(llb_ctx *ctx, int is_degenerate, const float matP[36]);

/* ---- device-resident key-frame store ----
 * cornerCloudKeyFrames / surfCloudKeyFrames / outlierCloudKeyFrames (MO:128-130) kept in HBM, so that a registration
 * moves only the new sweep over PCIe and the raw local map never exists on the host. */
/* saveKeyFramesAndFactor's cloud part MO:1443-1453: stores the CURRENT laserCloudCornerLastDS / SurfLastDS /
 * OutlierLastDS (already on the device after llb_downsample_current_scan); *id = index in the store */
int llb_keyframe_add(llb_ctx *ctx, int *id);
/* the same from host clouds (replaying / restoring a store) */
int llb_keyframe_add_clouds(llb_ctx *ctx, const llb_point *corner_ds, int nc, const llb_point *surf_ds, int ns,
                            const llb_point *outlier_ds, int no, int *id);
int llb_keyframe_count(llb_ctx *ctx, int *n);
int llb_keyframe_clear(llb_ctx *ctx);
/* cloud part of extractSurroundingKeyFrames: for the n key-frames ids[] IN ORDER (surroundingExistingKeyPosesID,
 * MO:1033-1049, or the recent-frames queue MO:962-1001) with their cloudKeyPoses6D entries poses[n][6] =
 * {roll, pitch, yaw, x, y, z}: transformPointCloud (MO:545-575) + concatenation (MO:1050-1054) in ONE launch,
 * then the two map voxel filters (MO:1057-1064) and the index build (MO:1333-1334).  The sin/cos of the poses are
 * taken on the host with the float libm overloads, as the reference does (MO:529-543). */
int llb_map_assemble(llb_ctx *ctx, const int *ids, const float *poses, int n);
/* laserCloudCornerFromMap (0) / laserCloudSurfFromMap (1) of the last llb_map_assemble (parity checks) */
int llb_map_get_raw(llb_ctx *ctx, int which, llb_point *out, int capacity, int *n);

/* ---- loop closure + global map on the device (loopClosureEnableFlag is off by default, UT:104) ----
 * The host keeps what it keeps in the reference: the radius search over the key poses (MO:824-837, MO:770-778), the 30 s
 * test, the pose-graph factor (MO:919-944).  The device does the cloud work on its key-frame store. */
typedef struct {
    int    max_iterations;                  /* 100   MO:894 */
    double max_correspondence_distance;     /* 100   MO:893 */
    double transformation_epsilon;          /* 1e-6  MO:895 */
    double euclidean_fitness_epsilon;       /* 1e-6  MO:896 */
} llb_loop_params;
typedef struct {
    float  T[16];                           /* icp.getFinalTransformation(), row-major */
    int    has_converged;                   /* icp.hasConverged() */
    int    iterations;
    int    convergence_state;               /* 0 too few correspondences, 1 iterations, 2 transformation epsilon, 3 absolute
                                               MSE, 4 relative MSE (pcl::registration::DefaultConvergenceCriteria) */
    int    n_correspondences;               /* of the last iteration */
    double fitness_score;                   /* icp.getFitnessScore() */
    double sums[17];                        /* of the last pass: n, sum p, sum q, sum q p^T, sum d^2 (parity checks) */
    float  device_ms;
    int    n_source, n_target;
} llb_icp_result;
void llb_loop_params_default(llb_loop_params *p);
/* cloud part of detectLoopClosure MO:838-861: latestSurfKeyFrameCloud = corner + surf clouds of key-frame latest_id at
 * latest_pose {roll, pitch, yaw, x, y, z} (points with (int)intensity >= 0, MO:845-849); nearHistorySurfKeyFrameCloud =
 * corner + surf clouds of hist_ids[] (closestHistoryFrameID - 25 .. + 25 clipped, MO:853-858) at hist_poses[], then
 * VoxelGrid(history_leaf) (0.4, MO:253).  counts = {latest, history DS} */
int llb_loop_set_clouds(llb_ctx *ctx, int latest_id, const float latest_pose[6], const int *hist_ids, const float *hist_poses,
                        int n_hist, float history_leaf, int counts[2]);
/* the same two clouds handed over by a caller that built them itself (the drop-in members of the adapter) */
int llb_loop_set_clouds_host(llb_ctx *ctx, const llb_point *latest, int n_latest, const llb_point *history_ds, int n_hist);
/* pcl::IterativeClosestPoint<PointType, PointType>::align + getFitnessScore as performLoopClosure configures them
 * (MO:892-904), PCL 1.8's algorithm: one persistent kernel for all iterations (p NULL = the reference's settings) */
int llb_loop_icp(llb_ctx *ctx, const llb_loop_params *p, llb_icp_result *out);
/* which: 0 latestSurfKeyFrameCloud, 1 nearHistorySurfKeyFrameCloud, 2 nearHistorySurfKeyFrameCloudDS, 3 globalMapKeyFramesDS */
int llb_loop_get_cloud(llb_ctx *ctx, int which, llb_point *out, int capacity, int *n);
/* nearest target index + squared distance of every source point in the last search of llb_loop_icp (the fitness pass) */
int llb_loop_get_nn(llb_ctx *ctx, int *idx, float *sqdist, int capacity, int *n);
/* cloud part of publishGlobalMap MO:780-788: corner + surf + outlier clouds of ids[] (globalMapKeyPosesDS order) at
 * poses[], then VoxelGrid(leaf) (0.4, MO:257); read the result with llb_loop_get_cloud(which = 3) */
int llb_global_map_assemble(llb_ctx *ctx, const int *ids, const float *poses, int n, float leaf, int *n_out);

/* ---- featureAssociation: feature extraction ----
 * What laserCloudHandler / laserCloudInfoHandler leave in the node (FA:461-489): segmentedCloud in the LIDAR frame with
 * intensity = row + col / 10000 (IP:253), and cloud_msgs::cloud_info. */
typedef struct {
    const llb_point *cloud; int n;          /* segmentedCloud */
    const int *start_ring, *end_ring;       /* startRingIndex / endRingIndex, n_scan entries each (IP:318, IP:358) */
    float start_orientation, end_orientation, orientation_diff;   /* IP:199-211 */
    const unsigned char *ground_flag;       /* segmentedCloudGroundFlag[n] */
    const unsigned *col_ind;                /* segmentedCloudColInd[n] */
    const float *range;                     /* segmentedCloudRange[n] */
} llb_segmented_cloud;
/* N_SCAN / Horizon_SCAN (UT:63-84); allocates the per-point state the reference keeps between sweeps (FA:210-223) */
int llb_features_init(llb_ctx *ctx, int n_scan, int horizon_scan);
/* adjustDistortion (IMU branch FA:525-613 when llb_features_set_imu gave ring buffers), calculateSmoothness, markOccludedPoints, extractFeatures
 * = runFeatureAssociation FA:1827-1833.  counts = sizes of cornerPointsSharp, cornerPointsLessSharp, surfPointsFlat,
 * surfPointsLessFlat.  Selection, order and coordinates are bit-identical to the reference (std::sort's order of equal
 * curvatures included), and so are the intensities: the kernel restates glibc's atan2f operation by operation. */
int llb_features_extract(llb_ctx *ctx, const llb_segmented_cloud *seg, int counts[4], float *device_ms);
/* which: 0 cornerPointsSharp, 1 cornerPointsLessSharp, 2 surfPointsFlat, 3 surfPointsLessFlat, 4 segmentedCloud after
 * adjustDistortion, 5 / 6 laserCloudCornerLast / laserCloudSurfLast after llb_features_publish_last */
int llb_features_get(llb_ctx *ctx, int which, llb_point *out, int capacity, int *n);
/* cloudCurvature / cloudNeighborPicked / cloudLabel of the last sweep (parity checks) */
int llb_features_get_state(llb_ctx *ctx, float *curvature, int *neighbor_picked, int *label, int capacity);
/* cloud part of publishCloudsLast (FA:1759-1788) after updateTransformation: TransformToEnd (FA:885-953, no IMU data) of
 * cornerPointsLessSharp / surfPointsLessFlat with transformCur, the results become laserCloudCornerLast /
 * laserCloudSurfLast of the odometry (= llb_odom_set_last on them, without leaving the device; the clouds are always
 * indexed, see llb_odom_set_last).  llb_features_get(which = 5 / 6) reads them back. */
int llb_features_publish_last(llb_ctx *ctx, const float transformCur[6]);
/* ---- IMU branches of featureAssociation (imuHandler FA:417-448 and AccumulateIMUShiftAndRotation
 * FA:390-415 stay on the host: one message at a time, a few flops each — host/lego_loam_b200.hpp keeps the ring
 * buffers; the per-point work, adjustDistortion's IMU branch FA:525-613 and TransformToEnd's IMU terms FA:927-950,
 * runs on the device) ---- */
#define LLB_IMU_QUEUE 200                   /* imuQueLength UT:109 */
typedef struct llb_imu_queue {              /* the ring buffers FA:82-135 as they stand when the sweep arrives */
    double time[LLB_IMU_QUEUE];             /* imuTime */
    float roll[LLB_IMU_QUEUE], pitch[LLB_IMU_QUEUE], yaw[LLB_IMU_QUEUE];
    float velo[3][LLB_IMU_QUEUE];           /* imuVeloX / Y / Z */
    float shift[3][LLB_IMU_QUEUE];          /* imuShiftX / Y / Z */
    float angular[3][LLB_IMU_QUEUE];        /* imuAngularRotationX / Y / Z */
    double time_scan_cur;                   /* timeScanCur FA:453 */
    int pointer_last;                       /* imuPointerLast; < 0: no message yet (FA:525 skips the branch) */
    int pointer_last_iteration;             /* imuPointerLastIteration FA:489 / FA:616 */
} llb_imu_queue;
typedef struct llb_imu_sweep {              /* what adjustDistortion's IMU branch leaves in the members */
    float start[9];                         /* imuRoll/Pitch/YawStart, imuVeloX/Y/ZStart, imuShiftX/Y/ZStart (point 0) */
    float angular_cur[3];                   /* imuAngularRotationX/Y/ZCur (point 0, FA:580-594) */
    float cur[3];                           /* imuRoll/Pitch/YawCur after the last point */
    float velo_from_start_cur[3];           /* imuVeloFromStartX/Y/ZCur after the last point (valid when has_velo) */
    int valid, has_velo;                    /* valid: the branch ran; has_velo: the sweep had a point after the first */
} llb_imu_sweep;
typedef struct llb_imu_end {                /* IMU terms of TransformToEnd FA:927-950 */
    float cs_start[6];                      /* cos, sin of imuRollStart; of imuPitchStart; of imuYawStart (FA:317-324) */
    float shift_from_start[3];              /* imuShiftFromStartX/Y/Z */
    float last[3];                          /* imuRollLast, imuPitchLast, imuYawLast */
} llb_imu_end;
/* ring buffers for the NEXT llb_features_extract / llb_projection_to_features (copied; NULL: back to "no IMU data") */
int llb_features_set_imu(llb_ctx *ctx, const llb_imu_queue *queue);
/* members after the last extraction (out->valid = 0 when the branch did not run) */
int llb_features_get_imu(llb_ctx *ctx, llb_imu_sweep *out);
/* llb_features_publish_last with the IMU terms of TransformToEnd (NULL = llb_features_publish_last) */
int llb_features_publish_last_imu(llb_ctx *ctx, const float transformCur[6], const llb_imu_end *imu);
/* SM cycles of the slowest ring of the last sweep (profiling): [0] sort phase, [1] picks; then the slowest warp / ring
 * per part: [2] partitions, [3] leaf ranges, [4] edge picks, [5] flat picks, [6..9] reserved */
int llb_features_get_profile(llb_ctx *ctx, int cycles[10]);
/* cornerPointsSharp / surfPointsFlat of the last extraction become the odometry's features without leaving the device
 * (= llb_odom_set_features on them) */
int llb_features_to_odometry(llb_ctx *ctx);

/* ---- featureAssociation ---- */
/* laserCloudCornerLast / laserCloudSurfLast + kdtree rebuild (FA:1615-1619, FA:1774-1788) */
int llb_odom_set_last(llb_ctx *ctx, const llb_point *corner_last, int ncl, const llb_point *surf_last, int nsl);
/* cornerPointsSharp / surfPointsFlat of the current sweep */
int llb_odom_set_features(llb_ctx *ctx, const llb_point *corner_sharp, int nsharp, const llb_point *surf_flat, int nflat);
/* updateTransformation FA:1666-1695: both <= 25-iteration loops on device. T = transformCur in/out */
int llb_odom_optimize(llb_ctx *ctx, float T[6], llb_stats *stats_surf, llb_stats *stats_corner);
/* single steps for per-function parity: which 0 = surf (findCorrespondingSurfFeatures +
 * calculateTransformationSurf), 1 = corner.  *more = the reference's return value
 * (false = converged, C8); *n_correspondences = laserCloudOri size; the LM step is
 * skipped when it is < 10 (FA:1677) */
int llb_odom_iterate(llb_ctx *ctx, int which, float T[6], int iter, int *more, int *n_correspondences);
int llb_odom_get_correspondences(llb_ctx *ctx, llb_point *ori, llb_point *coeff, int capacity, int *n);
/* which: 0 = surf (pointSearchSurfInd1 / 2 / 3, FA:148-152), 1 = corner (pointSearchCornerInd1 / 2; ind3 untouched) */
int llb_odom_get_search_ind(llb_ctx *ctx, int which, float *ind1, float *ind2, float *ind3, int capacity, int *n);
int llb_odom_get_degeneracy(llb_ctx *ctx, int *is_degenerate, float matP[9]);

/* ---- device-resident family (inputs already in HBM as float4 {x,y,z,intensity}) ---- */
int llb_map_set_ds_dev(llb_ctx *ctx, const void *corner_ds_f4, int mc, const void *surf_ds_f4, int ms);
int llb_map_set_raw_dev(llb_ctx *ctx, const void *corner_f4, int rc, const void *surf_f4, int rs);
/* the sweep clouds are borrowed, not copied: keep them valid and unchanged until the registration has finished */
int llb_scan_set_dev(llb_ctx *ctx, const void *corner_f4, int nc, const void *surf_f4, int ns,
                     const void *outlier_f4, int no);
/* pose in/out in device memory (6 floats); nothing is copied to the host */
int llb_s2m_optimize_dev(llb_ctx *ctx, float *T_dev);
/* sharded large-map mode (BASELINE config 4): this rank accumulates the normal
 * equations of ITS share of the queries into 28 doubles (21 upper-tri AtA, 6 AtB,
 * count) at a device address the caller all-reduces (NCCL) before llb_s2m_solve */
int llb_s2m_accumulate(llb_ctx *ctx, int iter, int rank, int world, double **acc28_dev);
int llb_s2m_solve(llb_ctx *ctx, int iter, int *converged);
/* the same exchange FUSED into the persistent kernel (no NCCL call, no host round trip per iteration): every rank's
 * CTA 0 stores its 28 sums into every peer's mailbox over NVLink (P2P stores into cudaIpc-mapped memory), waits for
 * the peers' flags, adds the contributions in rank order - bit-identical normal equations on all ranks - and takes the
 * identical 6x6 LM step.  Set-up: each rank exports its mailbox handle, the 64-byte handles are all-gathered by the
 * caller (torch.distributed / MPI: plumbing) and imported.  llb_s2m_optimize_sharded must then be called by every rank
 * with the same map, scan and pose; a peer that never arrives is reported as LLB_ERR_STATE after ~2 s, not a hang. */
int llb_p2p_export(llb_ctx *ctx, unsigned char handle[64]);
int llb_p2p_import(llb_ctx *ctx, int rank, int world, const unsigned char *handles /* world x 64 bytes */);
int llb_s2m_optimize_sharded(llb_ctx *ctx, float T[6], llb_stats *stats);
/* ---- the same registration with the MAP sharded (preferred form): every rank is given the same raw local
 * map (its key-frame stores are replicas: each key-frame crossed PCIe once) but keeps, voxel-filters (MO:1057-1064) and
 * indexes (MO:1333-1334) only a slab of it - the voxels that can hold a centroid within the kNN gate of a query inside
 * the slab - on the lattice of the whole map, so voxel membership, order and centroids are those of the unsharded
 * filter and the map-side work divides by the number of ranks.  The slab borders are quantiles of a deterministic
 * sample of the raw map (one small D2H + host sort inside the call).  llb_s2m_optimize_sharded /
 * llb_s2m_accumulate(rank 0, world 1) then take the queries whose mapped position lies inside the slab; the exchange
 * per LM iteration stays the 28 fp64 sums.  The guard MO:1331 needs the sizes of the UNSHARDED DS maps: all-reduce
 * llb_shard_info.ds_owned over the ranks (plumbing) and hand the sums to llb_map_shard_set_global; without that call
 * the guard looks at this rank's part.  While a context holds a slab (world > 1) the single-rank entry points
 * llb_s2m_optimize / _async / _dev / llb_s2m_iterate return LLB_ERR_STATE: they would see this rank's queries only. */
typedef struct {
    int axis; float lo, hi;                 /* this rank owns mapped coordinates lo <= p[axis] < hi */
    int rank, world;
    int raw_kept[2];                        /* raw corner / surf points this rank filtered */
    int ds_local[2];                        /* centroids it holds (slab + halo) */
    int ds_owned[2];                        /* centroids inside the slab: sums over the ranks = sizes of the unsharded maps */
} llb_shard_info;
int llb_map_set_raw_sharded(llb_ctx *ctx, const llb_point *corner, int rc, const llb_point *surf, int rs, int rank, int world);
int llb_map_set_raw_sharded_dev(llb_ctx *ctx, const void *corner_f4, int rc, const void *surf_f4, int rs, int rank, int world);
int llb_map_shard_info(llb_ctx *ctx, llb_shard_info *out);
/* the slab planning alone (pure host function, no context): sample_xyz = nsamp x {x, y, z} map points; every rank must
 * pass the same sample.  axis = longest extent of the sample, [lo, hi) = quantile interval of `rank` (open at the ends) */
int llb_shard_plan(const float *sample_xyz, int nsamp, int rank, int world, int *axis, float *lo, float *hi);
int llb_map_shard_set_global(llb_ctx *ctx, const int global_ds[2]);
int llb_s2m_pose_set(llb_ctx *ctx, const float T[6]);
int llb_s2m_pose_get(llb_ctx *ctx, float T[6]);

/* measurement hook for the roofline line of bench.py: launches the fused K3+K4 iteration
 * kernel `reps` times on the context's stream at pose T with the LM step disabled (the
 * state is left untouched), bracketed by CUDA events on that stream; *ms_per_launch is the
 * average device time of one launch, *n_queries the number of queries one launch processes */
int llb_s2m_time_iteration(llb_ctx *ctx, const float T[6], int reps, float *ms_per_launch, int *n_queries);

/* clock64 stamps taken by CTA 0 during the last iteration of the persistent scan-to-map
 * kernel: start, end of phase A (kNN), B (fits), C (products), first grid sync, reduction,
 * LM step, second grid sync (SM clock cycles; diagnostics for profiles/) */
int llb_s2m_get_profile(llb_ctx *ctx, int iter, long long stamps[8]);
/* per-CTA cycles {phase A, phase B, phase C, wait at the first grid barrier} of the last iteration of the
 * last run: out[4*cta + k]; *n_ctas receives the grid size (diagnostics for profiles/) */
int llb_s2m_get_cta_profile(llb_ctx *ctx, double *out, int capacity_ctas, int *n_ctas);

/* ---- imageProjection on the device (IP = LeGO-LOAM/src/imageProjection.cpp) ----
 * llb_projection_init: the sensor block of utility.h (UT:62-84): N_SCAN, Horizon_SCAN, ang_res_x, ang_res_y,
 * groundScanInd.  llb_projection_process replaces cloudHandler IP:181-197 without the publishing: findStartEndAngle
 * (IP:199-211), projectPointCloud (IP:213-257, useCloudRing: the row of a point is its ring), groundRemoval
 * (IP:259-310), cloudSegmentation (IP:312-368) with labelComponents (IP:370-448) as connected-component labelling.
 * cloud: the sweep in firing order without NaN points (IP:170), pcl::PointXYZI layout; ring: its ring channel.
 * Results stay on the device: llb_projection_get_* copy them out (segmentedCloud + cloud_msgs::cloud_info fields,
 * outlierCloud, the three images for parity checks), llb_projection_to_features hands them to the feature extraction
 * (llb_features_init with the same N_SCAN / Horizon_SCAN) without leaving the device: raw sweep -> features with one
 * upload per sweep. */
int llb_projection_init(llb_ctx *ctx, int n_scan, int horizon_scan, float ang_res_x, float ang_res_y, int ground_scan_ind);
int llb_projection_process(llb_ctx *ctx, const llb_point *cloud, const uint16_t *ring, int n, int *n_segmented,
                           int *n_outlier, float *device_ms /* may be NULL */);
/* which: 0 segmentedCloud, 1 outlierCloud */
int llb_projection_get_cloud(llb_ctx *ctx, int which, llb_point *out, int capacity, int *n);
/* cloud_info of the last sweep: start / end ring index (N_SCAN each), {startOrientation, endOrientation,
 * orientationDiff}, per segmented point: ground flag, column index, range (capacity entries each; any may be NULL) */
int llb_projection_get_info(llb_ctx *ctx, int32_t *start_ring, int32_t *end_ring, float orientation[3],
                            uint8_t *ground_flag, uint32_t *col_ind, float *range, int capacity);
/* rangeMat (FLT_MAX: no return), groundMat, labelMat of the last sweep, N_SCAN x Horizon_SCAN each (parity checks) */
int llb_projection_get_images(llb_ctx *ctx, float *range_mat, int8_t *ground_mat, int32_t *label_mat);
/* adjustDistortion .. extractFeatures (llb_features_extract) on the segmented cloud of the last llb_projection_process */
int llb_projection_to_features(llb_ctx *ctx, int counts[4], float *device_ms /* may be NULL */);

/* number of kernels launched by this context since creation (bench.py gpu_launches) */
long long llb_launch_count(const llb_ctx *ctx);

/* ---- batched multi-registration engine (BASELINE config 5: batches of independent sequences) ----
 * n_slots independent sequences share one GPU and one stream; llb_batch_register performs, for EVERY slot,
 * downsampleCurrentScan (MO:1067-1091) + the two kdtree->setInputCloud replacements (MO:1333-1334, only for
 * slots whose map was set since the last step) + scan2MapOptimization (MO:1329-1350) with a number of kernel
 * launches that does not depend on n_slots.  Each slot keeps its own isDegenerate / matP across steps (C6).
 * Results are bit-identical to n_slots separate llb_ctx registrations up to the fp64 summation order of the
 * normal equations.  max_scan_points bounds every one of the three scan clouds of a slot (<= 16384, and surf + outlier
 * of one sweep <= 16384: the four filters of downsampleCurrentScan sort in shared memory; VLP-16 / HDL-32E sweeps fit,
 * VLS-128 sweeps use llb_ctx); max_map_points bounds each DS map. */
typedef struct llb_batch llb_batch;
int  llb_batch_create(const llb_params *p /* NULL = defaults */, int device, int n_slots, int max_scan_points,
                      int max_map_points, llb_batch **out);
int  llb_batch_destroy(llb_batch *b);
const char *llb_batch_last_error(const llb_batch *b);
void *llb_batch_stream(llb_batch *b);
int  llb_batch_slots(const llb_batch *b);
long long llb_batch_launch_count(const llb_batch *b);
/* laserCloudCornerLast / SurfLast / OutlierLast of one slot (handlers MO:608-627); host clouds are DMA'd
 * asynchronously and must stay unchanged until the next llb_batch_result (params.pin_host_clouds as for llb_ctx) */
int  llb_batch_scan_set(llb_batch *b, int slot, const llb_point *corner_last, int nc, const llb_point *surf_last,
                        int ns, const llb_point *outlier_last, int no);
/* laserCloudCornerFromMapDS / SurfFromMapDS of one slot; a slot whose map is not set again keeps its index */
int  llb_batch_map_set_ds(llb_batch *b, int slot, const llb_point *corner_ds, int mc, const llb_point *surf_ds, int ms);
/* device-resident inputs (float4 {x,y,z,intensity}); the pointers are borrowed until the step has finished */
int  llb_batch_scan_set_dev(llb_batch *b, int slot, const void *corner_f4, int nc, const void *surf_f4, int ns,
                            const void *outlier_f4, int no);
int  llb_batch_map_set_ds_dev(llb_batch *b, int slot, const void *corner_ds_f4, int mc, const void *surf_ds_f4, int ms);
/* the same for ALL slots with one call: arrays of n_slots pointers / counts */
int  llb_batch_scan_set_all(llb_batch *b, const llb_point *const *corner_last, const int *nc,
                            const llb_point *const *surf_last, const int *ns, const llb_point *const *outlier_last, const int *no);
int  llb_batch_map_set_ds_all(llb_batch *b, const llb_point *const *corner_ds, const int *mc,
                              const llb_point *const *surf_ds, const int *ms);
int  llb_batch_scan_set_dev_all(llb_batch *b, const void *const *corner_f4, const int *nc, const void *const *surf_f4,
                                const int *ns, const void *const *outlier_f4, const int *no);
int  llb_batch_map_set_ds_dev_all(llb_batch *b, const void *const *corner_ds_f4, const int *mc,
                                  const void *const *surf_ds_f4, const int *ms);
/* T: n_slots x 6 transformTobeMapped in/out; stats: n_slots entries (device_ms = the whole step) or NULL */
int  llb_batch_register(llb_batch *b, float *T, llb_stats *stats);
int  llb_batch_register_async(llb_batch *b, const float *T);
int  llb_batch_result(llb_batch *b, float *T, llb_stats *stats);
/* which: 0 cornerLastDS, 1 surfLastDS, 2 outlierLastDS, 3 surfTotalLastDS of the slot's last step */
int  llb_batch_scan_get_ds(llb_batch *b, int slot, int which, llb_point *out, int capacity, int *n);
int  llb_batch_get_degeneracy(llb_batch *b, int slot, int *is_degenerate);
/* ---- device-resident key-frame stores for the slots of a batch (the llb_ctx calls llb_keyframe_add / llb_map_assemble
 * with a leading slot index).  llb_batch_enable_keyframes sizes the per-slot arenas and the scratch of the batched
 * map voxel filters: raw local maps of up to max_raw_map_points points each, up to max_keyframes key-frames per
 * assembled map.  llb_batch_keyframe_add stores the DS clouds of the slot's LAST completed step (MO:1443-1453);
 * llb_batch_map_assemble (MO:1033-1064) takes effect in the NEXT step, where the raw maps of all requesting slots are
 * assembled by one launch and voxel-filtered by one set of 18 launches; a slot's map then stays as it is until the
 * next llb_batch_map_assemble / llb_batch_map_set_ds. */
int  llb_batch_enable_keyframes(llb_batch *b, int max_raw_map_points, int max_keyframes);
int  llb_batch_keyframe_add(llb_batch *b, int slot, int *id);
int  llb_batch_keyframe_count(llb_batch *b, int slot, int *n);
int  llb_batch_map_assemble(llb_batch *b, int slot, const int *ids, const float *poses, int n);
/* the same for every slot with one call: slot s takes ids[offset[s] .. offset[s+1]) and the poses at the same positions
 * (offset: n_slots + 1 entries, offset[0] = 0); slots with an empty range keep their map */
int  llb_batch_map_assemble_all(llb_batch *b, const int *ids, const float *poses, const int *offset);
/* which: 0 / 1 raw corner / surf map, 2 / 3 DS corner / surf map of the slot's last assembled map (parity checks) */
int  llb_batch_map_get(llb_batch *b, int slot, int which, llb_point *out, int capacity, int *n);
/* ---- feature extraction of the slots: llb_features_init / _extract / _get with one sweep per slot (segs[slots],
 * counts[slots][4]); all slots share ONE set of five launches per step, every slot keeps its own per-point state */
int  llb_batch_features_init(llb_batch *b, int n_scan, int horizon_scan);
int  llb_batch_features_extract(llb_batch *b, const llb_segmented_cloud *segs, int *counts, float *device_ms);
int  llb_batch_features_get(llb_batch *b, int slot, int which, llb_point *out, int capacity, int *n);
/* ---- featureAssociation of the slots: llb_odom_set_last + llb_odom_set_features + llb_odom_optimize with a leading slot
 * index; updateTransformation (FA:1666-1695) of ALL slots is one launch (one persistent CTA per slot).  T: n_slots x 6
 * transformCur in/out; stats arrays of n_slots entries or NULL.  isDegenerate / matP and the neighbour indices kept
 * between iterations persist per slot, as in the reference. */
int  llb_batch_odom_set(llb_batch *b, int slot, const llb_point *corner_last, int ncl, const llb_point *surf_last, int nsl,
                        const llb_point *corner_sharp, int nsharp, const llb_point *surf_flat, int nflat);
int  llb_batch_odom_optimize(llb_batch *b, float *T, llb_stats *stats_surf, llb_stats *stats_corner);
/* per-stage CUDA-event times of the last step when enabled: ms[6] = {host-cloud unpack, downsampleCurrentScan,
 * index build, kNN kernels, fit kernels, LM-step + prepare + collect kernels}; geometry[4] = {kNN CTAs per slot,
 * fit CTAs per slot, index-build CTAs per map, query capacity per slot} */
int  llb_batch_set_profile(llb_batch *b, int on);
int  llb_batch_get_profile(llb_batch *b, float ms[6], int geometry[4]);

#ifdef __cplusplus
}
#endif
#endif /* LLB200_H_ */
