/*
 * llo.h — CPU ORACLE for the LeGO-LOAM scan-matching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm (featureAssociation.cpp = FA, mapOptmization.cpp = MO, utility.h = UT
 * under /root/reference/LeGO-LOAM) plus restatements of the un-vendored third
 * party primitives it calls (pcl::VoxelGrid, pcl::KdTreeFLANN, cv::solve,
 * cv::eigen, cv::Mat::inv, cv::gemm).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (lego_loam_b200/) never links or calls anything in this directory.
 *
 * Parity pinning status:
 *   - cv::solve(DECOMP_QR), cv::eigen, cv::invert, cv::gemm restatements are
 *     pinned bit-for-bit against the OpenCV 4.13 wheel present in the build
 *     container (tests/test_oracle_linalg.py, fixtures in tests/golden/).
 *   - pcl::VoxelGrid / pcl::KdTreeFLANN are NOT available anywhere offline:
 *     restated from the published algorithm ("parity unpinned" for those two,
 *     exact-kNN is cross-checked against scipy.cKDTree and cv2.flann).
 *   - control flow / thresholds / Jacobians are pinned against the UNMODIFIED
 *     reference sources compiled against shim headers (oracle/_ref).
 *   - feature extraction (llo_features.c) incl. its restatement of libstdc++'s
 *     std::sort: pinned bit-for-bit against oracle/_ref and against libstdc++
 *     itself (tests/test_oracle_features.py), golden vectors committed.
 *
 * Numerics contract: IEEE float32 with the double-promoted sub-expressions the
 * C++ reference has.  Build with -O2 -ffp-contract=off
 * and no -march=native so that no FMA contraction happens.
 */
#ifndef LLO_H_
#define LLO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* compact point: x, y, z, intensity (the 4 meaningful floats of pcl::PointXYZI) */
typedef struct { float x, y, z, intensity; } llo_point;

/* sin/cos of float arguments: mode 0 (default) = the host libm's sinf/cosf, i.e. what the
 * reference compiled on this machine calls (utility.h pulls the std:: float overloads in);
 * mode 1 = (float)sin((double)x), the correctly-rounded value, which is what the CUDA path
 * computes.  glibc's sinf/cosf are within 1 ulp but not always correctly rounded, so mode 1
 * lets the tests separate "libm flavour" from real arithmetic differences. */
void llo_set_trig_mode(int mode);
float llo_sinf(float x);
float llo_cosf(float x);

/* ---------------- third-party primitive restatements ---------------- */

/* cv::eigen on a symmetric CV_32F n x n matrix (row-major, n <= 8): cyclic
 * Jacobi with max-pivot bookkeeping, eigenvalues descending in W[n],
 * eigenvectors as ROWS of V[n*n].  A is destroyed. */
int llo_cv_eigen_f32(int n, float *A, float *W, float *V);

/* cv::solve(A, b, x, DECOMP_QR) for CV_32F, A m x n (m >= n, row-major), b m x 1.
 * Householder QR as hal::QR32f; on a singular system x is set to 0 and 0 is
 * returned (cv::solve's behaviour), else 1. */
int llo_cv_solve_qr_f32(int m, int n, const float *A, const float *b, float *x);

/* cv::Mat::inv() (DECOMP_LU) for CV_32F n x n: closed form in double for n<=3,
 * LU with partial pivoting in float for n>3.  Returns 0 when singular (dst=0). */
int llo_cv_inv_f32(int n, const float *A, float *Ainv);

/* cv::gemm for small CV_32F products D(m x n) = A(m x k) * B(k x n): every dot
 * product accumulated sequentially in double, rounded once to float. */
void llo_cv_gemm_f32(int m, int k, int n, const float *A, const float *B, float *D);

/* pcl::VoxelGrid<PointXYZI>::filter with setLeafSize(leaf,leaf,leaf) and all
 * defaults (downsample_all_data = true, min_points_per_voxel = 0).  out must
 * have room for n points.  Returns the number of output points; *overflow is
 * set to 1 when the int32 voxel index would overflow and the input is passed
 * through unchanged (PCL prints a warning and copies input to output).
 * Points inside a voxel are summed in ascending input-index order (stable). */
int llo_voxel_grid(const llo_point *in, int n, float leaf, llo_point *out, int *overflow);
/* the same on pcl::PointXYZI-layout clouds (8 floats per point: x y z _ intensity _ _ _), written in that layout (the
 * untouched floats of an output point keep what they held); out32 must have room for n points and must not alias in32 */
int llo_voxel_grid_pcl(const float *in32, int n, float leaf, float *out32, int *overflow);

/* pcl::KdTreeFLANN<PointXYZI>: exact k-NN, squared L2 in float
 * (flann::L2_Simple), results ascending; ties broken by smaller index. */
typedef struct llo_kdtree llo_kdtree;
llo_kdtree *llo_kdtree_build(const llo_point *pts, int n);
void llo_kdtree_free(llo_kdtree *t);
/* returns number found (= min(k, n)) */
int llo_kdtree_knn(const llo_kdtree *t, const float *q, int k, int *idx, float *d2);
/* brute-force definition of the same query (used to validate the tree) */
int llo_knn_bruteforce(const llo_point *pts, int n, const float *q, int k, int *idx, float *d2);

/* ---------------- loop closure (llo_loop.c; PARITY UNPINNED: PCL / Eigen absent) ----------------
 * pcl::IterativeClosestPoint<PointXYZI,PointXYZI>::align + getFitnessScore as performLoopClosure configures them
 * (MO:892-904), restated from PCL 1.8.  T_final: final_transformation_ (row-major 4x4); state: 0 too few
 * correspondences (hasConverged() false), 1 iterations, 2 transform, 3 abs MSE, 4 rel MSE. */
int llo_icp_align(const llo_point *src, int ns, const llo_point *tgt, int nt, int max_iterations, double max_corr_dist,
                  double transformation_epsilon, double euclidean_fitness_epsilon, float T_final[16], int *converged,
                  int *iterations, int *state, double *fitness);
/* one step for per-function parity: correspondences of the CURRENT source cloud (exact 1-NN, kept when d^2 <= max_d2),
 * the sums n (return value), sum p, sum q, sum q p^T, and the mean squared distance; nn_idx (may be NULL): target index or -1 */
int llo_icp_correspondence_sums(const llo_point *cur, int ns, const llo_kdtree *tree, const llo_point *tgt, double max_d2,
                                double sp[3], double sq[3], double sqp[9], double *mse, int *nn_idx);
/* pcl::umeyama (no scaling) from those sums -> row-major 4x4 */
void llo_umeyama_from_sums(double n, const double sp[3], const double sq[3], const double sqp[9], float Rt[16]);

/* ---------------- mapOptimization hot path (MO:498-527, MO:1067-1350) ---------------- */

typedef struct llo_mapopt llo_mapopt;

llo_mapopt *llo_mapopt_create(void);
void llo_mapopt_destroy(llo_mapopt *m);

/* members written by the caller side (extractSurroundingKeyFrames tail MO:1057-1064
 * hands over the DS maps; handlers MO:608-627 hand over the scan) */
void llo_mapopt_set_map_ds(llo_mapopt *m, const llo_point *corner_ds, int mc,
                           const llo_point *surf_ds, int ms);
/* raw map -> DS map (MO:1057-1064) */
void llo_mapopt_set_map_raw(llo_mapopt *m, const llo_point *corner, int rc,
                            const llo_point *surf, int rs);
void llo_mapopt_set_scan(llo_mapopt *m, const llo_point *corner_last, int nc,
                         const llo_point *surf_last, int ns,
                         const llo_point *outlier_last, int no);
void llo_mapopt_set_pose(llo_mapopt *m, const float tobe_mapped[6]);
void llo_mapopt_get_pose(const llo_mapopt *m, float tobe_mapped[6]);
void llo_mapopt_set_transform_sum(llo_mapopt *m, const float sum[6]);
void llo_mapopt_get_bef_aft(const llo_mapopt *m, float bef[6], float aft[6]);
/* isDegenerate / matP persist across registrations */
void llo_mapopt_get_degenerate(const llo_mapopt *m, int *is_degenerate, float matP[36]);

void llo_mapopt_downsampleCurrentScan(llo_mapopt *m);                 /* MO:1067 */
void llo_mapopt_build_kdtrees(llo_mapopt *m);                          /* MO:1333-1334 */
void llo_mapopt_clear_correspondences(llo_mapopt *m);                  /* MO:1338-1339 */
void llo_mapopt_cornerOptimization(llo_mapopt *m, int iterCount);      /* MO:1093 */
void llo_mapopt_surfOptimization(llo_mapopt *m, int iterCount);        /* MO:1176 */
int  llo_mapopt_LMOptimization(llo_mapopt *m, int iterCount);          /* MO:1229 */
/* returns the number of LM iterations executed (0 when the guard MO:1331 fails) */
int  llo_mapopt_scan2MapOptimization(llo_mapopt *m);                   /* MO:1329 */

/* read-back of members */
int llo_mapopt_get_scan_ds(const llo_mapopt *m, int which /*0 corner,1 surf,2 outlier,3 surfTotal*/,
                           llo_point *out, int cap);
int llo_mapopt_get_map_ds(const llo_mapopt *m, int which /*0 corner,1 surf*/, llo_point *out, int cap);
int llo_mapopt_get_correspondences(const llo_mapopt *m, llo_point *ori, llo_point *coeff, int cap);
/* per-query diagnostics of the last corner/surfOptimization call: 5 neighbour
 * indices (or -1) per query, in query order */
int llo_mapopt_get_knn(const llo_mapopt *m, int which /*0 corner,1 surf*/, int *idx5, float *d2_5, int cap);
/* last AtA (36) / AtB (6) / matX (6) of LMOptimization */
void llo_mapopt_get_normal_eq(const llo_mapopt *m, float AtA[36], float AtB[6], float X[6]);

/* ---------------- featureAssociation hot path (FA:860-883, FA:1044-1478, FA:1666-1695) ------- */

typedef struct llo_featassoc llo_featassoc;

llo_featassoc *llo_featassoc_create(void);
void llo_featassoc_destroy(llo_featassoc *f);
/* laserCloudCornerLast / laserCloudSurfLast + kd-tree rebuild (FA:1615-1619, FA:1774-1788).
 * The trees are rebuilt only if nc > 10 && ns > 100 (FA:1785, quirk C20) unless force != 0. */
void llo_featassoc_set_last(llo_featassoc *f, const llo_point *corner_last, int nc,
                            const llo_point *surf_last, int ns, int force);
void llo_featassoc_set_features(llo_featassoc *f, const llo_point *corner_sharp, int nsharp,
                                const llo_point *surf_flat, int nflat);
void llo_featassoc_set_transform(llo_featassoc *f, const float cur[6]);
void llo_featassoc_get_transform(const llo_featassoc *f, float cur[6]);
void llo_featassoc_get_degenerate(const llo_featassoc *f, int *is_degenerate, float matP[9]);
void llo_featassoc_clear_correspondences(llo_featassoc *f);
void llo_featassoc_findCorrespondingCornerFeatures(llo_featassoc *f, int iterCount); /* FA:1044 */
void llo_featassoc_findCorrespondingSurfFeatures(llo_featassoc *f, int iterCount);   /* FA:1155 */
int  llo_featassoc_calculateTransformationSurf(llo_featassoc *f, int iterCount);     /* FA:1270 */
int  llo_featassoc_calculateTransformationCorner(llo_featassoc *f, int iterCount);   /* FA:1379 */
/* returns iterations executed: low 16 bits surf loop, high 16 bits corner loop */
int  llo_featassoc_updateTransformation(llo_featassoc *f);                           /* FA:1666 */
int  llo_featassoc_get_correspondences(const llo_featassoc *f, llo_point *ori, llo_point *coeff, int cap);
/* pointSearch{Corner,Surf}Ind{1,2,3} (stored as float in the reference, C3) */
int  llo_featassoc_get_search_ind(const llo_featassoc *f, int which /*0 corner,1 surf*/,
                                  float *ind1, float *ind2, float *ind3, int cap);

/* ---------------- imageProjection (llo_projection.c) ---------------- */
typedef struct llo_projection llo_projection;
llo_projection *llo_projection_create(int n_scan, int horizon, float ang_res_x, float ang_res_y, int ground_scan_ind);
void llo_projection_destroy(llo_projection *p);
/* cloudHandler IP:181-197 without publishing: raw sweep (lidar frame, firing order, ring channel) -> range / ground / label
 * images, segmented cloud + cloud_info, outlier cloud */
void llo_projection_process(llo_projection *p, const llo_point *cloud, const uint16_t *ring, int n);
int llo_projection_get_cloud(const llo_projection *p, int which /* 0 segmented, 1 outlier */, llo_point *out, int cap);
void llo_projection_get_info(const llo_projection *p, int *start_ring, int *end_ring, float ori[3], uint8_t *ground,
                             uint32_t *col, float *range, int n);
void llo_projection_get_images(const llo_projection *p, float *range_mat, int8_t *ground_mat, int32_t *label_mat);

/* ---------------- feature extraction (llo_features.c) ---------------- */
typedef struct llo_features llo_features;
llo_features *llo_features_create(int n_scan, int horizon);
void llo_features_destroy(llo_features *f);
/* adjustDistortion (no IMU), calculateSmoothness, markOccludedPoints, extractFeatures (FA:491-784) on one segmented
 * sweep; cloud is adjusted in place; ground/col/range hold n_scan*horizon entries (zero beyond n).
 * out[0..3] = cornerPointsSharp, cornerPointsLessSharp, surfPointsFlat, surfPointsLessFlat (room for n each). */
void llo_features_extract(llo_features *f, llo_point *cloud, int n, const int *start_ring, const int *end_ring,
                          float start_ori, float end_ori, float ori_diff,
                          const uint8_t *ground, const uint32_t *col, const float *range,
                          llo_point *const out[4], int n_out[4]);
void llo_features_get_state(const llo_features *f, int n, float *curv, int *picked, int *label);
void llo_adjust_distortion(llo_point *cloud, int n, float start_ori, float end_ori, float ori_diff, float scan_period);
/* TransformToEnd FA:885-953 (no IMU messages received) on every point, in place */
void llo_transform_to_end(const float T[6], llo_point *cloud, int n);
/* libstdc++ std::sort of (value, ind) records compared by value only (FA:57-61, FA:699); depth_limit < 0 = std::sort's own */
void llo_std_sort_by_value(float *value, uint32_t *ind, int n, int depth_limit);

#ifdef __cplusplus
}
#endif
#endif /* LLO_H_ */
