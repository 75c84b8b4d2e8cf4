/*
 * vo_b200.h — C ABI of libvo_b200.so, the sm_100a (NVIDIA B200) implementation of the
 * per-frame data-parallel hot path of lucanunz/Visual-odometry.
 *
 * Every entry point below replaces one piece of the reference's C++ surface (cited as
 * file:line relative to the reference checkout).  The C++ drop-in headers under
 * visual-odometry_b200/host/include/ (same file names as the reference's include/) are thin
 * marshalling layers over this ABI; bench.py and tests/ bind it with ctypes.
 *
 * Conventions
 *   - plain pointers + sizes only; no C++/torch types cross this boundary.
 *   - "_host" style entry points (no suffix) take HOST pointers in the reference's native
 *     strides (what std::vector<Eigen::...>::data() gives) and do their own H2D/D2H copies.
 *   - "_device" entry points take DEVICE pointers, enqueue on the handle's stream and
 *     return without synchronising (async).
 *   - all functions return 0 (VO_OK) or a negative vo_status; vo_last_error() gives text.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     VO_ERR_CUDA.
 *   - matrices are column-major (Eigen's default): K[9] = 3x3, T[16] = 4x4 homogeneous.
 *   - index pairs are int32[2] exactly as std::pair<int,int> is laid out (first, second).
 */
#ifndef VO_B200_H
#define VO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VO_B200_ABI_VERSION 1

typedef enum vo_status {
  VO_OK = 0,
  VO_ERR_ARG = -1,         /* null pointer, negative size, bad stride ...            */
  VO_ERR_CUDA = -2,        /* a CUDA runtime call failed (see vo_last_error)         */
  VO_ERR_UNSUPPORTED = -3, /* e.g. appearance dimension outside 1..VO_NN_MAX_DIM     */
  VO_ERR_STATE = -4        /* call order violated (query before set_map, ...)        */
} vo_status;

#define VO_NN_MAX_DIM 32

/* ---- library ----------------------------------------------------------------------- */
int vo_abi_version(void);
/* thread-local text of the last failure on the calling thread ("" if none)             */
const char* vo_last_error(void);
/* number of CUDA devices visible; <0 on error                                           */
int vo_device_count(void);
/* total kernel launches issued by this library in this process (for bench gpu_launches) */
int64_t vo_launch_count(void);
/* measured FP32 FFMA throughput (TFLOP/s) of a register-resident FMA loop on `device`;
 * used only as a roofline denominator in bench.py                                        */
int vo_measure_ffma_peak(int device, double* tflops_out);
/* the same with packed fma.rn.f32x2 (SASS FFMA2, two FMAs per lane per instruction)          */
int vo_measure_ffma2_peak(int device, double* tflops_out);

/* ==== (1) appearance nearest neighbour =================================================
 * replaces bruteForceBestMatch / bruteForceSearch   include/brute_force_search.h:3-41
 * and answers the queries of TreeNode_::bestMatchFull include/eigen_kdtree.h:90-115
 *
 * A map row / query is `row_stride` floats of which the first `skip_cols` are carried but
 * ignored (the reference's Vector11f = [float(id) | 10 appearance floats], defs.h:7,
 * vo_complete.cpp:22).  dim = row_stride - skip_cols.  Distance is the squared L2 norm over
 * the dim trailing floats, evaluated in the reference's FP32 order;
 * a row matches when d2 < norm*norm (strict); the best match is the strict minimum, the
 * lowest row index winning ties; -1 when no row matches.
 *
 * Two filters feed the exact re-rank (same answers, chosen per call): the FP32 FFMA2
 * partial-distance filter (csrc/nn.cu; small maps, small batches, data whose norms do not fit
 * f16) and the tcgen05 f16 tensor-core filter (csrc/nn_tc.cu; maps >= 32768 rows and batches
 * >= 2048 queries).  The launch log (queries per thread == 0) marks the tensor-core filter.
 * VO_NN_FORCE_PATH=ffma|tc in the environment at vo_nn_create pins one of them (tests, bench).
 */
typedef struct vo_nn_s* vo_nn_t;

int vo_nn_create(vo_nn_t* out, int device);
int vo_nn_destroy(vo_nn_t h);
/* use an existing cudaStream_t (passed as void*) instead of the handle's own stream      */
int vo_nn_set_stream(vo_nn_t h, void* cuda_stream);
int vo_nn_synchronize(vo_nn_t h);

/* upload (host) or adopt (device) the map rows; both re-pack the rows on the device into
 * the kernel's tile layout.  The caller's buffer is not referenced after the call returns
 * (host variant) / after the enqueued work completes (device variant).                    */
int vo_nn_set_map(vo_nn_t h, const float* rows_host, int64_t n_rows, int row_stride,
                  int skip_cols);
int vo_nn_set_map_device(vo_nn_t h, const float* rows_dev, int64_t n_rows, int row_stride,
                         int skip_cols);

/* bruteForceBestMatch for a batch of queries (same row layout as the map).
 * best_idx[q] = winning row index or -1;  best_d2 (nullable) = its squared distance
 * (undefined where best_idx == -1).                                                       */
int vo_nn_best_match(vo_nn_t h, const float* queries_host, int64_t n_queries,
                     int query_stride, float norm, int32_t* best_idx_host,
                     float* best_d2_host);
int vo_nn_best_match_device(vo_nn_t h, const float* queries_dev, int64_t n_queries,
                            int query_stride, float norm, int32_t* best_idx_dev,
                            float* best_d2_dev);

/* Introspection for tests and profiles: the filter-kernel launches the LAST best_match call on this
 * handle issued, four int32 per launch = (queries per thread, threads per block, query tiles, map
 * splits).  *n_launches = how many there were; at most `capacity` of them are written to out.    */
int vo_nn_last_launches(vo_nn_t h, int32_t* out, int capacity, int* n_launches);
/* (query, 128-row block) pairs the tensor-core filter of the last best_match call handed to the
 * exact re-rank; -1 when that call ran the FP32 filter.  Synchronises.                            */
int vo_nn_last_rescans(vo_nn_t h, int64_t* n_rescans);

/* bruteForceSearch for a batch of queries: counts[q] = number of rows with d2 < norm^2.
 * If idx_out != NULL, the matching row indices of query q are written in ascending row
 * order to idx_out[q*max_per_query ...], at most max_per_query of them.                    */
int vo_nn_radius_search(vo_nn_t h, const float* queries_host, int64_t n_queries,
                        int query_stride, float norm, int32_t* counts_host,
                        int32_t* idx_out_host, int32_t max_per_query);

/* ==== (1b) the same on several GPUs of one node ===========================================
 * BASELINE config 4 / SURVEY.md 8e: the map is replicated on every GPU, a query batch is cut into
 * contiguous blocks (GPU g answers [g*Q/G, (g+1)*Q/G)), and the int32 match indices are gathered
 * with ONE ncclAllGather over NVLink — the only collective; no data-path exchange.  One process
 * drives all GPUs (ncclCommInitAll over devices 0..n-1; NCCL is loaded at run time).  The reference
 * has no counterpart: it answers one query at a time on one thread (src/apps/vo_complete.cpp:12-49).
 */
typedef struct vo_comm_s* vo_comm_t;

int vo_comm_init_all(vo_comm_t* out, int n_gpus);
int vo_comm_destroy(vo_comm_t c);
int vo_comm_size(vo_comm_t c);
/* vo_nn_set_map on GPU 0, then the re-packed map is broadcast to the other GPUs over NVLink      */
int vo_nn_set_map_replicated(vo_comm_t c, const float* rows_host, int64_t n_rows, int row_stride,
                             int skip_cols);
/* vo_nn_best_match with the queries sharded over the communicator's GPUs                         */
int vo_nn_best_match_sharded(vo_comm_t c, const float* queries_host, int64_t n_queries,
                             int query_stride, float norm, int32_t* best_idx_host);

/* ==== (2) projective ICP ===============================================================
 * replaces PICPSolver::{init,oneRound,linearize,errorAndJacobian}  src/picp_solver.cpp:16-112
 * and Camera::projectPoint                                        include/camera.h:25-37
 */
typedef struct vo_camera {
  int32_t rows, cols;     /* image size                      camera.h:57-58                 */
  int32_t z_near, z_far;  /* integer depth range             camera.h:59-60                 */
  float K[9];             /* camera matrix, column-major     camera.h:61                    */
  float T[16];            /* world-in-camera pose, col-major camera.h:62                    */
} vo_camera;

typedef struct vo_picp_state {
  float T[16];            /* current world-in-camera pose (column-major 4x4)                */
  float H[36];            /* last linearisation, column-major, damping already added        */
  float b[6];
  float chi_inliers, chi_outliers;
  int32_t num_inliers;
  int32_t rounds_done;    /* rounds executed since init                                     */
  int32_t last_ok;        /* 0 if the last round was skipped (too few inliers)              */
} vo_picp_state;

typedef struct vo_picp_s* vo_picp_t;

int vo_picp_create(vo_picp_t* out, int device);
int vo_picp_destroy(vo_picp_t h);
int vo_picp_set_stream(vo_picp_t h, void* cuda_stream);
int vo_picp_synchronize(vo_picp_t h);
/* kernel_threshold / damping / min_num_inliers  (picp_solver.cpp:10-13; setKernelThreshold
 * picp_solver.h:35)                                                                        */
int vo_picp_set_params(vo_picp_t h, float kernel_threshold, float damping,
                       int32_t min_num_inliers);

/* PICPSolver::init: copies the camera and uploads the two point sets
 * (world: 3 floats/point, image: 2 floats/point, tightly packed as Eigen vectors are).     */
int vo_picp_init(vo_picp_t h, const vo_camera* cam, const float* world_host,
                 int64_t n_world, const float* image_host, int64_t n_image);
int vo_picp_init_device(vo_picp_t h, const vo_camera* cam, const float* world_dev,
                        int64_t n_world, const float* image_dev, int64_t n_image);

/* upload correspondences: int32 pairs (first = image/measurement index, second = world
 * index; picp_solver.cpp:66-71).                                                           */
int vo_picp_set_correspondences(vo_picp_t h, const int32_t* pairs_host, int64_t n_pairs);
int vo_picp_set_correspondences_device(vo_picp_t h, const int32_t* pairs_dev, int64_t n_pairs);

/* `rounds` x oneRound on the uploaded correspondences, all on the device, no host
 * round-trip between rounds; asynchronous.                                                 */
int vo_picp_compute(vo_picp_t h, int keep_outliers, int rounds);
/* frame-pipeline variant of vo_picp_compute (<= 65536 correspondences, resident kernel only):
 * n_pairs_dev (nullable) is a DEVICE int32 holding the actual number of uploaded pairs (the count
 * passed to set_correspondences_device is then only an upper bound); pre_transform (nullable) is
 * a column-major 4x4 isometry applied to every world point while it is gathered — the
 * `X_curr * triangulated_pc` of vo_complete.cpp:154 without a pass over the cloud.              */
int vo_picp_compute_ex(vo_picp_t h, int keep_outliers, int rounds, const int32_t* n_pairs_dev,
                       const float* pre_transform);
/* convenience == set_correspondences + compute                                             */
int vo_picp_one_round(vo_picp_t h, const int32_t* pairs_host, int64_t n_pairs,
                      int keep_outliers);
/* synchronises and copies the solver state to the host                                     */
int vo_picp_get_state(vo_picp_t h, vo_picp_state* out);

/* ==== (3) triangulation ================================================================
 * replaces triangulate_point / triangulate_points (3 overloads)   src/utils.cpp:36-134
 *
 * corr: int32 pairs (first -> index into p1, second -> index into p2).
 * Outputs are the order-preserving compaction of the successes:
 *   out_points[k]   (3 floats)       the k-th triangulated point           (utils.cpp:70,98)
 *   out_corr_new[k] = (second, k)    nullable                              (utils.cpp:97)
 *   out_app[k]      (10 floats) = app2[second]  nullable, with app2        (utils.cpp:127)
 *   out_src[k]      = position of the k-th success in corr   nullable (extension, for tests)
 * returns the number of successes in *n_success.
 */
int vo_triangulate(int device, const float K[9], const float X[16], const int32_t* corr_host,
                   int64_t n_corr, const float* p1_host, int64_t n_p1, const float* p2_host,
                   int64_t n_p2, const float* app2_host, float* out_points_host,
                   int32_t* out_corr_new_host, float* out_app_host, int32_t* out_src_host,
                   int64_t* n_success);
/* device-resident variant: all pointers are device pointers except n_success_dev which is a
 * device int64; asynchronous on `cuda_stream`; `workspace_dev` must hold
 * vo_triangulate_workspace_bytes(n_corr) bytes.                                            */
int64_t vo_triangulate_workspace_bytes(int64_t n_corr);
int vo_triangulate_device(void* cuda_stream, const float K[9], const float X[16],
                          const int32_t* corr_dev, int64_t n_corr, const float* p1_dev,
                          const float* p2_dev, const float* app2_dev, float* out_points_dev,
                          int32_t* out_corr_new_dev, float* out_app_dev, int32_t* out_src_dev,
                          int64_t* n_success_dev, void* workspace_dev);

/* as vo_triangulate_device, with the number of correspondences read from DEVICE memory
 * (n_corr_dev, int32); n_corr_max bounds it and sizes the workspace.                        */
int vo_triangulate_device_ex(void* cuda_stream, const float K[9], const float X[16],
                             const int32_t* corr_dev, int64_t n_corr_max, const int32_t* n_corr_dev,
                             const float* p1_dev, const float* p2_dev, const float* app2_dev,
                             float* out_points_dev, int32_t* out_corr_new_dev, float* out_app_dev,
                             int64_t* n_success_dev, void* workspace_dev);

/* ==== (4) batch projection =============================================================
 * replaces Camera::projectPoints   src/camera.cpp:16-37
 * keep_indices != 0: out_image has n_points entries, invalid ones are (-1,-1);
 * keep_indices == 0: out_image is the order-preserving compaction of the valid ones.
 * *n_inside = number of points that project inside the image; *n_out = entries written.   */
int vo_project_points(int device, const vo_camera* cam, const float* world_host,
                      int64_t n_points, int keep_indices, float* out_image_host,
                      int64_t* n_out, int64_t* n_inside);

/* ==== (5) device-resident frame loop ====================================================
 * replaces the loop body of src/apps/vo_complete.cpp:150-178 — compute_correspondences_images
 * (:12-48), extract_correspondences_world (:51-66), operator*(Isometry3f, PointCloudVector)
 * (PointCloud.h:77-82), PICPSolver::init + 100 x oneRound, triangulate_points (point-cloud
 * overload) and PointCloudVector::update (PointCloud.h:52-66) — with the frames, the match
 * lists, the triangulated cloud and the map resident on the device.  Per frame only the new
 * measurements go up and the pose comes down.  Epipolar initialisation stays on the host:
 *   vo_pipe_first_frame(f0); vo_pipe_second_frame(f1, corr, &n);   // appearance matches of (f0,f1)
 *   X = estimate_transform(K, corr, ...)  (host);  vo_pipe_bootstrap(X);
 *   for every further frame: vo_pipe_step(...)
 * A frame is n measurements: points (2 floats each) and appearances (10 floats each).          */
typedef struct vo_pipe_s* vo_pipe_t;

typedef struct vo_pipe_result {
  float T[16];               /* pose of the previous camera in the current one (column-major)  */
  int64_t n_measurements;    /* of this frame                                                   */
  int64_t n_matches;         /* appearance matches with the previous frame                      */
  int64_t n_correspondences; /* matches that also have a triangulated point (PICP input)        */
  int64_t map_points;        /* map size after the PREVIOUS frame's merge                       */
  float chi_inliers;
  int32_t n_inliers;
  int32_t map_overflow;      /* != 0: max_map_points was reached, later points were dropped     */
} vo_pipe_result;

int vo_pipe_create(vo_pipe_t* out, int device, const vo_camera* cam, int64_t max_points_per_frame,
                   int64_t max_map_points);
int vo_pipe_destroy(vo_pipe_t h);
int vo_pipe_first_frame(vo_pipe_t h, const float* points_host, const float* app_host, int64_t n);
/* uploads the second frame and returns the (first, second) index pairs of the appearance
 * matches: *n_corr = how many there are (at most min(n0, n1)); the first min(*n_corr,
 * corr_capacity) pairs are written to corr_host (int32 pairs; nullable with capacity 0)        */
int vo_pipe_second_frame(vo_pipe_t h, const float* points_host, const float* app_host, int64_t n,
                         int32_t* corr_host, int64_t corr_capacity, int64_t* n_corr);
/* X: pose of the first camera in the second (what estimate_transform returns); triangulates the
 * first pair, starts the map                                                                   */
int vo_pipe_bootstrap(vo_pipe_t h, const float X[16]);
/* one frame of the loop; synchronises once, to return the pose                                 */
int vo_pipe_step(vo_pipe_t h, const float* points_host, const float* app_host, int64_t n, int rounds,
                 float kernel_threshold, vo_pipe_result* out);
/* PointCloudVector::update (PointCloud.h:52-66) on the device map, for a host cloud moved by X:
 * a point whose appearance is already stored (float ==, first hit) replaces that position, every
 * other point is appended in order; appended points take part in the matching of later ones.   */
int vo_pipe_merge_cloud(vo_pipe_t h, const float* points_host, const float* app_host, int64_t n,
                        const float X[16]);
/* the global map in `history` coordinates (vo_complete.cpp applies cam_transform afterwards)   */
int vo_pipe_get_map(vo_pipe_t h, float* points_host, float* app_host, int64_t capacity, int64_t* n);

#ifdef __cplusplus
}
#endif
#endif /* VO_B200_H */
