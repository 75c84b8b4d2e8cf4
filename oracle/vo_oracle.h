/*
 * vo_oracle.h — CPU restatement (plain C, single-threaded, FP32 with every operation
 * individually rounded, no FMA contraction) of the reference's hot path.
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may link or call it; the product (libvo_b200.so) never
 * does.  Each function cites the reference lines it follows (paths relative to the reference
 * checkout).  Pinning status: see oracle/README.md — the reference ships no golden vectors;
 * this restatement is pinned against the reference's OWN sources compiled in this container
 * (oracle/_ref, built against third_party/mini_eigen because Eigen3 is not installed) and
 * against the bundled dataset's ground-truth landmark ids.
 */
#ifndef VO_ORACLE_H
#define VO_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Same POD layouts as include/vo_b200.h so tests can pass one buffer to both sides. */
typedef struct oracle_camera {
  int32_t rows, cols, z_near, z_far;
  float K[9];  /* column-major 3x3 */
  float T[16]; /* column-major 4x4 world-in-camera */
} oracle_camera;

typedef struct oracle_picp_state {
  float T[16];
  float H[36];
  float b[6];
  float chi_inliers, chi_outliers;
  int32_t num_inliers;
  int32_t rounds_done;
  int32_t last_ok;
} oracle_picp_state;

/* (p-q).tail(dim).squaredNorm() in Eigen's SSE2 linear-vectorised reduction order
 * (brute_force_search.h:14,34; order derivation in DESIGN.md / SURVEY.md Appendix B.1). */
float oracle_sqdist(const float* p, const float* q, int dim);

/* bruteForceBestMatch, brute_force_search.h:22-41, for a batch of queries. */
void oracle_nn_best_match(const float* map, int64_t n_rows, int row_stride, int skip_cols,
                          const float* queries, int64_t n_queries, int query_stride, float norm,
                          int32_t* best_idx, float* best_d2);

/* bruteForceSearch, brute_force_search.h:3-20. */
void oracle_nn_radius_search(const float* map, int64_t n_rows, int row_stride, int skip_cols,
                             const float* queries, int64_t n_queries, int query_stride,
                             float norm, int32_t* counts, int32_t* idx_out,
                             int32_t max_per_query);

/* Camera::projectPoint, camera.h:25-37. returns 1 if inside. */
int oracle_project_point(const oracle_camera* cam, const float wp[3], float out[2]);
/* Camera::projectPoints, camera.cpp:16-37. */
void oracle_project_points(const oracle_camera* cam, const float* world, int64_t n,
                           int keep_indices, float* out_image, int64_t* n_out,
                           int64_t* n_inside);

/* PICPSolver::oneRound (picp_solver.cpp:98-112) = linearize (:55-96) + damping + LDLT solve +
 * v2tEuler left-multiply (utils.h:64-78).  `st->T` is the pose, updated in place.
 * returns 1 (true) / 0 (too few inliers). */
int oracle_picp_one_round(oracle_picp_state* st, const oracle_camera* cam_params,
                          const float* world, const float* image, const int32_t* pairs,
                          int64_t n_pairs, int keep_outliers, float kernel_threshold,
                          float damping, int32_t min_num_inliers);

/* Same round in float64 with float64 accumulators ("truth" for large-N comparisons; not the
 * reference's arithmetic).  T is float64 column-major, updated in place. */
int oracle_picp_one_round_f64(double T[16], double H[36], double b[6], double stats[3],
                              const oracle_camera* cam_params, const float* world,
                              const float* image, const int32_t* pairs, int64_t n_pairs,
                              int keep_outliers, double kernel_threshold, double damping);

/* triangulate_points, utils.cpp:51-134 (all three overloads share this body).
 * out_* may be NULL except out_points. returns n_success. */
int64_t oracle_triangulate_points(const float K[9], const float X[16], const int32_t* corr,
                                  int64_t n_corr, const float* p1, const float* p2,
                                  const float* app2, float* out_points, int32_t* out_corr_new,
                                  float* out_app, int32_t* out_src);

/* helpers exposed for tests */
void oracle_ldlt_solve(int n, const float* A_colmajor, const float* rhs, float* x);
void oracle_v2t_euler(const float v[6], float T[16]);

#ifdef __cplusplus
}
#endif
#endif
