// ref_harness.cpp — thin extern "C" shim around the REFERENCE'S OWN code (compiled unchanged
// from /root/reference against third_party/mini_eigen; see build_ref.sh).  TEST INFRASTRUCTURE:
// it exists so the C restatement in vo_oracle.c can be validated against the reference's real
// control flow (index roles, thresholds, compaction order, kd-tree descent).  Nothing in the
// product links it.
#include <cstdint>
#include <cstring>
#include <vector>

#include "defs.h"
#include "camera.h"
#include "picp_solver.h"
#include "utils.h"
#include "eigen_kdtree.h"  // pulls split.h, eigen_covariance.h, brute_force_search.h

namespace {
Eigen::Matrix3f mat3(const float* colmajor) {
  Eigen::Matrix3f m;
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) m(i, j) = colmajor[j * 3 + i];
  return m;
}
Eigen::Isometry3f iso(const float* colmajor16) {
  Eigen::Isometry3f x = Eigen::Isometry3f::Identity();
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 3; ++i) x(i, j) = colmajor16[j * 4 + i];
  return x;
}
void iso_out(const Eigen::Isometry3f& x, float* colmajor16) {
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) colmajor16[j * 4 + i] = (i == 3) ? (j == 3 ? 1.f : 0.f) : x(i, j);
}
Vector11fVector rows11(const float* rows, int64_t n) {
  Vector11fVector v((size_t)n);
  for (int64_t r = 0; r < n; ++r)
    for (int k = 0; k < 11; ++k) v[(size_t)r](k) = rows[r * 11 + k];
  return v;
}
Vector3fVector vec3(const float* p, int64_t n) {
  Vector3fVector v((size_t)n);
  for (int64_t i = 0; i < n; ++i) v[(size_t)i] = Eigen::Vector3f(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
  return v;
}
Vector2fVector vec2(const float* p, int64_t n) {
  Vector2fVector v((size_t)n);
  for (int64_t i = 0; i < n; ++i) v[(size_t)i] = Eigen::Vector2f(p[2 * i], p[2 * i + 1]);
  return v;
}
IntPairVector pairs(const int32_t* p, int64_t n) {
  IntPairVector v((size_t)n);
  for (int64_t i = 0; i < n; ++i) v[(size_t)i] = IntPair(p[2 * i], p[2 * i + 1]);
  return v;
}
struct SolverAccess : public PICPSolver {  // the accumulators are protected members
  const Matrix6f& H() const { return _H; }
  const Vector6f& b() const { return _b; }
};
struct RefPicp {
  SolverAccess solver;
  Vector3fVector world;
  Vector2fVector image;
};
}  // namespace

extern "C" {

// bruteForceBestMatch over Vector11f rows (include/brute_force_search.h:22-41)
void ref_nn_best_match(const float* map, int64_t n_rows, const float* queries, int64_t n_queries,
                       float norm, int32_t* best_idx, float* best_d2) {
  Vector11fVector m = rows11(map, n_rows), q = rows11(queries, n_queries);
  for (int64_t i = 0; i < n_queries; ++i) {
    Vector11f* hit = bruteForceBestMatch(m.begin(), m.end(), q[(size_t)i], norm);
    best_idx[i] = hit ? (int32_t)(hit - &m[0]) : -1;
    if (best_d2) best_d2[i] = hit ? ((*hit) - q[(size_t)i]).tail(10).squaredNorm() : norm * norm;
  }
}

// The same template instantiated over the caller's buffer IN PLACE (no copy of the map): a
// minimal iterator over tightly packed Vector11f rows.  Read-only, so several host threads may
// search the same map concurrently (bench.py's CPU arm).
namespace {
struct RowIterator {
  using value_type = Vector11f;
  Vector11f* p;
  Vector11f& operator*() const { return *p; }
  RowIterator& operator++() {
    ++p;
    return *this;
  }
  bool operator!=(const RowIterator& o) const { return p != o.p; }
};
static_assert(sizeof(Vector11f) == 11 * sizeof(float), "rows must be tightly packed");
}  // namespace
void ref_nn_best_match_inplace(const float* map, int64_t n_rows, const float* queries,
                               int64_t n_queries, float norm, int32_t* best_idx) {
  Vector11f* rows = reinterpret_cast<Vector11f*>(const_cast<float*>(map));
  const Vector11f* q = reinterpret_cast<const Vector11f*>(queries);
  for (int64_t i = 0; i < n_queries; ++i) {
    Vector11f* hit = bruteForceBestMatch(RowIterator{rows}, RowIterator{rows + n_rows}, q[i], norm);
    best_idx[i] = hit ? (int32_t)(hit - rows) : -1;
  }
}

// bruteForceSearch (include/brute_force_search.h:3-20)
void ref_nn_radius_search(const float* map, int64_t n_rows, const float* queries, int64_t n_queries,
                          float norm, int32_t* counts, int32_t* idx_out, int32_t max_per_query) {
  Vector11fVector m = rows11(map, n_rows), q = rows11(queries, n_queries);
  for (int64_t i = 0; i < n_queries; ++i) {
    std::vector<Vector11f*> ans;
    counts[i] = bruteForceSearch(ans, m.begin(), m.end(), q[(size_t)i], norm);
    for (size_t k = 0; k < ans.size() && (int32_t)k < max_per_query; ++k)
      idx_out[i * max_per_query + (int64_t)k] = (int32_t)(ans[k] - &m[0]);
  }
}

// TreeNode_ build + bestMatchFull / bestMatchFast (include/eigen_kdtree.h), used the way
// compute_correspondences_images does (src/apps/vo_complete.cpp:12-49): id in column 0.
void ref_kdtree_best_match(const float* map, int64_t n_rows, const float* queries,
                           int64_t n_queries, float norm, int max_points_in_leaf, int full,
                           int32_t* best_idx) {
  Vector11fVector m = rows11(map, n_rows), q = rows11(queries, n_queries);
  for (int64_t r = 0; r < n_rows; ++r) m[(size_t)r](0) = float(r);
  TreeNode_<Vector11fVector::iterator> tree(m.begin(), m.end(), max_points_in_leaf);
  for (int64_t i = 0; i < n_queries; ++i) {
    Vector11f* hit = full ? tree.bestMatchFull(q[(size_t)i], norm) : tree.bestMatchFast(q[(size_t)i], norm);
    best_idx[i] = hit ? (int32_t)(*hit)(0) : -1;
  }
}

// Camera::projectPoints (src/camera.cpp:16-37)
void ref_project_points(int rows, int cols, int z_near, int z_far, const float* K, const float* T,
                        const float* world, int64_t n, int keep_indices, float* out_image,
                        int64_t* n_out, int64_t* n_inside) {
  Camera cam(rows, cols, z_near, z_far, mat3(K), iso(T));
  Vector2fVector img;
  *n_inside = cam.projectPoints(img, vec3(world, n), keep_indices != 0);
  *n_out = (int64_t)img.size();
  for (size_t i = 0; i < img.size(); ++i) {
    out_image[2 * i] = img[i].x();
    out_image[2 * i + 1] = img[i].y();
  }
}

// PICPSolver (src/picp_solver.cpp)
void* ref_picp_create(int rows, int cols, int z_near, int z_far, const float* K, const float* T,
                      const float* world, int64_t n_world, const float* image, int64_t n_image,
                      float kernel_threshold) {
  RefPicp* r = new RefPicp();
  r->world = vec3(world, n_world);
  r->image = vec2(image, n_image);
  r->solver.setKernelThreshold(kernel_threshold);
  r->solver.init(Camera(rows, cols, z_near, z_far, mat3(K), iso(T)), r->world, r->image);
  return r;
}
void ref_picp_destroy(void* h) { delete static_cast<RefPicp*>(h); }
int ref_picp_one_round(void* h, const int32_t* pr, int64_t n, int keep_outliers) {
  RefPicp* r = static_cast<RefPicp*>(h);
  return r->solver.oneRound(pairs(pr, n), keep_outliers != 0) ? 1 : 0;
}
void ref_picp_get_state(void* h, float* T16, float* H36, float* b6, float* chi_in, float* chi_out,
                        int32_t* n_in) {
  RefPicp* r = static_cast<RefPicp*>(h);
  iso_out(r->solver.camera().worldInCameraPose(), T16);
  for (int j = 0; j < 6; ++j) {
    for (int i = 0; i < 6; ++i) H36[j * 6 + i] = r->solver.H()(i, j);
    b6[j] = r->solver.b()(j);
  }
  *chi_in = r->solver.chiInliers();
  *chi_out = r->solver.chiOutliers();
  *n_in = r->solver.numInliers();
}

// triangulate_points, 7-argument overload (src/utils.cpp:77-105) and, when app2 != NULL, the
// PointCloud overload (:106-134)
int64_t ref_triangulate_points(const float* K, const float* X, const int32_t* corr, int64_t n_corr,
                               const float* p1, int64_t n_p1, const float* p2, int64_t n_p2,
                               const float* app2, float* out_points, int32_t* out_corr_new,
                               float* out_app) {
  IntPairVector c = pairs(corr, n_corr), cn;
  int ns;
  if (!app2) {
    Vector3fVector tri;
    ns = triangulate_points(mat3(K), iso(X), c, vec2(p1, n_p1), vec2(p2, n_p2), tri, cn);
    for (int i = 0; i < ns; ++i)
      for (int k = 0; k < 3; ++k) out_points[3 * i + k] = tri[(size_t)i](k);
  } else {
    PointCloudVector<2> pc1((size_t)n_p1), pc2((size_t)n_p2);
    for (int64_t i = 0; i < n_p1; ++i) pc1.points()[(size_t)i] = Eigen::Vector2f(p1[2 * i], p1[2 * i + 1]);
    for (int64_t i = 0; i < n_p2; ++i) {
      pc2.points()[(size_t)i] = Eigen::Vector2f(p2[2 * i], p2[2 * i + 1]);
      for (int k = 0; k < 10; ++k) pc2.appearances()[(size_t)i](k) = app2[10 * i + k];
    }
    PointCloudVector<3> tri;
    ns = triangulate_points(mat3(K), iso(X), c, pc1, pc2, tri, cn);
    for (int i = 0; i < ns; ++i) {
      for (int k = 0; k < 3; ++k) out_points[3 * i + k] = tri.points()[(size_t)i](k);
      if (out_app)
        for (int k = 0; k < 10; ++k) out_app[10 * i + k] = tri.appearances()[(size_t)i](k);
    }
  }
  if (out_corr_new)
    for (int i = 0; i < ns; ++i) {
      out_corr_new[2 * i] = cn[(size_t)i].first;
      out_corr_new[2 * i + 1] = cn[(size_t)i].second;
    }
  return ns;
}

}  // extern "C"
