// whole_synthetic.cpp — BASELINE config 2: the reference's whole_test (src/tests/essential_picp_test.cpp
// :45-106) at a chosen size and with FIXED seeds: 3 views of n synthetic points, epipolar
// initialisation of view 1 (host, both builds), triangulation, 100 PICP rounds against view 2.
//
// One source, two builds (as vo_sequence.cpp):
//   * against this repository's drop-in headers + libvo_b200.so  -> host/bin/whole_synthetic (B200)
//   * against the reference's own headers and sources             -> oracle/_ref/bin/whole_synthetic
// It only uses declarations both header sets share (Camera, estimate_transform, triangulate_points,
// PICPSolver), so the second build IS the reference on the same inputs.  Every intermediate result
// is dumped in binary so a test can compare the two runs per original correspondence id:
//   correspondences, X_est, triangulated points + correspondences_new, H and b after round 1
//   (damping included, picp_solver.cpp:102), the final pose.
// The reference's own main draws its data from std::random_device (utils.cpp:9,24): not
// reproducible, hence this driver.  Distributions:
//   ref      utils.cpp:11-19,26-30 verbatim (x,y ~ U(-10,10), z = U(-10,10)*0.1+1; axis ~ U(-1,1)^3,
//            angle ~ U(-1,1), t ~ U(-1,1)^3) — only a few per cent of the points are seen by all views;
//   frustum  points drawn inside view 0's frustum (pixel ~ U(image), depth ~ U(0.5,2)), poses scaled by
//            0.1 with a mostly lateral baseline, so most of the n points survive all three views.
//
//   whole_synthetic <n_points> <seed> <ref|frustum> <rounds> <dump_file>
// prints one JSON line (stage timings in ms, counts).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>

#include "utils.h"
#include "camera.h"
#include "picp_solver.h"
#include "epipolar_utils.h"
#if defined(__has_include)
#if __has_include("vo_b200_host.h")
#include "vo_b200_host.h"  // drop-in build only: defines VO_B200_DROPIN
#endif
#endif

namespace {
using Clock = std::chrono::steady_clock;
double ms_since(Clock::time_point t0) {
  return std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
}

Eigen::Isometry3f seeded_isometry(std::mt19937& gen, float scale) {  // utils.cpp:11-19
  std::uniform_real_distribution<float> dis(-1.0f, 1.0f);
  Eigen::Vector3f a(dis(gen), dis(gen), dis(gen));
  a.normalize();
  const float angle = dis(gen) * scale;
  Eigen::Isometry3f X = Eigen::Isometry3f::Identity();
  X.linear() = Eigen::Matrix3f(Eigen::AngleAxisf(angle, a));
  X.translation() = Eigen::Vector3f(dis(gen) * scale, dis(gen) * scale, dis(gen) * scale);
  return X;
}

Vector3fVector seeded_points(std::mt19937& gen, int n, bool frustum, const Eigen::Matrix3f& k) {
  Vector3fVector pts((size_t)n);
  if (!frustum) {  // utils.cpp:26-30
    std::uniform_real_distribution<float> dis(-10.f, 10.0f);
    for (int i = 0; i < n; ++i) {
      const float x = dis(gen), y = dis(gen), z = dis(gen) * 0.1f + 1.0f;
      pts[(size_t)i] = Eigen::Vector3f(x, y, z);
    }
  } else {
    std::uniform_real_distribution<float> u(40.f, 599.f), v(40.f, 439.f), z(0.5f, 2.0f);
    for (int i = 0; i < n; ++i) {
      const float pu = u(gen), pv = v(gen), pz = z(gen);
      pts[(size_t)i] = Eigen::Vector3f((pu - k(0, 2)) * pz / k(0, 0), (pv - k(1, 2)) * pz / k(1, 1), pz);
    }
  }
  return pts;
}

// essential_picp_test.cpp:11-29
void fake_correspondences(IntPairVector& correspondences, const Vector2fVector& reference_image_points,
                          const Vector2fVector& current_measurements) {
  correspondences.resize(current_measurements.size());
  int n = 0;
  for (size_t i = 0; i < reference_image_points.size(); i++) {
    if (reference_image_points[i].x() < 0 || current_measurements[i].x() < 0) continue;
    correspondences[(size_t)n].first = (int)i;
    correspondences[(size_t)n].second = (int)i;
    n++;
  }
  correspondences.resize((size_t)n);
}

// the accumulators are protected members in both header sets
struct SolverAccess : public PICPSolver {
  void snapshot(float* H36_colmajor, float* b6) {
    (void)numInliers();  // the drop-in refreshes its host mirrors in the accessors
    for (int j = 0; j < 6; ++j)
      for (int i = 0; i < 6; ++i) H36_colmajor[j * 6 + i] = _H(i, j);
    for (int i = 0; i < 6; ++i) b6[i] = _b(i);
  }
};

void put(FILE* f, const void* p, size_t bytes) { std::fwrite(p, 1, bytes, f); }
void put_i64(FILE* f, long long v) { put(f, &v, 8); }
void put_iso(FILE* f, const Eigen::Isometry3f& X) {
  float m[16];
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) m[j * 4 + i] = (i == 3) ? (j == 3 ? 1.f : 0.f) : X(i, j);
  put(f, m, sizeof(m));
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 6) {
    std::fprintf(stderr, "usage: %s n_points seed ref|frustum rounds dump_file\n", argv[0]);
    return 2;
  }
  const int n = std::atoi(argv[1]);
  const unsigned seed = (unsigned)std::atoi(argv[2]);
  const bool frustum = std::string(argv[3]) == "frustum";
  const int rounds = std::atoi(argv[4]);
  const std::string dump = argv[5];

  Eigen::Matrix3f k;
  k << 150.f, 0.f, 320.f, 0.f, 150.f, 240.f, 0.f, 0.f, 1.f;  // essential_picp_test.cpp:54-57
  std::mt19937 gen(seed);
  Eigen::Isometry3f X_gt1 = seeded_isometry(gen, frustum ? 0.1f : 1.0f);
  Eigen::Isometry3f X_gt2 = seeded_isometry(gen, frustum ? 0.1f : 1.0f);
  if (frustum) {  // a mostly lateral baseline keeps the two-ray intersection well conditioned
    X_gt1.translation() = Eigen::Vector3f(0.45f, 0.05f, -0.03f);
    X_gt2.translation() = Eigen::Vector3f(-0.35f, -0.04f, 0.06f);
  }
  const Vector3fVector world_points_gt = seeded_points(gen, n, frustum, k);

#ifdef VO_B200_DROPIN
  {  // CUDA context + module load: what every GPU executable pays once, outside the stage timings
    PICPSolver warm;
    Vector2fVector tmp;
    Camera c0(480, 640, 0, 10, k);
    Vector3fVector one(1, Eigen::Vector3f(0.f, 0.f, 1.f));
    c0.projectPoints(tmp, one, true);
  }
#endif

  // The hot path runs twice and the SECOND pass is reported (both builds): the first one pays what a
  // process pays once — lazy loading of every kernel on its first launch, first-touch of the buffers —
  // which says nothing about the path and varied 7x from run to run.
  IntPairVector correspondences, correspondences_new;
  Vector3fVector world_points_est;
  Eigen::Isometry3f X_est = Eigen::Isometry3f::Identity(), X_picp = Eigen::Isometry3f::Identity();
  float H1[36], b1[6];
  int n_inliers = 0;
  float chi_in = 0.f;
  double ms_project = 0, ms_epipolar = 0, ms_triangulate = 0, ms_project2 = 0, ms_transform = 0, ms_picp = 0;
  for (int pass = 0; pass < 2; ++pass) {
    Camera cam(480, 640, 0, 10, k);
    Vector2fVector reference_image_points, current_measurements;
    auto t0 = Clock::now();
    cam.projectPoints(reference_image_points, world_points_gt, true);  // :65
    cam.setWorldInCameraPose(X_gt1);
    cam.projectPoints(current_measurements, world_points_gt, true);    // :67
    ms_project = ms_since(t0);

    correspondences.clear();
    fake_correspondences(correspondences, reference_image_points, current_measurements);

    t0 = Clock::now();
    X_est =
        estimate_transform(cam.cameraMatrix(), correspondences, reference_image_points, current_measurements);
    ms_epipolar = ms_since(t0);

    world_points_est.clear();
    correspondences_new.clear();
    t0 = Clock::now();
    triangulate_points(k, X_est, correspondences, reference_image_points, current_measurements, world_points_est,
                       correspondences_new);  // :78-79
    ms_triangulate = ms_since(t0);

    cam.setWorldInCameraPose(X_gt2);
    t0 = Clock::now();
    cam.projectPoints(current_measurements, world_points_gt, true);  // :87
    ms_project2 = ms_since(t0);

    SolverAccess solver;
    solver.setKernelThreshold(10000);
    t0 = Clock::now();
    Vector3fVector points_in_cameraframe1;
    points_in_cameraframe1.reserve(world_points_est.size());
    for (const auto& p : world_points_est) points_in_cameraframe1.push_back(X_est * p);  // :93-94
    ms_transform = ms_since(t0);

    cam.setWorldInCameraPose(Eigen::Isometry3f::Identity());
    t0 = Clock::now();
    solver.init(cam, points_in_cameraframe1, current_measurements);  // :97
    solver.oneRound(correspondences_new, false);
    solver.snapshot(H1, b1);
    for (int i = 1; i < rounds; i++) solver.oneRound(correspondences_new, false);  // :98-99
    X_picp = solver.camera().worldInCameraPose();          // :101 (synchronises)
    ms_picp = ms_since(t0);
    n_inliers = solver.numInliers();
    chi_in = solver.chiInliers();

  }

  if (FILE* f = std::fopen(dump.c_str(), "wb")) {
    put_i64(f, n);
    put_i64(f, (long long)correspondences.size());
    put_i64(f, (long long)world_points_est.size());
    put_iso(f, X_gt1);
    put_iso(f, X_gt2);
    put_iso(f, X_est);
    put_iso(f, X_picp);
    put(f, H1, sizeof(H1));
    put(f, b1, sizeof(b1));
    if (!correspondences.empty()) put(f, &correspondences[0].first, correspondences.size() * 8);
    if (!correspondences_new.empty()) put(f, &correspondences_new[0].first, correspondences_new.size() * 8);
    if (!world_points_est.empty()) put(f, world_points_est[0].data(), world_points_est.size() * 12);
    std::fclose(f);
  }
  std::printf(
      "{\"impl\": \"%s\", \"n_points\": %d, \"seed\": %u, \"dist\": \"%s\", \"rounds\": %d, "
      "\"n_correspondences\": %zu, \"n_triangulated\": %zu, \"n_inliers\": %d, \"chi_inliers\": %.9g, "
      "\"ms\": {\"project_2_views\": %.4f, \"epipolar_host\": %.4f, \"triangulate\": %.4f, "
      "\"project_view_2\": %.4f, \"transform_host\": %.4f, \"picp\": %.4f}}\n",
#ifdef VO_B200_DROPIN
      "b200",
#else
      "reference-cpu",
#endif
      n, seed, frustum ? "frustum" : "ref", rounds, correspondences.size(), world_points_est.size(), n_inliers,
      (double)chi_in, ms_project, ms_epipolar, ms_triangulate, ms_project2, ms_transform, ms_picp);
  return 0;
}
