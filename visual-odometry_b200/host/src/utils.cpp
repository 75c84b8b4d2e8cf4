// utils.cpp — host side of the drop-in utils: clock, synthetic generators, the one-pair
// triangulation (host) and the batch triangulations (GPU, vo_triangulate).
#include "utils.h"

#include "vo_b200_host.h"

double getTime() {
  struct timeval now;
  gettimeofday(&now, 0);
  return 1e3 * now.tv_sec + 1e-3 * now.tv_usec;
}

// random rigid motion: unit axis from U(-1,1)^3, angle and translation from U(-1,1)
// (reference utils.cpp:8-20); seeded from the OS like the reference, or from VO_B200_SEED
static std::mt19937& generator() {
  static std::mt19937 gen([] {
    const char* e = std::getenv("VO_B200_SEED");
    return e ? (unsigned)std::atoi(e) : std::random_device{}();
  }());
  return gen;
}

void generate_isometry3f(Eigen::Isometry3f& X) {
  std::uniform_real_distribution<float> unit(-1.0f, 1.0f);
  std::mt19937& gen = generator();
  Eigen::Vector3f axis;
  for (int i = 0; i < 3; ++i) axis(i) = unit(gen);
  axis.normalize();
  const float angle = unit(gen);
  X.linear() = Eigen::Matrix3f(Eigen::AngleAxisf(angle, axis));
  Eigen::Vector3f t;
  for (int i = 0; i < 3; ++i) t(i) = unit(gen);
  X.translation() = t;
}

// x,y in U(-10,10), z = U(-10,10)*0.1+1 (reference utils.cpp:22-34)
Vector3fVector generate_points3d(const int& num_points) {
  std::uniform_real_distribution<float> wide(-10.f, 10.0f);
  std::mt19937& gen = generator();
  Vector3fVector points(num_points);
  for (auto& p : points) {
    const float x = wide(gen), y = wide(gen), z = wide(gen);
    p = Eigen::Vector3f(x, y, z * 0.1f + 1.0f);
  }
  return points;
}

bool triangulate_point(const Eigen::Vector3f& d1, const Eigen::Vector3f& d2,
                       const Eigen::Vector3f& p2, Eigen::Vector3f& p) {
  Eigen::Matrix<float, 3, 2> D;
  D.col(0) = -d1;
  D.col(1) = d2;
  const Eigen::Vector2f s = -(D.transpose() * D).ldlt().solve(D.transpose() * p2);
  if (s(0) < 0 || s(1) < 0) return false;
  const Eigen::Vector3f on_first = s(0) * d1;
  const Eigen::Vector3f on_second = p2 + s(1) * d2;
  p = 0.5f * (on_first + on_second);
  return true;
}

namespace {
// shared body of the three overloads
int64_t triangulate_on_gpu(const Eigen::Matrix3f& k, const Eigen::Isometry3f& X,
                           const IntPairVector& correspondences, const float* p1, size_t n1,
                           const float* p2, size_t n2, const float* app2, float* out_points,
                           int* out_corr, float* out_app) {
  float K[9], T[16];
  vo_b200::pack3(k, K);
  vo_b200::pack_iso(X, T);
  int64_t n_success = 0;
  vo_b200::check(vo_triangulate(vo_b200::device(), K, T,
                                correspondences.empty() ? nullptr : &correspondences[0].first,
                                (int64_t)correspondences.size(), p1, (int64_t)n1, p2, (int64_t)n2,
                                app2, out_points, out_corr, out_app, nullptr, &n_success),
                 "vo_triangulate");
  return n_success;
}
const float* data2(const Vector2fVector& v) { return v.empty() ? nullptr : v[0].data(); }
}  // namespace

int triangulate_points(const Eigen::Matrix3f& k, const Eigen::Isometry3f& X,
                       const IntPairVector& correspondences, const Vector2fVector& p1_img,
                       const Vector2fVector& p2_img, Vector3fVector& triangulated) {
  triangulated.resize(correspondences.size());
  const int64_t n = triangulate_on_gpu(k, X, correspondences, data2(p1_img), p1_img.size(),
                                       data2(p2_img), p2_img.size(), nullptr,
                                       triangulated.empty() ? nullptr : triangulated[0].data(),
                                       nullptr, nullptr);
  triangulated.resize((size_t)n);
  return (int)n;
}

int triangulate_points(const Eigen::Matrix3f& k, const Eigen::Isometry3f& X,
                       const IntPairVector& correspondences, const Vector2fVector& p1_img,
                       const Vector2fVector& p2_img, Vector3fVector& triangulated,
                       IntPairVector& correspondences_new) {
  triangulated.resize(correspondences.size());
  correspondences_new.resize(correspondences.size());
  const int64_t n = triangulate_on_gpu(
      k, X, correspondences, data2(p1_img), p1_img.size(), data2(p2_img), p2_img.size(), nullptr,
      triangulated.empty() ? nullptr : triangulated[0].data(),
      correspondences_new.empty() ? nullptr : &correspondences_new[0].first, nullptr);
  triangulated.resize((size_t)n);
  correspondences_new.resize((size_t)n);
  return (int)n;
}

int triangulate_points(const Eigen::Matrix3f& k, const Eigen::Isometry3f& X,
                       const IntPairVector& correspondences, const PointCloudVector<2>& pc_1,
                       const PointCloudVector<2>& pc_2, PointCloudVector<3>& triangulated,
                       IntPairVector& correspondences_new) {
  // points()/appearances() of a const cloud return copies (PointCloud.h): take them once
  const auto p1 = pc_1.points();
  const auto p2 = pc_2.points();
  const Vector10fVector app2 = pc_2.appearances();
  triangulated.clear();
  triangulated.resize(correspondences.size());
  correspondences_new.resize(correspondences.size());
  const int64_t n = triangulate_on_gpu(
      k, X, correspondences, p1.empty() ? nullptr : p1[0].data(), p1.size(),
      p2.empty() ? nullptr : p2[0].data(), p2.size(), app2.empty() ? nullptr : app2[0].data(),
      triangulated.size() ? triangulated.points()[0].data() : nullptr,
      correspondences_new.empty() ? nullptr : &correspondences_new[0].first,
      triangulated.size() ? triangulated.appearances()[0].data() : nullptr);
  triangulated.resize((size_t)n);
  correspondences_new.resize((size_t)n);
  return (int)n;
}
