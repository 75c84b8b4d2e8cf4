// utils.h — drop-in for the reference's include/utils.h: small rotation helpers (host, inline),
// the synthetic generators (host) and the triangulation entry points, whose batch overloads run
// on the GPU (vo_triangulate).
#pragma once
#include <Eigen/Eigenvalues>
#include <sys/time.h>

#include <cmath>
#include <iostream>
#include <random>
#include <unordered_set>

#include "PointCloud.h"
#include "defs.h"

// wall clock in milliseconds (reference utils.cpp:2-6)
double getTime();

// elementary rotations and their XYZ composition (reference utils.h:12-61)
template <typename Scalar_>
Eigen::Matrix<Scalar_, 3, 3> RotationX(const Scalar_& angle) {
  const Scalar_ c = cos(angle), s = sin(angle), o(1.), z(0.);
  Eigen::Matrix<Scalar_, 3, 3> R;
  R << o, z, z,
       z, c, -s,
       z, s, c;
  return R;
}
template <typename Scalar_>
Eigen::Matrix<Scalar_, 3, 3> RotationY(const Scalar_& angle) {
  const Scalar_ c = cos(angle), s = sin(angle), o(1.), z(0.);
  Eigen::Matrix<Scalar_, 3, 3> R;
  R << c, z, s,
       z, o, z,
       -s, z, c;
  return R;
}
template <typename Scalar_>
Eigen::Matrix<Scalar_, 3, 3> RotationZ(const Scalar_& angle) {
  const Scalar_ c = cos(angle), s = sin(angle), o(1.), z(0.);
  Eigen::Matrix<Scalar_, 3, 3> R;
  R << c, -s, z,
       s, c, z,
       z, z, o;
  return R;
}
template <typename Scalar_>
Eigen::Matrix<Scalar_, 3, 3> Rotation(const Eigen::Matrix<Scalar_, 3, 1>& angles) {
  return RotationX(angles.x()) * RotationY(angles.y()) * RotationZ(angles.z());
}

// (x y z thx thy thz) -> isometry, Euler XYZ (reference utils.h:64-78)
inline Eigen::Isometry3f v2tEuler(const Vector6f& v) {
  Eigen::Isometry3f T = Eigen::Isometry3f::Identity();
  T.translation() = v.head<3>();
  T.linear() = Rotation(Eigen::Vector3f(v.tail<3>()));
  return T;
}

// eigenvector of the smallest eigenvalue of a symmetric matrix (reference utils.h:80-92)
template <typename SquareMatrixType_>
Eigen::Matrix<typename SquareMatrixType_::Scalar, SquareMatrixType_::RowsAtCompileTime, 1>
smallestEigenVector(const SquareMatrixType_& m) {
  Eigen::SelfAdjointEigenSolver<SquareMatrixType_> solver;
  solver.compute(m);
  return solver.eigenvectors().col(0);
}

// [v]x (reference utils.h:96-102)
inline Eigen::Matrix3f skew(const Eigen::Vector3f& v) {
  Eigen::Matrix3f S;
  S << 0, -v.z(), v.y(),
       v.z(), 0, -v.x(),
       -v.y(), v.x(), 0;
  return S;
}

// synthetic data (reference utils.cpp:8-34); host code, same distributions
void generate_isometry3f(Eigen::Isometry3f& X);
Vector3fVector generate_points3d(const int& num_points);

// one pair of rays, on the host (reference utils.cpp:36-49): midpoint of the closest approach of
// the ray {s*d1} and the ray {p2 + s*d2}; false when either ray parameter is negative.
bool triangulate_point(const Eigen::Vector3f& d1, const Eigen::Vector3f& d2,
                       const Eigen::Vector3f& p2, Eigen::Vector3f& p);

// batches, on the GPU (reference utils.cpp:51-134).  k: camera matrix; X: pose of the first
// camera in the frame of the second; correspondences: (index in image 1, index in image 2).
// Outputs are resized to the number of successes, which is also returned; successes keep the
// order of `correspondences`; correspondences_new[n] = (index in image 2, n).
int triangulate_points(const Eigen::Matrix3f& k, const Eigen::Isometry3f& X,
                       const IntPairVector& correspondences, const Vector2fVector& p1_img,
                       const Vector2fVector& p2_img, Vector3fVector& triangulated);
int triangulate_points(const Eigen::Matrix3f& k, const Eigen::Isometry3f& X,
                       const IntPairVector& correspondences, const Vector2fVector& p1_img,
                       const Vector2fVector& p2_img, Vector3fVector& triangulated,
                       IntPairVector& correspondences_new);
// point-cloud flavour: the appearance of the image-2 point travels with the triangulated point
int triangulate_points(const Eigen::Matrix3f& k, const Eigen::Isometry3f& X,
                       const IntPairVector& correspondences, const PointCloudVector<2>& pc_1,
                       const PointCloudVector<2>& pc_2, PointCloudVector<3>& triangulated,
                       IntPairVector& correspondences_new);
