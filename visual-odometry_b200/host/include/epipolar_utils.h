// epipolar_utils.h — declarations of the reference's epipolar initialisation
// (include/epipolar_utils.h).  It runs once per sequence on the CPU and is outside the GPU hot
// path: the definitions are the reference's own src/epipolar_utils.cpp, compiled unchanged by
// build_dropin.sh (its four cheirality tests call triangulate_points, i.e. the GPU).
#pragma once
#include "defs.h"
#include "utils.h"

const Eigen::Matrix3f transform2essential(const Eigen::Isometry3f X);
const Eigen::Matrix3f estimate_essential(const Eigen::Matrix3f& k, const IntPairVector& correspondences,
                                         const Vector2fVector& p1_img, const Vector2fVector& p2_img);
const IsometryPair essential2transformPair(const Eigen::Matrix3f& E);
Vector2fVector normalize(const Vector2fVector& p, Eigen::Matrix3f& T);
Vector2fVector normalizeGauss(const Vector2fVector& p, Eigen::Matrix3f& T);
const Eigen::Isometry3f estimate_transform(const Eigen::Matrix3f k, const IntPairVector& correspondences,
                                           const Vector2fVector& p1_img, const Vector2fVector& p2_img);
