// defs.h — drop-in for the reference's include/defs.h: the same type NAMES (Vector11f,
// IntPairVector, Vector3fVector, ...) so that the reference's mains compile unchanged against
// the B200 host layer.  Only names are shared; everything is expressed through two aliases.
#pragma once
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <Eigen/Cholesky>
#include <Eigen/StdVector>

#include <iostream>
#include <utility>
#include <vector>

namespace vo_b200 {
template <int R, int C = 1>
using Mat = Eigen::Matrix<float, R, C>;
template <class T>
using AlignedVector = std::vector<T, Eigen::aligned_allocator<T>>;
}  // namespace vo_b200

// fixed-size vectors / matrices (reference include/defs.h:7-16)
using Vector6f = vo_b200::Mat<6>;
using Vector9f = vo_b200::Mat<9>;
using Vector10f = vo_b200::Mat<10>;   // appearance descriptor
using Vector11f = vo_b200::Mat<11>;   // [id | appearance], the kd-tree / brute-force point type
using RowVector9f = vo_b200::Mat<1, 9>;
using Matrix2_3f = vo_b200::Mat<2, 3>;
using Matrix2_6f = vo_b200::Mat<2, 6>;
using Matrix3_6f = vo_b200::Mat<3, 6>;
using Matrix6f = vo_b200::Mat<6, 6>;
using Matrix9f = vo_b200::Mat<9, 9>;

// index pairs (defs.h:18-20): std::pair<int,int> is two packed int32 — the layout the C ABI takes
using IntPair = std::pair<int, int>;
using IntPairVector = std::vector<IntPair>;
using IsometryPair = std::pair<Eigen::Isometry3f, Eigen::Isometry3f>;
static_assert(sizeof(IntPair) == 2 * sizeof(int), "IntPair must be two packed ints");

// containers (defs.h:22-29)
using Vector2fVector = vo_b200::AlignedVector<Eigen::Vector2f>;
using Vector2iVector = vo_b200::AlignedVector<Eigen::Vector2i>;
using Vector3fVector = vo_b200::AlignedVector<Eigen::Vector3f>;
using Vector4fVector = vo_b200::AlignedVector<Eigen::Vector4f>;
using Vector6fVector = vo_b200::AlignedVector<Vector6f>;
using Vector10fVector = vo_b200::AlignedVector<Vector10f>;
using Vector11fVector = vo_b200::AlignedVector<Vector11f>;
using IsometryVector = vo_b200::AlignedVector<Eigen::Isometry3f>;
static_assert(sizeof(Eigen::Vector2f) == 8 && sizeof(Eigen::Vector3f) == 12 &&
                  sizeof(Vector10f) == 40 && sizeof(Vector11f) == 44,
              "the C ABI reads these containers in place and relies on tight packing");
