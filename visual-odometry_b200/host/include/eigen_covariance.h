// eigen_covariance.h — drop-in for the reference's include/eigen_covariance.h: sample mean /
// covariance of the appearance part of a point range and its principal direction (host code;
// kept for source compatibility, the GPU-backed TreeNode_ does not need a split plane).
#pragma once
#include <Eigen/Core>
#include <Eigen/Eigenvalues>

template <typename Iterator_>
int computeMeanAndCovariance(
    Eigen::Matrix<typename Iterator_::value_type::Scalar, Iterator_::value_type::RowsAtCompileTime - 1, 1>& mean,
    Eigen::Matrix<typename Iterator_::value_type::Scalar, Iterator_::value_type::RowsAtCompileTime - 1,
                  Iterator_::value_type::RowsAtCompileTime - 1>& cov,
    Iterator_ begin, Iterator_ end) {
  using Scalar = typename Iterator_::value_type::Scalar;
  constexpr int D = Iterator_::value_type::RowsAtCompileTime - 1;
  mean.setZero();
  cov.setZero();
  int count = 0;
  for (Iterator_ it = begin; it != end; ++it, ++count) {
    const Eigen::Matrix<Scalar, D, 1> a = (*it).tail(D);
    mean += a;
    cov += a * a.transpose();
  }
  mean *= (1. / count);
  cov *= (1. / count);
  cov -= mean * mean.transpose();
  cov *= Scalar(count) / Scalar(count - 1);  // unbiased
  return count;
}

template <typename SquareMatrixType_>
Eigen::Matrix<typename SquareMatrixType_::Scalar, SquareMatrixType_::RowsAtCompileTime, 1>
largestEigenVector(const SquareMatrixType_& m) {
  Eigen::SelfAdjointEigenSolver<SquareMatrixType_> solver;
  solver.compute(m);
  return solver.eigenvectors().col(SquareMatrixType_::RowsAtCompileTime - 1);  // ascending order
}
