// PointCloud.h — drop-in for the reference's include/PointCloud.h: a point with its 10-D
// appearance, and the structure-of-arrays container the pipeline passes around.  Host-side data
// structure; the GPU entry points read points() / appearances() in place.
#pragma once
#include "defs.h"

template <int dim>
class PointCloud {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  using Point = Eigen::Matrix<float, dim, 1>;
  PointCloud() {}
  PointCloud(const Point& point, const Vector10f& appearance) : _point(point), _appearance(appearance) {}
  inline Point point() const { return _point; }
  inline Vector10f appearance() const { return _appearance; }

 protected:
  Point _point;
  Vector10f _appearance;
};

template <int dim>
class PointCloudVector {
  using Point = Eigen::Matrix<float, dim, 1>;
  using PointsVec = std::vector<Point, Eigen::aligned_allocator<Point>>;

 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  PointCloudVector() {}
  PointCloudVector(size_t N) : _points(N), _appearances(N) {}

  inline size_t size() const { return _points.size(); }
  void clear() {
    _points.clear();
    _appearances.clear();
  }
  void resize(size_t N) {
    _points.resize(N);
    _appearances.resize(N);
  }
  void reserve(size_t N) {
    _points.reserve(N);
    _appearances.reserve(N);
  }
  void push_back(const PointCloud<dim>& pc) {
    _points.push_back(pc.point());
    _appearances.push_back(pc.appearance());
  }
  // merge `cloud` into this one (reference PointCloud.h:52-66): a point whose appearance is
  // already present (exact float equality, first hit) replaces the stored position; every other
  // point is appended, in order.  Appended points take part in the matching of later ones.
  void update(const PointCloudVector<dim>& cloud) {
    const PointsVec& new_points = cloud._points;
    const Vector10fVector& new_appearances = cloud._appearances;
    for (size_t i = 0; i < new_points.size(); ++i) {
      size_t hit = _appearances.size();
      for (size_t j = 0; j < _appearances.size(); ++j)
        if (_appearances[j] == new_appearances[i]) {
          hit = j;
          break;
        }
      if (hit < _appearances.size()) {
        _points[hit] = new_points[i];
      } else {
        _points.push_back(new_points[i]);
        _appearances.push_back(new_appearances[i]);
      }
    }
  }

  inline PointsVec& points() { return _points; }
  inline Vector10fVector& appearances() { return _appearances; }
  inline PointsVec points() const { return _points; }
  inline Vector10fVector appearances() const { return _appearances; }

 protected:
  PointsVec _points;
  Vector10fVector _appearances;
};

// every point of the cloud moved by X; appearances are carried over (PointCloud.h:77-82)
inline PointCloudVector<3> operator*(Eigen::Isometry3f X, const PointCloudVector<3> pc) {
  PointCloudVector<3> moved(pc.size());
  const auto pts = pc.points();
  for (size_t i = 0; i < pts.size(); ++i) moved.points()[i] = X * pts[i];
  moved.appearances() = pc.appearances();
  return moved;
}
