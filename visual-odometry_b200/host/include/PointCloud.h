// PointCloud.h — drop-in for the reference's include/PointCloud.h: a point with its 10-D
// appearance, and the structure-of-arrays container the pipeline passes around.  Host-side data
// structure; the GPU entry points read points() / appearances() in place.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

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
    _index_valid = false;
  }
  void resize(size_t N) {
    _points.resize(N);
    _appearances.resize(N);
    _index_valid = false;
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
  // The reference scans the whole map for every new point (O(|cloud| x |map|), the dominant cost
  // of a long sequence: SURVEY.md §8f.1); here the first index of every distinct appearance is
  // kept in a hash index that stays valid across update()/push_back() calls and is rebuilt
  // only after the containers were exposed for writing.  Same results, O(|cloud|) per call.
  void update(const PointCloudVector<dim>& cloud) {
    if (!_index_valid || _indexed != _appearances.size()) rebuild_index();
    const PointsVec& new_points = cloud._points;
    const Vector10fVector& new_appearances = cloud._appearances;
    const size_t n = new_points.size();
    // the table must not grow while slots are being prefetched: make room for the whole cloud
    if ((_appearances.size() + n + 1) * 2 > _slots.size()) rebuild_index(n);
    // Random probes into a multi-megabyte table are bound by memory latency, not by work: hash a
    // block of new points first and prefetch their home slots, then prefetch the stored
    // appearances those slots point at, and only then resolve the block in order.
    constexpr size_t kBlock = 64;
    uint64_t h[kBlock];
    bool comparable[kBlock];
    const size_t mask = _slots.size() - 1;
    for (size_t b0 = 0; b0 < n; b0 += kBlock) {
      const size_t bn = std::min(kBlock, n - b0);
      for (size_t k = 0; k < bn; ++k) {
        comparable[k] = hash_of(new_appearances[b0 + k], h[k]);  // false: holds a NaN, equals nothing
        __builtin_prefetch(&_slots[(size_t)h[k] & mask]);
      }
      for (size_t k = 0; k < bn; ++k) {
        const uint32_t j = _slots[(size_t)h[k] & mask];
        if (comparable[k] && j != kEmpty) {
          __builtin_prefetch(&_appearances[j]);
          __builtin_prefetch(&_points[j], 1);
        }
      }
      for (size_t k = 0; k < bn; ++k) {
        const size_t i = b0 + k;
        if (comparable[k]) {
          const size_t slot = find_slot(new_appearances[i], h[k]);
          if (_slots[slot] != kEmpty) {
            _points[_slots[slot]] = new_points[i];
            continue;
          }
          _slots[slot] = (uint32_t)_appearances.size();
        }
        _points.push_back(new_points[i]);
        _appearances.push_back(new_appearances[i]);
        _indexed = _appearances.size();
      }
    }
  }

  // non-const access may change appearances behind the index: it is rebuilt on the next update()
  inline PointsVec& points() { return _points; }
  inline Vector10fVector& appearances() {
    _index_valid = false;
    return _appearances;
  }
  inline PointsVec points() const { return _points; }
  inline Vector10fVector appearances() const { return _appearances; }

 protected:
  // Open-addressing table over the stored appearances: a slot holds the FIRST index whose
  // appearance hashes there (float equality: -0.0 == +0.0 hash alike, a NaN equals nothing and is
  // never indexed).  Linear probing; at most half full.
  static constexpr uint32_t kEmpty = 0xFFFFFFFFu;
  static bool hash_of(const Vector10f& a, uint64_t& h) {
    h = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < 10; ++i) {
      const float v = a(i);
      if (v != v) return false;
      const float z = (v == 0.f) ? 0.f : v;
      uint32_t w;
      std::memcpy(&w, &z, sizeof(float));
      h = (h ^ w) * 0x100000001B3ull;
      h ^= h >> 29;
    }
    return true;
  }
  // the slot holding an appearance equal to `a`, or the empty slot where it belongs
  size_t find_slot(const Vector10f& a, uint64_t h) const {
    const size_t mask = _slots.size() - 1;
    size_t s = (size_t)h & mask;
    while (_slots[s] != kEmpty && !(_appearances[_slots[s]] == a)) s = (s + 1) & mask;
    return s;
  }
  void rebuild_index(size_t extra = 0) {
    size_t cap = 1024;
    while (cap < 4 * (_appearances.size() + extra + 1)) cap *= 2;
    _slots.assign(cap, kEmpty);
    for (size_t j = 0; j < _appearances.size(); ++j) {
      uint64_t h;
      if (!hash_of(_appearances[j], h)) continue;
      const size_t s = find_slot(_appearances[j], h);
      if (_slots[s] == kEmpty) _slots[s] = (uint32_t)j;  // keeps the first
    }
    _indexed = _appearances.size();
    _index_valid = true;
  }

  PointsVec _points;
  Vector10fVector _appearances;
  std::vector<uint32_t> _slots;
  size_t _indexed = 0;
  bool _index_valid = false;
};

// every point of the cloud moved by X; appearances are carried over (PointCloud.h:77-82)
inline PointCloudVector<3> operator*(Eigen::Isometry3f X, const PointCloudVector<3> pc) {
  PointCloudVector<3> moved(pc.size());
  const auto pts = pc.points();
  for (size_t i = 0; i < pts.size(); ++i) moved.points()[i] = X * pts[i];
  moved.appearances() = pc.appearances();
  return moved;
}
