// brute_force_search.h — drop-in for the reference's include/brute_force_search.h.
// Same two function templates over an iterator range of fixed-size float points whose first
// coefficient is an id that does not take part in the distance.  The scan runs on the GPU:
// the range is uploaded as the map of a vo_nn handle and the query answered by the exact
// brute-force kernel (strict '<' against norm*norm, lowest address wins ties, nullptr if none).
#pragma once
#include <iterator>
#include <mutex>
#include <vector>

#include "vo_b200_host.h"

namespace vo_b200 {
// RAII wrapper of one resident map.  Handles (stream + grow-only device buffers) are recycled
// through a small process-wide pool: the pipeline builds a search structure per frame
// (vo_complete.cpp:35), and creating / destroying CUDA streams and allocations at that rate costs
// more than the search itself.
class NNMap {
 public:
  NNMap() : _h(acquire()) {}
  ~NNMap() { release(_h); }
  NNMap(const NNMap&) = delete;
  NNMap& operator=(const NNMap&) = delete;
  void setRows(const float* rows, long n, int stride) {
    check(vo_nn_set_map(_h, rows, n, stride, 1), "vo_nn_set_map");
  }
  int bestMatch(const float* query, int stride, float norm) {
    int32_t idx = -1;
    check(vo_nn_best_match(_h, query, 1, stride, norm, &idx, nullptr), "vo_nn_best_match");
    return idx;
  }
  void bestMatches(const float* queries, long n, int stride, float norm, int32_t* out) {
    check(vo_nn_best_match(_h, queries, n, stride, norm, out, nullptr), "vo_nn_best_match");
  }
  std::vector<int32_t> within(const float* query, int stride, float norm) {
    int32_t count = 0;
    check(vo_nn_radius_search(_h, query, 1, stride, norm, &count, nullptr, 0), "vo_nn_radius_search");
    std::vector<int32_t> rows((size_t)count);
    if (count > 0)
      check(vo_nn_radius_search(_h, query, 1, stride, norm, &count, rows.data(), count),
            "vo_nn_radius_search");
    return rows;
  }

 private:
  struct Pool {
    std::mutex mu;
    std::vector<vo_nn_t> idle;
    ~Pool() {
      for (vo_nn_t h : idle) vo_nn_destroy(h);
    }
  };
  static Pool& pool() {
    static Pool p;
    return p;
  }
  static vo_nn_t acquire() {
    {
      Pool& p = pool();
      std::lock_guard<std::mutex> lock(p.mu);
      if (!p.idle.empty()) {
        vo_nn_t h = p.idle.back();
        p.idle.pop_back();
        return h;
      }
    }
    vo_nn_t h = nullptr;
    check(vo_nn_create(&h, device()), "vo_nn_create");
    return h;
  }
  static void release(vo_nn_t h) {
    Pool& p = pool();
    std::lock_guard<std::mutex> lock(p.mu);
    if (p.idle.size() < 8) p.idle.push_back(h);
    else vo_nn_destroy(h);
  }
  vo_nn_t _h;
};

template <typename PointType>
struct point_layout {
  static_assert(sizeof(typename PointType::Scalar) == sizeof(float), "float points only");
  enum { stride = sizeof(PointType) / sizeof(float) };
  static_assert(stride == PointType::RowsAtCompileTime, "points must be tightly packed");
};
}  // namespace vo_b200

// every point of [begin,end) closer than `norm` to `query` (id column ignored); pointers are
// appended to `answers` in container order; returns how many (reference :3-20)
template <typename IteratorType_>
int bruteForceSearch(std::vector<typename IteratorType_::value_type*>& answers, IteratorType_ begin,
                     IteratorType_ end, const typename IteratorType_::value_type& query,
                     const typename IteratorType_::value_type::Scalar norm) {
  using Point = typename IteratorType_::value_type;
  const long n = (long)std::distance(begin, end);
  if (n <= 0) return 0;
  vo_b200::NNMap map;
  map.setRows((*begin).data(), n, vo_b200::point_layout<Point>::stride);
  const std::vector<int32_t> rows = map.within(query.data(), vo_b200::point_layout<Point>::stride, norm);
  for (int32_t r : rows) answers.push_back(&*(begin + r));
  return (int)rows.size();
}

// the closest point of [begin,end) among those closer than `norm`, or nullptr (reference :22-41)
template <typename IteratorType_>
typename IteratorType_::value_type* bruteForceBestMatch(
    IteratorType_ begin, IteratorType_ end, const typename IteratorType_::value_type& query,
    const typename IteratorType_::value_type::Scalar norm) {
  using Point = typename IteratorType_::value_type;
  const long n = (long)std::distance(begin, end);
  if (n <= 0) return nullptr;
  vo_b200::NNMap map;
  map.setRows((*begin).data(), n, vo_b200::point_layout<Point>::stride);
  const int row = map.bestMatch(query.data(), vo_b200::point_layout<Point>::stride, norm);
  return row < 0 ? nullptr : &*(begin + row);
}
