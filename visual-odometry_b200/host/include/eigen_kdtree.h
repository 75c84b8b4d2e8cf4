// eigen_kdtree.h — drop-in for the reference's include/eigen_kdtree.h (class TreeNode_).
//
// The reference builds a PCA kd-tree over the range (reordering it in place) and answers
// within-radius queries by descending it.  Here the constructor uploads the range once as the
// resident map of the exact brute-force GPU kernel, and every query method asks that kernel:
//   bestMatchFull  — identical answers (it is exact within the radius in the reference too),
//                    except on exact distance ties, where this returns the lowest-index point
//                    and the reference the one from its right subtree;
//   bestMatchFast  — the reference may MISS the true neighbour near a split plane; this never does;
//   fullSearch / fastSearch — all points within the radius, in container order.
// The range is NOT reordered; like the reference, returned pointers point into the caller's
// container, whose id column tells the caller which point it is (vo_complete.cpp:41-43).
#include <memory>
#include <vector>

#include "brute_force_search.h"

template <typename IteratorType_>
class TreeNode_ {
 public:
  using IteratorType = IteratorType_;
  using PointType = typename IteratorType_::value_type;
  using Scalar = typename PointType::Scalar;
  static constexpr int Dim = PointType::RowsAtCompileTime;
  using ThisType = TreeNode_<IteratorType>;
  using PtrType = std::unique_ptr<ThisType>;
  using AnswerType = std::vector<PointType*>;

  // max_points_in_leaf is accepted for compatibility; there are no leaves any more
  TreeNode_(IteratorType begin_, IteratorType end_, int max_points_in_leaf = 20)
      : _begin(begin_), _end(end_), _size((long)std::distance(begin_, end_)) {
    (void)max_points_in_leaf;
    if (_size > 0) _map.setRows((*_begin).data(), _size, stride());
  }

  void fastSearch(AnswerType& answers, const PointType& query, Scalar norm) { fullSearch(answers, query, norm); }
  void fullSearch(AnswerType& answers, const PointType& query, Scalar norm) {
    if (_size <= 0) return;
    for (int32_t r : _map.within(query.data(), stride(), norm)) answers.push_back(&*(_begin + r));
  }
  PointType* bestMatchFast(const PointType& query, Scalar norm) { return bestMatchFull(query, norm); }
  PointType* bestMatchFull(const PointType& query, Scalar norm) {
    if (_size <= 0) return nullptr;
    const int row = _map.bestMatch(query.data(), stride(), norm);
    return row < 0 ? nullptr : &*(_begin + row);
  }
  // addition: a whole batch of queries in one launch; out[i] = index into [begin,end) or -1
  template <typename QueryIterator>
  void bestMatchFullBatch(QueryIterator qbegin, QueryIterator qend, Scalar norm, std::vector<int>& out) {
    const long nq = (long)std::distance(qbegin, qend);
    out.assign((size_t)nq, -1);
    if (nq <= 0 || _size <= 0) return;
    _map.bestMatches((*qbegin).data(), nq, stride(), norm, out.data());
  }

 protected:
  static constexpr int stride() { return vo_b200::point_layout<PointType>::stride; }
  IteratorType _begin, _end;
  long _size;
  vo_b200::NNMap _map;
};
