// split.h — drop-in for the reference's include/split.h: in-place two-way partition.
// Kept for source compatibility (the GPU-backed TreeNode_ no longer reorders its points).
#pragma once
#include <algorithm>
#include <iterator>

// elements satisfying `predicate` first; returns the first element of the second class
template <typename IteratorType_, typename PredicateType_>
IteratorType_ split(IteratorType_ begin, IteratorType_ end, PredicateType_ predicate) {
  IteratorType_ left = begin, right = end;
  while (left != right) {
    if (predicate(*left)) {
      ++left;
    } else {
      --right;
      std::iter_swap(left, right);
    }
  }
  return left;
}
