// TEST INFRASTRUCTURE ONLY - stands in for <tbb/tbb.h> (TBB is not in this image): the parallel algorithms the reference's solver and
// multigrid headers use, executed serially (parallel_for over an index range or a blocked_range, parallel_reduce over ONE range - the
// reference's results do not depend on the partition beyond the rounding of a sum).
#pragma once
#include <cstddef>
#include <mutex>
namespace tbb {
struct split {};
template <class T>
class blocked_range {
    T b_, e_;
    std::size_t g_;

public:
    typedef T const_iterator;
    blocked_range(T b, T e, std::size_t grain = 1) : b_(b), e_(e), g_(grain) {}
    T begin() const { return b_; }
    T end() const { return e_; }
    std::size_t size() const { return (std::size_t)(e_ - b_); }
    std::size_t grainsize() const { return g_; }
    bool empty() const { return !(b_ < e_); }
};
template <class Index, class F>
inline void parallel_for(Index first, Index last, const F& f) { for (Index i = first; i < last; ++i) f(i); }
template <class T, class F>
inline void parallel_for(const blocked_range<T>& r, const F& f) { f(r); }
template <class Range, class Value, class F, class Join>
inline Value parallel_reduce(const Range& r, const Value& init, const F& f, const Join&) { return f(r, init); }
} // namespace tbb
