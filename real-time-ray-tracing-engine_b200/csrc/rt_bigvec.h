// rt_bigvec.h - a std::vector for the per-primitive host arrays of very large scenes.
#pragma once

#include <memory>
#include <utility>
#include <vector>

namespace rtflat {

// The per-primitive arrays of a million-primitive scene are hundreds of megabytes: a std::vector would zero-fill
// them (and take their page faults) on the calling thread before the worker threads overwrite every element.
// BigVec::resize leaves new elements as raw storage - whoever resizes writes every element, on whatever thread.
template <class T> struct RawAlloc : std::allocator<T> {
  template <class U> struct rebind {
    using other = RawAlloc<U>;
  };
  RawAlloc() = default;
  template <class U> RawAlloc(const RawAlloc<U> &) {}
  template <class U, class... A> void construct(U *p, A &&...a) {
    if constexpr (sizeof...(A) > 0)
      ::new ((void *)p) U(std::forward<A>(a)...);
  }
};
template <class T> using BigVec = std::vector<T, RawAlloc<T>>;

} // namespace rtflat
