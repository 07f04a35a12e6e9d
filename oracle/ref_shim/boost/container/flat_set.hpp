// TEST INFRASTRUCTURE.  Boost is not installed in this image; the reference's Laia scheduler
// (laia/include/utils.h:7) uses boost::container::flat_set only as an ordered set of keys
// (emplace / clear / begin / end / find).  std::set has the same observable behaviour.
#pragma once
#include <set>
namespace boost {
namespace container {
template <class T>
using flat_set = std::set<T>;
}
} // namespace boost
