// TEST INFRASTRUCTURE: see interprocess_mutex.hpp.
#pragma once
#include <cstdint>
namespace boost {
namespace interprocess {
struct interprocess_condition {
    uint64_t word[6] = {0, 0, 0, 0, 0, 0}; // 48 bytes, the size of a pthread_cond_t
};
} // namespace interprocess
} // namespace boost
