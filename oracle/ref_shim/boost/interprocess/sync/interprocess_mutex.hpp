// TEST INFRASTRUCTURE: share_mem.h only placement-constructs these in its control region; the
// locking calls are commented out in the reference (share_mem.h:128,143).
#pragma once
#include <cstdint>
namespace boost {
namespace interprocess {
struct interprocess_mutex {
    uint64_t word[5] = {0, 0, 0, 0, 0}; // 40 bytes, the size of a pthread_mutex_t
};
} // namespace interprocess
} // namespace boost
