// TEST INFRASTRUCTURE: see shared_memory_object.hpp.
#pragma once
#include "shared_memory_object.hpp"

namespace boost {
namespace interprocess {

class mapped_region {
public:
    mapped_region() = default;
    mapped_region(const shared_memory_object &shm, mode_t mode) {
        struct stat st;
        if (fstat(shm.fd(), &st) != 0)
            throw interprocess_exception(std::string("fstat: ") + std::strerror(errno));
        size_ = (size_t)st.st_size;
        addr_ = mmap(nullptr, size_, mode == read_write ? PROT_READ | PROT_WRITE : PROT_READ, MAP_SHARED,
                     shm.fd(), 0);
        if (addr_ == MAP_FAILED) {
            addr_ = nullptr;
            throw interprocess_exception(std::string("mmap: ") + std::strerror(errno));
        }
    }
    mapped_region(mapped_region &&o) noexcept : addr_(o.addr_), size_(o.size_) {
        o.addr_ = nullptr;
        o.size_ = 0;
    }
    mapped_region &operator=(mapped_region &&o) noexcept {
        if (this != &o) {
            unmap();
            addr_ = o.addr_;
            size_ = o.size_;
            o.addr_ = nullptr;
            o.size_ = 0;
        }
        return *this;
    }
    mapped_region(const mapped_region &) = delete;
    mapped_region &operator=(const mapped_region &) = delete;
    ~mapped_region() {
        unmap();
    }
    size_t get_size() const {
        return size_;
    }
    void *get_address() const {
        return addr_;
    }

private:
    void unmap() {
        if (addr_)
            munmap(addr_, size_);
        addr_ = nullptr;
    }
    void *addr_ = nullptr;
    size_t size_ = 0;
};

} // namespace interprocess
} // namespace boost
