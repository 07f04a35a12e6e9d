// TEST INFRASTRUCTURE.  Minimal POSIX-backed stand-ins for the boost::interprocess names that
// laia/include/share_mem.h uses (shared_memory_object, mapped_region, interprocess_mutex /
// _condition, the open tags and modes).  Enough to compile the reference's TopkScheduler
// unmodified and to run its local-shared ring buffers on this box.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>
#include <stdexcept>
#include <string>

namespace boost {
namespace interprocess {

enum mode_t { read_only = 0, read_write = 1 };
struct open_only_t {};
struct open_or_create_t {};
struct create_only_t {};
static const open_only_t open_only = open_only_t();
static const open_or_create_t open_or_create = open_or_create_t();
static const create_only_t create_only = create_only_t();

class interprocess_exception : public std::runtime_error {
public:
    explicit interprocess_exception(const std::string &m) : std::runtime_error(m) {}
};

class shared_memory_object {
public:
    shared_memory_object() = default;
    shared_memory_object(open_only_t, const char *name, mode_t mode) {
        open(name, mode == read_write ? O_RDWR : O_RDONLY);
    }
    shared_memory_object(open_or_create_t, const char *name, mode_t mode) {
        open(name, (mode == read_write ? O_RDWR : O_RDONLY) | O_CREAT);
    }
    shared_memory_object(shared_memory_object &&o) noexcept : fd_(o.fd_) {
        o.fd_ = -1;
    }
    shared_memory_object &operator=(shared_memory_object &&o) noexcept {
        if (this != &o) {
            close_fd();
            fd_ = o.fd_;
            o.fd_ = -1;
        }
        return *this;
    }
    shared_memory_object(const shared_memory_object &) = delete;
    shared_memory_object &operator=(const shared_memory_object &) = delete;
    ~shared_memory_object() {
        close_fd();
    }
    void truncate(long long size) {
        if (ftruncate(fd_, size) != 0)
            throw interprocess_exception(std::string("ftruncate: ") + std::strerror(errno));
    }
    static bool remove(const char *name) {
        return shm_unlink((std::string("/") + name).c_str()) == 0;
    }
    int fd() const {
        return fd_;
    }

private:
    void open(const char *name, int flags) {
        fd_ = shm_open((std::string("/") + name).c_str(), flags, 0644);
        if (fd_ < 0)
            throw interprocess_exception(std::string("shm_open: ") + std::strerror(errno));
    }
    void close_fd() {
        if (fd_ >= 0)
            ::close(fd_);
        fd_ = -1;
    }
    int fd_ = -1;
};

} // namespace interprocess
} // namespace boost
