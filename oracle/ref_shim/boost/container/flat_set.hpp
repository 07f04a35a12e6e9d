// TEST INFRASTRUCTURE.  Boost is not installed in this image.  The reference's Laia schedulers
// (laia/include/utils.h:7,48, laia/src/laia_scheduler.cc, laia/src/topk_scheduler.cc) use
// boost::container::flat_set<uint64_t> as an ordered set of keys: emplace / insert / clear /
// find / size / begin / end / erase(key) and construction from an iterator range.
//
// This stand-in keeps boost's MEMORY BEHAVIOUR, not only its set semantics: a sorted contiguous
// array, iterators that are plain pointers, and an erase that moves the tail down and shrinks the
// size without touching the vacated cells.  topk_scheduler.cc:478-482 erases from the set inside a
// range-for over it; with pointer iterators and a cached end() that loop visits every ORIGINAL cell
// (skipping the element that slides into an erased cell, re-reading stale copies at the tail), and
// the plan it produces depends on exactly that.  (A node-based std::set would free the node under
// the running iterator.)
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstring>
#include <type_traits>
#include <utility>

namespace boost {
namespace container {

template <class T>
class flat_set {
    static_assert(std::is_trivially_copyable<T>::value, "stand-in for trivially copyable keys only");

public:
    using value_type = T;
    using key_type = T;
    using size_type = std::size_t;
    using iterator = T *;
    using const_iterator = const T *;

    flat_set() = default;
    template <class It>
    flat_set(It first, It last) {
        for (; first != last; ++first)
            emplace(*first);
    }
    flat_set(const flat_set &o) {
        assign(o);
    }
    flat_set(flat_set &&o) noexcept : d_(o.d_), n_(o.n_), cap_(o.cap_) {
        o.d_ = nullptr;
        o.n_ = o.cap_ = 0;
    }
    flat_set &operator=(const flat_set &o) {
        if (this != &o)
            assign(o);
        return *this;
    }
    flat_set &operator=(flat_set &&o) noexcept {
        if (this != &o) {
            delete[] d_;
            d_ = o.d_;
            n_ = o.n_;
            cap_ = o.cap_;
            o.d_ = nullptr;
            o.n_ = o.cap_ = 0;
        }
        return *this;
    }
    ~flat_set() {
        delete[] d_;
    }

    iterator begin() { return d_; }
    iterator end() { return d_ + n_; }
    const_iterator begin() const { return d_; }
    const_iterator end() const { return d_ + n_; }
    size_type size() const { return n_; }
    bool empty() const { return n_ == 0; }
    void clear() { n_ = 0; }
    void reserve(size_type c) { grow(c); }

    std::pair<iterator, bool> emplace(const T &v) {
        T *p = std::lower_bound(d_, d_ + n_, v);
        if (p != d_ + n_ && *p == v)
            return {p, false};
        const size_type at = p - d_;
        if (n_ == cap_)
            grow(cap_ ? 2 * cap_ : 16);
        std::memmove(d_ + at + 1, d_ + at, (n_ - at) * sizeof(T));
        d_[at] = v;
        n_++;
        return {d_ + at, true};
    }
    std::pair<iterator, bool> insert(const T &v) { return emplace(v); }
    template <class It>
    void insert(It first, It last) {
        for (; first != last; ++first)
            emplace(*first);
    }
    iterator find(const T &v) {
        T *p = std::lower_bound(d_, d_ + n_, v);
        return (p != d_ + n_ && *p == v) ? p : d_ + n_;
    }
    const_iterator find(const T &v) const {
        const T *p = std::lower_bound(d_, d_ + n_, v);
        return (p != d_ + n_ && *p == v) ? p : d_ + n_;
    }
    size_type count(const T &v) const { return find(v) != end(); }
    // erase by key: the tail slides down, the vacated cell keeps its old contents
    size_type erase(const T &key) {
        const T v = key; // `key` may alias a cell that is about to be overwritten
        T *p = std::lower_bound(d_, d_ + n_, v);
        if (p == d_ + n_ || !(*p == v))
            return 0;
        std::memmove(p, p + 1, (d_ + n_ - (p + 1)) * sizeof(T));
        n_--;
        return 1;
    }

private:
    void grow(size_type c) {
        if (c <= cap_)
            return;
        T *q = new T[c];
        if (n_)
            std::memcpy(q, d_, n_ * sizeof(T));
        delete[] d_;
        d_ = q;
        cap_ = c;
    }
    void assign(const flat_set &o) {
        n_ = 0;
        grow(o.n_);
        if (o.n_)
            std::memcpy(d_, o.d_, o.n_ * sizeof(T));
        n_ = o.n_;
    }
    T *d_ = nullptr;
    size_type n_ = 0, cap_ = 0;
};

} // namespace container
} // namespace boost
