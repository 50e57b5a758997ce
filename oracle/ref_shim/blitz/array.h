// blitz/array.h -- SHIM, test infrastructure only (never linked into the product).
//
// The reference (pw359/Impec2D) typedefs blitz::Array<double,2/3> in src/GenSimDefs.h:23-31.  blitz++ is not installed in this image,
// so oracle/Makefile.ref compiles the reference's UNMODIFIED translation units against this header instead.  It implements exactly
// the blitz surface those files use, with blitz's documented semantics:
//   * Array(const Array&) is a REFERENCE copy (shares the data block); operator=(const Array&) copies elements; copy() deep-copies
//   * resize() keeps the index bases, reindexSelf() moves them, reverseSelf(d) flips the stride of dimension d in place
//   * operator()(i,j[,k]) indexes with the current bases; extent() returns the TinyVector of lengths
//   * whole-array expressions a + b, a - b (same shape, combined by position), a = scalar, max(a), min(a)
// New memory is zero-filled (blitz leaves it uninitialised, so any value is conforming).
#ifndef IMPEC_REF_SHIM_BLITZ_ARRAY_H
#define IMPEC_REF_SHIM_BLITZ_ARRAY_H
#include <memory>
#include <cstddef>
#include <cstring>
#include <cstdlib>
#include <cassert>
#include <algorithm>
// the real blitz/array.h drags these in, and the reference relies on it (e.g. numeric_limits and fabs in src/HyperbolicSolver.cpp)
#include <cmath>
#include <math.h>
#include <limits>
#include <iostream>
#include <fstream>
#include <string>

// The reference is C++03: src/Printer.cpp:131 streams a std::ofstream into cerr (pre-C++11 that printed its void* conversion).
// This overload keeps that single diagnostic line compiling under the C++11 this shim needs.
inline std::ostream &operator<<(std::ostream &o, const std::ofstream &f) { return o << (f.good() ? "<ofstream>" : "<bad ofstream>"); }

namespace blitz {

template <typename T, int N> class TinyVector {
    T v_[N];
public:
    TinyVector() { for (int k = 0; k < N; k++) v_[k] = T(); }
    TinyVector(T a) { for (int k = 0; k < N; k++) v_[k] = a; }
    T &operator[](int k) { return v_[k]; }
    const T &operator[](int k) const { return v_[k]; }
    T &operator()(int k) { return v_[k]; }
    const T &operator()(int k) const { return v_[k]; }
    TinyVector &operator=(T a) { for (int k = 0; k < N; k++) v_[k] = a; return *this; }
    int length() const { return N; }
};

template <typename T, int N> class Array {
    std::shared_ptr<T> block_;
    T *zero_ = nullptr;                  // address of element (0,...,0) in the current indexing
    TinyVector<int, N> base_, len_;
    std::ptrdiff_t stride_[N];

    void allocate() {
        std::size_t n = 1;
        for (int k = 0; k < N; k++) n *= (std::size_t)(len_[k] > 0 ? len_[k] : 0);
        block_.reset(n ? new T[n]() : nullptr, std::default_delete<T[]>());
        std::ptrdiff_t s = 1;
        for (int k = N - 1; k >= 0; k--) { stride_[k] = s; s *= len_[k]; }          // C storage order: last index fastest
        zero_ = block_.get();
        for (int k = 0; k < N; k++) zero_ -= (std::ptrdiff_t)base_[k] * stride_[k];
    }
    template <typename F> void for_each_index(F f) const {
        TinyVector<int, N> i;
        for (int k = 0; k < N; k++) { if (len_[k] <= 0) return; i[k] = base_[k]; }
        for (;;) {
            f(i);
            int k = N - 1;
            while (k >= 0 && ++i[k] >= base_[k] + len_[k]) { i[k] = base_[k]; k--; }
            if (k < 0) return;
        }
    }
    T &at(const TinyVector<int, N> &i) const { std::ptrdiff_t o = 0; for (int k = 0; k < N; k++) o += (std::ptrdiff_t)i[k] * stride_[k]; return zero_[o]; }
    // element at POSITION p (offsets from the lower bounds): what blitz's expression evaluation pairs up
    T &at_pos(const Array &ref, const TinyVector<int, N> &i) const {
        std::ptrdiff_t o = 0; for (int k = 0; k < N; k++) o += (std::ptrdiff_t)(i[k] - ref.base_[k] + base_[k]) * stride_[k]; return zero_[o];
    }
    template <typename Op> Array combine(const Array &b, Op op) const {
        Array r; r.base_ = base_; r.len_ = len_; r.allocate();
        for (int k = 0; k < N; k++) assert(len_[k] == b.len_[k]);
        for_each_index([&](const TinyVector<int, N> &i) { r.at(i) = op(at(i), b.at_pos(*this, i)); });
        return r;
    }
public:
    Array() { for (int k = 0; k < N; k++) { base_[k] = 0; len_[k] = 0; stride_[k] = 0; } }
    explicit Array(int n0) { static_assert(N == 1, "rank"); base_[0] = 0; len_[0] = n0; allocate(); }
    Array(int n0, int n1) { static_assert(N == 2, "rank"); base_[0] = base_[1] = 0; len_[0] = n0; len_[1] = n1; allocate(); }
    Array(int n0, int n1, int n2) { static_assert(N == 3, "rank"); base_[0] = base_[1] = base_[2] = 0; len_[0] = n0; len_[1] = n1; len_[2] = n2; allocate(); }
    Array(const Array &o) = default;                                   // reference semantics, as blitz
    Array &operator=(const Array &o) {                                  // element-wise copy into the existing storage
        if (&o == this) return *this;
        for (int k = 0; k < N; k++) assert(len_[k] == o.len_[k]);
        for_each_index([&](const TinyVector<int, N> &i) { at(i) = o.at_pos(*this, i); });
        return *this;
    }
    Array &operator=(T a) { for_each_index([&](const TinyVector<int, N> &i) { at(i) = a; }); return *this; }

    void resize(int n0) { static_assert(N == 1, "rank"); if (len_[0] == n0 && block_) return; len_[0] = n0; allocate(); }
    void resize(int n0, int n1) { static_assert(N == 2, "rank"); if (len_[0] == n0 && len_[1] == n1 && block_) return; len_[0] = n0; len_[1] = n1; allocate(); }
    void resize(int n0, int n1, int n2) {
        static_assert(N == 3, "rank");
        if (len_[0] == n0 && len_[1] == n1 && len_[2] == n2 && block_) return;
        len_[0] = n0; len_[1] = n1; len_[2] = n2; allocate();
    }
    void reindexSelf(const TinyVector<int, N> &nb) {
        for (int k = 0; k < N; k++) { zero_ += (std::ptrdiff_t)(base_[k] - nb[k]) * stride_[k]; base_[k] = nb[k]; }
    }
    void reverseSelf(int d) { zero_ += (std::ptrdiff_t)(len_[d] - 1 + 2 * base_[d]) * stride_[d]; stride_[d] = -stride_[d]; }
    Array copy() const {
        Array r; r.base_ = base_; r.len_ = len_; r.allocate();
        for_each_index([&](const TinyVector<int, N> &i) { r.at(i) = at(i); });
        return r;
    }

    T &operator()(int i) { static_assert(N == 1, "rank"); return zero_[(std::ptrdiff_t)i * stride_[0]]; }
    const T &operator()(int i) const { static_assert(N == 1, "rank"); return zero_[(std::ptrdiff_t)i * stride_[0]]; }
    T &operator()(int i, int j) { static_assert(N == 2, "rank"); return zero_[(std::ptrdiff_t)i * stride_[0] + (std::ptrdiff_t)j * stride_[1]]; }
    const T &operator()(int i, int j) const { static_assert(N == 2, "rank"); return zero_[(std::ptrdiff_t)i * stride_[0] + (std::ptrdiff_t)j * stride_[1]]; }
    T &operator()(int i, int j, int k) { static_assert(N == 3, "rank"); return zero_[(std::ptrdiff_t)i * stride_[0] + (std::ptrdiff_t)j * stride_[1] + (std::ptrdiff_t)k * stride_[2]]; }
    const T &operator()(int i, int j, int k) const { static_assert(N == 3, "rank"); return zero_[(std::ptrdiff_t)i * stride_[0] + (std::ptrdiff_t)j * stride_[1] + (std::ptrdiff_t)k * stride_[2]]; }

    const TinyVector<int, N> &extent() const { return len_; }
    const TinyVector<int, N> &shape() const { return len_; }
    const TinyVector<int, N> &length() const { return len_; }
    int extent(int d) const { return len_[d]; }
    const TinyVector<int, N> &base() const { return base_; }
    int lbound(int d) const { return base_[d]; }
    int ubound(int d) const { return base_[d] + len_[d] - 1; }
    int rows() const { return len_[0]; }
    int cols() const { return len_[1]; }
    std::size_t size() const { std::size_t n = 1; for (int k = 0; k < N; k++) n *= (std::size_t)len_[k]; return n; }
    std::size_t numElements() const { return size(); }
    T *data() { return block_.get(); }
    const T *data() const { return block_.get(); }

    Array operator+(const Array &b) const { return combine(b, [](T x, T y) { return x + y; }); }
    Array operator-(const Array &b) const { return combine(b, [](T x, T y) { return x - y; }); }

    T reduce_max() const { bool first = true; T m = T(); for_each_index([&](const TinyVector<int, N> &i) { const T v = at(i); if (first || v > m) m = v; first = false; }); return m; }
    T reduce_min() const { bool first = true; T m = T(); for_each_index([&](const TinyVector<int, N> &i) { const T v = at(i); if (first || v < m) m = v; first = false; }); return m; }
};

template <typename T, int N> T max(const Array<T, N> &a) { return a.reduce_max(); }
template <typename T, int N> T min(const Array<T, N> &a) { return a.reduce_min(); }

}  // namespace blitz
#endif
