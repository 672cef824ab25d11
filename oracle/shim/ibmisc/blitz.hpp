// Stand-in for <ibmisc/blitz.hpp> (ibmisc and blitz++ are not vendored with the reference and are
// absent from this image).  TEST INFRASTRUCTURE ONLY: lets the genuine reference headers under
// /root/reference/slib compile so that oracle/_ref can serve as ground truth.
// Only what the reference touches is provided:
//   spsparse.hpp:78 (blitz::Array typedef), VectorCooArray.hpp:70-73,128,150,313-321,
//   accum.hpp:110-140 (DenseAccum copies the array it writes into => copies must SHARE storage),
//   algorithm.hpp:433-440 (to_sparse: iterator with position()).
#pragma once
#include <array>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <iostream>
#include <memory>
#include <vector>

namespace blitz {

template <class T, int N>
struct TinyVector {
    T d[N];
    T &operator[](int i) { return d[i]; }
    T const &operator[](int i) const { return d[i]; }
};

template <class T, int N>
class Array {
    std::shared_ptr<std::vector<T>> store_;  // shared on copy, like blitz reference counting
    int ext_[N];

    size_t flat(const int *ix) const {
        size_t o = 0;
        for (int k = 0; k < N; ++k) o = o * (size_t)ext_[k] + (size_t)ix[k];
        return o;
    }
    void alloc() {
        size_t n = 1;
        for (int k = 0; k < N; ++k) n *= (size_t)ext_[k];
        store_ = std::make_shared<std::vector<T>>(n, T());
    }

public:
    Array() { for (int k = 0; k < N; ++k) ext_[k] = 0; alloc(); }
    explicit Array(TinyVector<int, N> const &e) { for (int k = 0; k < N; ++k) ext_[k] = e[k]; alloc(); }
    explicit Array(int e0) { static_assert(N == 1, "rank"); ext_[0] = e0; alloc(); }
    Array(int e0, int e1) { static_assert(N == 2, "rank"); ext_[0] = e0; ext_[1] = e1; alloc(); }

    Array &operator=(T v) { for (auto &x : *store_) x = v; return *this; }
    int extent(int k) const { return ext_[k]; }
    size_t size() const { return store_->size(); }

    T &operator()(TinyVector<int, N> const &ix) { return (*store_)[flat(ix.d)]; }
    T const &operator()(TinyVector<int, N> const &ix) const { return (*store_)[flat(ix.d)]; }
    T &operator()(int i) { int ix[1] = {i}; return (*store_)[flat(ix)]; }
    T const &operator()(int i) const { int ix[1] = {i}; return (*store_)[flat(ix)]; }
    T &operator()(int i, int j) { int ix[2] = {i, j}; return (*store_)[flat(ix)]; }
    T const &operator()(int i, int j) const { int ix[2] = {i, j}; return (*store_)[flat(ix)]; }

    struct const_iterator {
        Array const *a;
        size_t off;
        T const &operator*() const { return (*a->store_)[off]; }
        const_iterator &operator++() { ++off; return *this; }
        bool operator!=(const_iterator const &o) const { return off != o.off; }
        bool operator==(const_iterator const &o) const { return off == o.off; }
        TinyVector<int, N> position() const {
            TinyVector<int, N> p;
            size_t r = off;
            for (int k = N - 1; k >= 0; --k) { p[k] = (int)(r % (size_t)a->ext_[k]); r /= (size_t)a->ext_[k]; }
            return p;
        }
    };
    const_iterator begin() const { return const_iterator{this, 0}; }
    const_iterator end() const { return const_iterator{this, store_->size()}; }
    T *data() { return store_->data(); }
    T const *data() const { return store_->data(); }
};

}  // namespace blitz

namespace ibmisc {

template <class T>
blitz::Array<T, 1> to_blitz(std::vector<T> const &v) {
    blitz::Array<T, 1> a((int)v.size());
    for (size_t i = 0; i < v.size(); ++i) a((int)i) = v[i];
    return a;
}

template <class T>
std::vector<T> to_vector(blitz::Array<T, 1> const &a) {
    std::vector<T> v;
    for (int i = 0; i < a.extent(0); ++i) v.push_back(a(i));
    return v;
}

template <class D, class S, int N>
blitz::TinyVector<D, N> to_tiny(std::array<S, N> const &s) {
    blitz::TinyVector<D, N> t;
    for (int k = 0; k < N; ++k) t[k] = (D)s[k];
    return t;
}

template <class D, class S, int N>
void to_tiny(blitz::TinyVector<D, N> &t, std::array<S, N> const &s) {
    for (int k = 0; k < N; ++k) t[k] = (D)s[k];
}

}  // namespace ibmisc
