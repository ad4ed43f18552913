// TEST INFRASTRUCTURE: a declaration-only stand-in for <RcppArmadillo.h>, just wide enough for
// `g++ -fsyntax-only -DUSE_SAIGE_B200 rcpp_shim/SAIGE_fitGLMM_fast_b200.cpp` (tests/test_abi.py).  R, Rcpp and Armadillo are not
// installed in the build container; this lets the compiler check every sgb_* call of the shim (names, arity, pointer types)
// against include/saige_b200.h.  Nothing here has a body worth running and nothing in the product includes it.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <initializer_list>
#include <iostream>
#include <string>
#include <type_traits>
#include <vector>

namespace arma {
typedef unsigned long long uword;
typedef long long sword;
namespace fill { struct fill_zeros {}; static const fill_zeros zeros = fill_zeros(); }
template <typename T> struct Mat {
    uword n_rows = 0, n_cols = 0, n_elem = 0;
    Mat() {}
    Mat(uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c) {}
    Mat(uword r, uword c, fill::fill_zeros) : n_rows(r), n_cols(c), n_elem(r * c) {}
    Mat(std::initializer_list<std::initializer_list<T> >) {}
    T *memptr() { return nullptr; }
    const T *memptr() const { return nullptr; }
    T *begin() { return nullptr; }
    T *end() { return nullptr; }
    T &operator()(uword, uword) { static T t; return t; }
};
template <typename T> struct Col : Mat<T> {
    Col() {}
    template <typename I, typename = typename std::enable_if<std::is_integral<I>::value>::type> Col(I n) : Mat<T>((uword)n, 1) {}
    template <typename I, typename = typename std::enable_if<std::is_integral<I>::value>::type> Col(I n, fill::fill_zeros z) : Mat<T>((uword)n, 1, z) {}
    Col(std::initializer_list<T>) {}
    T &operator[](uword) { static T t; return t; }
    T &operator()(uword) { static T t; return t; }
};
typedef Col<double> vec;
typedef Col<float> fvec;
typedef Col<float> fcolvec;
typedef Col<sword> ivec;
typedef Mat<double> mat;
typedef Mat<float> fmat;
template <typename Out> struct conv_to { template <typename In> static Out from(const In &) { return Out(); } };
template <typename T> Col<T> sqrt(const Col<T> &v) { return v; }
}  // namespace arma

namespace Rcpp {
struct SEXPish {};
struct NamedArg { template <typename T> NamedArg operator=(const T &) const { return *this; } };
inline NamedArg Named(const char *) { return NamedArg(); }
template <typename T> inline NamedArg Named(const char *, const T &) { return NamedArg(); }
struct List { template <typename... A> static List create(const A &...) { return List(); } };
[[noreturn]] inline void stop(const std::string &) { throw 1; }
template <typename T> struct Vector {
    std::vector<T> v;
    Vector() {}
    Vector(long long n) : v((size_t)n) {}
    Vector(const SEXPish &) {}
    T &operator[](long long i) { return v[(size_t)i]; }
    long long size() const { return (long long)v.size(); }
    typename std::vector<T>::iterator begin() { return v.begin(); }
    typename std::vector<T>::iterator end() { return v.end(); }
    Vector operator-(int) const { return *this; }
};
typedef Vector<double> NumericVector;
typedef Vector<int> IntegerVector;
typedef Vector<unsigned char> RawVector;
struct Function { template <typename... A> SEXPish operator()(const A &...) const { return SEXPish(); } };
struct Environment { Environment(const char *) {} Function operator[](const char *) const { return Function(); } };
template <typename T> T as(const SEXPish &) { return T(); }
inline NumericVector rbinom(long long n, double, double) { return NumericVector(n); }
inline IntegerVector sample(int, int, bool) { return IntegerVector(0); }
static std::ostream &Rcout = std::cout;
}  // namespace Rcpp
inline void Rprintf(const char *, ...) {}
