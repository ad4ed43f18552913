// TEST INFRASTRUCTURE (oracle/): a small value-semantics stand-in for the part of Armadillo that the reference's solver layer
// uses, so that the reference's OWN functions -- getDiagOfSigma, getCrossprod, getPCG1ofSigmaAndVector, getCoefficients,
// GetTrace[_q], getAIScore[_q], fitglmmaiRPCG[_q], getSigma_X / _G, calCV and the _LOCO twins
// (/root/reference/src/SAIGE/src/SAIGE_fitGLMM_fast.cpp) -- can be compiled from where they lie (oracle/ref_fg/extract_ref.py
// cuts them out at build time into oracle/_ref/, never into the repository) and used to pin the restatement in oracle/oracle.py.
// Armadillo itself is not installed here.  Only what those functions touch is provided: dense column vectors and matrices,
// element-wise and matrix arithmetic, dot / sum / mean / stddev, inv_sympd / pinv / solve for the p x p covariance algebra.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <limits>
#include <stdexcept>
#include <type_traits>
#include <vector>

namespace arma {
typedef unsigned long long uword;
#define ARMA_ARITH(S) typename = typename std::enable_if<std::is_arithmetic<S>::value>::type
#define ARMA_INTEG(S) typename = typename std::enable_if<std::is_integral<S>::value>::type

template <typename T> struct Col;
template <typename T> struct Mat;
namespace fill { struct zeros_t {}; static const zeros_t zeros = zeros_t(); }
struct distr_param { long long a, b; distr_param(long long a_, long long b_) : a(a_), b(b_) {} };

// s * v, not yet evaluated: Armadillo's expression templates fuse `acc += s * v` into one pass without a temporary, and the
// reference's marker loop (FG.cpp:1591) lives on exactly that statement -- a stand-in that allocated N floats per marker there
// would make the reference look slower than it is.  Everywhere else the product is simply materialised.
template <typename T> struct Scaled { T s; const Col<T> *v; };
template <typename T> struct RowView { const Col<T> *v; };            // v.t()
template <typename T> struct ElemView {                                // v.elem(idx)
    Col<T> *v; std::vector<uword> idx;
    void zeros() { for (uword i : idx) v->d[i] = T(0); }
};

template <typename T> struct Col {
    std::vector<T> d;
    uword n_elem = 0;
    Col() {}
    template <typename I, ARMA_INTEG(I)> explicit Col(I n) : d((size_t)n, T(0)), n_elem((uword)n) {}
    template <typename I, ARMA_INTEG(I)> Col(I n, fill::zeros_t) : d((size_t)n, T(0)), n_elem((uword)n) {}
    void zeros() { std::fill(d.begin(), d.end(), T(0)); }
    template <typename I, ARMA_INTEG(I)> void zeros(I n) { d.assign((size_t)n, T(0)); n_elem = (uword)n; }
    void clear() { d.clear(); n_elem = 0; }
    template <typename I> void resize(I n) { d.resize((size_t)n, T(0)); n_elem = (uword)n; }
    template <typename I> void set_size(I n) { d.resize((size_t)n); n_elem = (uword)n; }
    Col(const Scaled<T> &e) : d(e.v->n_elem), n_elem(e.v->n_elem) { for (uword i = 0; i < n_elem; i++) d[i] = e.s * e.v->d[i]; }
    Col &operator+=(const Scaled<T> &e)
    {
        if (e.v->n_elem != n_elem) throw std::logic_error("mini_arma: += on vectors of different length");
        const T s = e.s; const T *x = e.v->d.data(); T *y = d.data();
        for (uword i = 0; i < n_elem; i++) y[i] += s * x[i];
        return *this;
    }
    Col &operator+=(const Col &o) { if (o.n_elem != n_elem) throw std::logic_error("mini_arma: += on vectors of different length"); for (uword i = 0; i < n_elem; i++) d[i] += o.d[i]; return *this; }
    Col &operator-=(const Col &o) { if (o.n_elem != n_elem) throw std::logic_error("mini_arma: -= on vectors of different length"); for (uword i = 0; i < n_elem; i++) d[i] -= o.d[i]; return *this; }
    template <typename I> T &operator()(I i) { return d[(size_t)i]; }
    template <typename I> const T &operator()(I i) const { return d[(size_t)i]; }
    template <typename I> T &operator[](I i) { return d[(size_t)i]; }
    template <typename I> const T &operator[](I i) const { return d[(size_t)i]; }
    RowView<T> t() const { return RowView<T>{this}; }
    ElemView<T> elem(const Col<uword> &idx) { return ElemView<T>{this, idx.d}; }
    T *memptr() { return d.data(); }
    const T *memptr() const { return d.data(); }
};
typedef Col<uword> uvec;

template <typename T, typename F> Col<T> zip(const Col<T> &a, const Col<T> &b, F f)
{
    if (a.n_elem != b.n_elem) throw std::logic_error("mini_arma: element-wise operation on vectors of different length");
    Col<T> r(a.n_elem);
    for (uword i = 0; i < a.n_elem; i++) r.d[i] = f(a.d[i], b.d[i]);
    return r;
}
template <typename T, typename F> Col<T> map(const Col<T> &a, F f)
{
    Col<T> r(a.n_elem);
    for (uword i = 0; i < a.n_elem; i++) r.d[i] = f(a.d[i]);
    return r;
}
template <typename T> Col<T> operator+(const Col<T> &a, const Col<T> &b) { return zip(a, b, [](T x, T y) { return x + y; }); }
template <typename T> Col<T> operator-(const Col<T> &a, const Col<T> &b) { return zip(a, b, [](T x, T y) { return x - y; }); }
template <typename T> Col<T> operator%(const Col<T> &a, const Col<T> &b) { return zip(a, b, [](T x, T y) { return x * y; }); }
template <typename T> Col<T> operator/(const Col<T> &a, const Col<T> &b) { return zip(a, b, [](T x, T y) { return x / y; }); }
template <typename T, typename S, ARMA_ARITH(S)> Col<T> operator*(const Col<T> &a, S s) { return map(a, [s](T x) { return x * (T)s; }); }
template <typename T, typename S, ARMA_ARITH(S)> Scaled<T> operator*(S s, const Col<T> &a) { return Scaled<T>{(T)s, &a}; }
template <typename T> Col<T> operator+(const Col<T> &a, const Scaled<T> &b) { Col<T> r = a; r += b; return r; }
template <typename T> Col<T> operator-(const Col<T> &a, const Scaled<T> &b) { Col<T> r = a; r += Scaled<T>{-b.s, b.v}; return r; }
template <typename T> Col<T> operator+(const Scaled<T> &a, const Scaled<T> &b) { Col<T> r(a); r += b; return r; }
template <typename T> Col<T> operator+(const Scaled<T> &a, const Col<T> &b) { Col<T> r(a); r += b; return r; }
template <typename T> Col<T> operator/(const Scaled<T> &a, const Col<T> &b) { return Col<T>(a) / b; }
template <typename T, typename S, ARMA_ARITH(S)> Col<T> operator/(const Scaled<T> &a, S s) { return Col<T>(a) / s; }
template <typename T> Col<T> operator%(const Scaled<T> &a, const Col<T> &b) { return Col<T>(a) % b; }
template <typename T, typename S, ARMA_ARITH(S)> Col<T> operator/(const Col<T> &a, S s) { return map(a, [s](T x) { return x / (T)s; }); }
template <typename T, typename S, ARMA_ARITH(S)> Col<T> operator/(S s, const Col<T> &a) { return map(a, [s](T x) { return (T)s / x; }); }
template <typename T, typename S, ARMA_ARITH(S)> Col<T> operator+(const Col<T> &a, S s) { return map(a, [s](T x) { return x + (T)s; }); }
template <typename T, typename S, ARMA_ARITH(S)> Col<T> operator+(S s, const Col<T> &a) { return map(a, [s](T x) { return (T)s + x; }); }
template <typename T, typename S, ARMA_ARITH(S)> Col<T> operator-(const Col<T> &a, S s) { return map(a, [s](T x) { return x - (T)s; }); }
template <typename T, typename S, ARMA_ARITH(S)> uvec operator<(const Col<T> &a, S s)
{
    uvec r(a.n_elem);
    for (uword i = 0; i < a.n_elem; i++) r.d[i] = a.d[i] < (T)s ? 1 : 0;
    return r;
}
struct SizeMat { uword r, c; uword operator[](int i) const { return i == 0 ? r : c; } };
template <typename T> SizeMat size(const Col<T> &v) { return SizeMat{v.n_elem, 1}; }
template <typename T, typename S, ARMA_ARITH(S)> uvec operator==(const Col<T> &a, S s)
{
    uvec r(a.n_elem);
    for (uword i = 0; i < a.n_elem; i++) r.d[i] = a.d[i] == (T)s ? 1 : 0;
    return r;
}
inline bool any(const uvec &m) { for (uword x : m.d) if (x) return true; return false; }
template <typename T> Col<T> sort(Col<T> v) { std::sort(v.d.begin(), v.d.end()); return v; }
template <typename T> Col<T> unique(Col<T> v)
{
    std::sort(v.d.begin(), v.d.end());
    v.d.erase(std::unique(v.d.begin(), v.d.end()), v.d.end());
    v.n_elem = v.d.size();
    return v;
}
// arma::randi(n, distr_param(a, b)): the reference draws its variance-ratio hold-out candidates with it (FG.cpp:866-868).  A
// parity harness must see the SAME draw on every side, so the stand-in returns the vector its caller installed.
inline std::vector<long long> &randi_supply() { static std::vector<long long> v; return v; }
inline Col<long long> randi(long long n, const distr_param &)
{
    Col<long long> r;
    r.d = randi_supply();
    r.n_elem = r.d.size();
    (void)n;
    return r;
}
inline uvec find(const uvec &m)
{
    uvec r;
    for (uword i = 0; i < m.n_elem; i++) if (m.d[i]) r.d.push_back(i);
    r.n_elem = r.d.size();
    return r;
}
template <typename T> T dot(const Col<T> &a, const Col<T> &b)
{
    if (a.n_elem != b.n_elem) throw std::logic_error("mini_arma: dot of vectors of different length");
    T s = T(0);
    for (uword i = 0; i < a.n_elem; i++) s += a.d[i] * b.d[i];
    return s;
}
template <typename T> T sum(const Col<T> &a) { T s = T(0); for (T x : a.d) s += x; return s; }
template <typename T> T accu(const Col<T> &a) { return sum(a); }
template <typename T> T mean(const Col<T> &a) { return sum(a) / (T)a.n_elem; }
template <typename T> T stddev(const Col<T> &a)                       // normalised by n - 1, Armadillo's default
{
    const T m = mean(a);
    T s = T(0);
    for (T x : a.d) s += (x - m) * (x - m);
    return a.n_elem > 1 ? std::sqrt(s / (T)(a.n_elem - 1)) : T(0);
}
template <typename T> Col<T> operator*(const RowView<T> &r, const Col<T> &b) { Col<T> o(1); o.d[0] = dot(*r.v, b); return o; }

template <typename T> struct ColView {                                 // M.col(j): readable and assignable
    Mat<T> *m; uword j;
    operator Col<T>() const;
    ColView &operator=(const Col<T> &v);
};
template <typename T> struct Mat {
    std::vector<T> d;                                                  // column-major
    uword n_rows = 0, n_cols = 0, n_elem = 0;
    Mat() {}
    template <typename I, typename J, ARMA_INTEG(I)> Mat(I r, J c) : d((size_t)r * (size_t)c, T(0)), n_rows((uword)r), n_cols((uword)c), n_elem((uword)r * (uword)c) {}
    template <typename I, typename J> T &operator()(I i, J j) { return d[(size_t)i + (size_t)j * n_rows]; }
    template <typename I, typename J> const T &operator()(I i, J j) const { return d[(size_t)i + (size_t)j * n_rows]; }
    template <typename I> ColView<T> col(I j) { return ColView<T>{this, (uword)j}; }
    template <typename I, typename J> void zeros(I r, J c) { d.assign((size_t)r * (size_t)c, T(0)); n_rows = (uword)r; n_cols = (uword)c; n_elem = n_rows * n_cols; }
    Mat t() const
    {
        Mat r(n_cols, n_rows);
        for (uword i = 0; i < n_rows; i++) for (uword j = 0; j < n_cols; j++) r(j, i) = (*this)(i, j);
        return r;
    }
    T *memptr() { return d.data(); }
    const T *memptr() const { return d.data(); }
};
template <typename T> ColView<T>::operator Col<T>() const
{
    Col<T> r(m->n_rows);
    for (uword i = 0; i < m->n_rows; i++) r.d[i] = (*m)(i, j);
    return r;
}
template <typename T> ColView<T> &ColView<T>::operator=(const Col<T> &v)
{
    if (v.n_elem != m->n_rows) throw std::logic_error("mini_arma: column assignment of the wrong length");
    for (uword i = 0; i < m->n_rows; i++) (*m)(i, j) = v.d[i];
    return *this;
}
template <typename T> Col<T> operator+(const ColView<T> &a, const Col<T> &b) { return Col<T>(a) + b; }
template <typename T> Col<T> operator-(const Col<T> &a, const ColView<T> &b) { return a - Col<T>(b); }
template <typename T> Col<T> operator*(const Mat<T> &A, const Col<T> &x)
{
    if (A.n_cols != x.n_elem) throw std::logic_error("mini_arma: matrix * vector dimension mismatch");
    Col<T> r(A.n_rows);
    for (uword j = 0; j < A.n_cols; j++) { const T xj = x.d[j]; for (uword i = 0; i < A.n_rows; i++) r.d[i] += A(i, j) * xj; }
    return r;
}
template <typename T> Mat<T> operator*(const Mat<T> &A, const Mat<T> &B)
{
    if (A.n_cols != B.n_rows) throw std::logic_error("mini_arma: matrix * matrix dimension mismatch");
    Mat<T> r(A.n_rows, B.n_cols);
    for (uword j = 0; j < B.n_cols; j++)
        for (uword k = 0; k < A.n_cols; k++) { const T b = B(k, j); for (uword i = 0; i < A.n_rows; i++) r(i, j) += A(i, k) * b; }
    return r;
}
template <typename T> Mat<T> symmatu(const Mat<T> &A)                 // upper triangle reflected to the lower
{
    Mat<T> r = A;
    for (uword i = 0; i < A.n_rows; i++) for (uword j = 0; j < i; j++) r(i, j) = A(j, i);
    return r;
}
template <typename T> Mat<T> inv_sympd(const Mat<T> &A)               // Cholesky; throws like Armadillo when A is not SPD
{
    const uword p = A.n_rows;
    Mat<T> L = A;
    for (uword j = 0; j < p; j++) {
        T dj = L(j, j);
        for (uword q = 0; q < j; q++) dj -= L(j, q) * L(j, q);
        if (!(dj > T(0))) throw std::runtime_error("inv_sympd(): matrix is singular or not positive definite");
        dj = std::sqrt(dj);
        L(j, j) = dj;
        for (uword i = j + 1; i < p; i++) {
            T v = L(i, j);
            for (uword q = 0; q < j; q++) v -= L(i, q) * L(j, q);
            L(i, j) = v / dj;
        }
    }
    Mat<T> inv(p, p);
    for (uword c = 0; c < p; c++) {
        std::vector<T> y(p, T(0));
        for (uword i = 0; i < p; i++) {
            T v = (i == c) ? T(1) : T(0);
            for (uword q = 0; q < i; q++) v -= L(i, q) * y[q];
            y[i] = v / L(i, i);
        }
        for (uword ii = p; ii-- > 0;) {
            T v = y[ii];
            for (uword q = ii + 1; q < p; q++) v -= L(q, ii) * inv(q, c);
            inv(ii, c) = v / L(ii, ii);
        }
    }
    return inv;
}
template <typename T> Mat<T> pinv(const Mat<T> &A)                    // symmetric input only (that is all the reference feeds it)
{
    const uword p = A.n_rows;
    Mat<T> S = A, V(p, p);
    for (uword i = 0; i < p; i++) V(i, i) = T(1);
    for (int sweep = 0; sweep < 100; sweep++) {
        T off = T(0);
        for (uword i = 0; i < p; i++) for (uword j = 0; j < p; j++) if (i != j) off += S(i, j) * S(i, j);
        if (off < T(1e-30)) break;
        for (uword a = 0; a < p; a++)
            for (uword b = a + 1; b < p; b++) {
                const T apq = S(a, b);
                if (std::fabs(apq) < T(1e-300)) continue;
                const T th = (S(b, b) - S(a, a)) / (T(2) * apq);
                const T tt = (th >= 0 ? T(1) : T(-1)) / (std::fabs(th) + std::sqrt(th * th + T(1)));
                const T c = T(1) / std::sqrt(tt * tt + T(1)), s = tt * c;
                for (uword q = 0; q < p; q++) { const T x = S(q, a), y = S(q, b); S(q, a) = c * x - s * y; S(q, b) = s * x + c * y; }
                for (uword q = 0; q < p; q++) { const T x = S(a, q), y = S(b, q); S(a, q) = c * x - s * y; S(b, q) = s * x + c * y; }
                for (uword q = 0; q < p; q++) { const T x = V(q, a), y = V(q, b); V(q, a) = c * x - s * y; V(q, b) = s * x + c * y; }
            }
    }
    T smax = T(0);
    for (uword i = 0; i < p; i++) smax = std::max(smax, (T)std::fabs(S(i, i)));
    const T tol = (T)p * smax * std::numeric_limits<T>::epsilon();
    Mat<T> inv(p, p);
    for (uword q = 0; q < p; q++) {
        const T ev = S(q, q);
        if (std::fabs(ev) <= tol) continue;
        for (uword i = 0; i < p; i++) for (uword j = 0; j < p; j++) inv(i, j) += V(i, q) * V(j, q) / ev;
    }
    return inv;
}
template <typename T> Col<T> solve(Mat<T> A, Col<T> b)                // Gaussian elimination with partial pivoting
{
    const uword p = A.n_rows;
    for (uword c = 0; c < p; c++) {
        uword piv = c;
        for (uword r = c + 1; r < p; r++) if (std::fabs(A(r, c)) > std::fabs(A(piv, c))) piv = r;
        if (A(piv, c) == T(0)) throw std::runtime_error("solve(): solution not found");
        if (piv != c) { for (uword q = 0; q < p; q++) std::swap(A(c, q), A(piv, q)); std::swap(b.d[c], b.d[piv]); }
        for (uword r = c + 1; r < p; r++) {
            const T f = A(r, c) / A(c, c);
            for (uword q = c; q < p; q++) A(r, q) -= f * A(c, q);
            b.d[r] -= f * b.d[c];
        }
    }
    for (uword rr = p; rr-- > 0;) {
        T v = b.d[rr];
        for (uword q = rr + 1; q < p; q++) v -= A(rr, q) * b.d[q];
        b.d[rr] = v / A(rr, rr);
    }
    return b;
}
}  // namespace arma
