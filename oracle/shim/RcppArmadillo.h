// oracle/shim/RcppArmadillo.h -- TEST INFRASTRUCTURE, not product code.
//
// A small, eager, allocation-happy stand-in for the slice of (Rcpp)Armadillo
// that the reference headers under
//   /root/reference/src/single_group/src/cpp/{singleGroup.h,model/,algorithms/,misc/,rng/}
// touch.  R, Rcpp and Armadillo are not installed in this image, so the
// reference's own C++ cannot be built as shipped; with this header on the
// include path the reference headers compile UNMODIFIED (they are #include'd
// from /root/reference, never copied) and give us a runnable CPU reference
// (oracle/_ref/).  Semantics that carry meaning on the hot path:
//   * uword is 64-bit unsigned, matrices are column-major,
//   * resize(n) preserves and zero-fills, set_size does not initialise,
//   * sort_index(v,"descend") is std::sort on (value,index) pairs with a
//     value-only comparator (Armadillo's non-stable variant),
//   * find_finite keeps indices with std::isfinite,
//   * sum(umat-vector) is an integer count,
//   * randu()/randn() draw from injectable std::function hooks
//     (arma::g_randu / arma::g_randn) defined by the driver TU.
// Everything is written from the published Armadillo API documentation; no
// Armadillo source was consulted.
#ifndef HYG_ORACLE_SHIM_RCPPARMADILLO_H
#define HYG_ORACLE_SHIM_RCPPARMADILLO_H

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <iostream>
#include <limits>
#include <numeric>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace arma {

typedef unsigned long long uword;

extern std::function<double()> g_randu;  // defined by the driver
extern std::function<double()> g_randn;  // defined by the driver

struct span {
  uword a, b;
  span(uword a_, uword b_) : a(a_), b(b_) {}
};

namespace fill {
struct fill_zeros {};
struct fill_ones {};
static const fill_zeros zeros = fill_zeros();
static const fill_ones ones = fill_ones();
}  // namespace fill

struct SizeMat {
  uword n_rows, n_cols;
  SizeMat(uword r, uword c) : n_rows(r), n_cols(c) {}
};

template <class T> class subview;

template <class T> class Mat {
 public:
  typedef T elem_type;
  uword n_rows, n_cols, n_elem;
  std::vector<T> mem;

  Mat() : n_rows(0), n_cols(0), n_elem(0) {}
  explicit Mat(uword n) : n_rows(n), n_cols(1), n_elem(n), mem(n) {}
  Mat(uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), mem(r * c) {}
  Mat(uword r, uword c, fill::fill_zeros) : n_rows(r), n_cols(c), n_elem(r * c), mem(r * c, T(0)) {}
  Mat(uword r, uword c, fill::fill_ones) : n_rows(r), n_cols(c), n_elem(r * c), mem(r * c, T(1)) {}
  template <class U> Mat(const Mat<U>& o) : n_rows(o.n_rows), n_cols(o.n_cols), n_elem(o.n_elem), mem(o.n_elem) {
    for (uword i = 0; i < n_elem; i++) mem[i] = static_cast<T>(o.mem[i]);
  }
  virtual ~Mat() {}

  void init(uword r, uword c) { n_rows = r; n_cols = c; n_elem = r * c; mem.resize(n_elem); }

  // element access
  T& operator()(uword i) { return mem[i]; }
  const T& operator()(uword i) const { return mem[i]; }
  T& operator[](uword i) { return mem[i]; }
  const T& operator[](uword i) const { return mem[i]; }
  T& operator()(uword i, uword j) { return mem[i + j * n_rows]; }
  const T& operator()(uword i, uword j) const { return mem[i + j * n_rows]; }
  T& at(uword i) { return mem[i]; }
  const T& at(uword i) const { return mem[i]; }

  // sub-views (write-back proxies on non-const objects, copies on const ones)
  subview<T> operator()(const span& s);
  Mat<T> operator()(const span& s) const;
  subview<T> subvec(uword a, uword b);
  Mat<T> subvec(uword a, uword b) const;
  subview<T> col(uword j);
  Mat<T> col(uword j) const;
  subview<T> elem(const Mat<uword>& idx);
  Mat<T> elem(const Mat<uword>& idx) const;
  subview<T> diag();
  Mat<T> diag() const;

  uword size() const { return n_elem; }
  bool is_empty() const { return n_elem == 0; }
  void set_size(uword n) { init(n, 1); }
  void set_size(uword r, uword c) { init(r, c); }
  void resize(uword n) {  // preserving, zero-filling
    std::vector<T> old = mem;
    n_rows = n; n_cols = 1; n_elem = n;
    mem.assign(n, T(0));
    for (uword i = 0; i < std::min<uword>(n, old.size()); i++) mem[i] = old[i];
  }
  void zeros() { std::fill(mem.begin(), mem.end(), T(0)); }
  void zeros(uword n) { init(n, 1); zeros(); }
  void zeros(uword r, uword c) { init(r, c); zeros(); }
  void ones() { std::fill(mem.begin(), mem.end(), T(1)); }
  void ones(uword n) { init(n, 1); ones(); }
  void fill(T v) { std::fill(mem.begin(), mem.end(), v); }
  bool has_nan() const {
    for (uword i = 0; i < n_elem; i++) if (std::isnan(static_cast<double>(mem[i]))) return true;
    return false;
  }
  Mat<T> t() const {
    Mat<T> out(n_cols, n_rows);
    for (uword i = 0; i < n_rows; i++) for (uword j = 0; j < n_cols; j++) out(j, i) = (*this)(i, j);
    return out;
  }
  void insert_rows(uword row, uword n) {  // vectors / matrices, zero-filled
    Mat<T> out(n_rows + n, n_cols);
    for (uword j = 0; j < n_cols; j++) {
      for (uword i = 0; i < row; i++) out(i, j) = (*this)(i, j);
      for (uword i = 0; i < n; i++) out(row + i, j) = T(0);
      for (uword i = row; i < n_rows; i++) out(i + n, j) = (*this)(i, j);
    }
    n_rows = out.n_rows; n_elem = out.n_elem; mem.swap(out.mem);
  }
  void shed_row(uword row) {
    Mat<T> out(n_rows - 1, n_cols);
    for (uword j = 0; j < n_cols; j++) {
      uword k = 0;
      for (uword i = 0; i < n_rows; i++) if (i != row) out(k++, j) = (*this)(i, j);
    }
    n_rows = out.n_rows; n_elem = out.n_elem; mem.swap(out.mem);
  }
  void shed_rows(const Mat<uword>& rows) {
    std::vector<char> drop(n_rows, 0);
    for (uword i = 0; i < rows.n_elem; i++) drop[rows.mem[i]] = 1;
    uword keep = 0;
    for (uword i = 0; i < n_rows; i++) keep += !drop[i];
    Mat<T> out(keep, n_cols);
    for (uword j = 0; j < n_cols; j++) {
      uword k = 0;
      for (uword i = 0; i < n_rows; i++) if (!drop[i]) out(k++, j) = (*this)(i, j);
    }
    n_rows = out.n_rows; n_elem = out.n_elem; mem.swap(out.mem);
  }
  template <class F> Mat<T>& transform(F f) {
    for (uword i = 0; i < n_elem; i++) mem[i] = f(mem[i]);
    return *this;
  }
  void swap(Mat<T>& o) {
    std::swap(n_rows, o.n_rows); std::swap(n_cols, o.n_cols); std::swap(n_elem, o.n_elem); mem.swap(o.mem);
  }
};

template <class T> class Col : public Mat<T> {
 public:
  Col() : Mat<T>() {}
  explicit Col(uword n) : Mat<T>(n, 1) {}
  Col(uword n, fill::fill_zeros z) : Mat<T>(n, 1, z) {}
  Col(uword n, fill::fill_ones o) : Mat<T>(n, 1, o) {}
  template <class U> Col(const Mat<U>& o) : Mat<T>(o) { this->n_rows = this->n_elem; this->n_cols = 1; }
  template <class U> Col& operator=(const Mat<U>& o) {
    Mat<T> tmp(o);
    this->mem.swap(tmp.mem); this->n_elem = tmp.n_elem; this->n_rows = tmp.n_elem; this->n_cols = 1;
    return *this;
  }
  using Mat<T>::operator();
};

template <class T> class Row : public Mat<T> {
 public:
  Row() : Mat<T>() {}
  explicit Row(uword n) : Mat<T>(1, n) {}
  template <class U> Row(const Mat<U>& o) : Mat<T>(o) { this->n_cols = this->n_elem; this->n_rows = 1; }
  template <class U> Row& operator=(const Mat<U>& o) {
    Mat<T> tmp(o);
    this->mem.swap(tmp.mem); this->n_elem = tmp.n_elem; this->n_cols = tmp.n_elem; this->n_rows = 1;
    return *this;
  }
  using Mat<T>::operator();
};

typedef Mat<double> mat;
typedef Mat<uword> umat;
typedef Col<double> colvec;
typedef Col<double> vec;
typedef Row<double> rowvec;
typedef Col<uword> uvec;
typedef Col<uword> ucolvec;
typedef Row<uword> urowvec;

// Eager sub-view: a copy of the selected elements that knows where they came
// from; assignment / fill write through to the parent.
template <class T> class subview : public Mat<T> {
 public:
  Mat<T>* parent;
  std::vector<uword> map;
  subview(Mat<T>* p, const std::vector<uword>& m, uword r, uword c) : Mat<T>(r, c), parent(p), map(m) {
    for (uword i = 0; i < map.size(); i++) this->mem[i] = p->mem[map[i]];
  }
  void push() { for (uword i = 0; i < map.size(); i++) parent->mem[map[i]] = this->mem[i]; }
  template <class U> subview& operator=(const Mat<U>& o) {
    for (uword i = 0; i < map.size(); i++) this->mem[i] = static_cast<T>(o.mem[i]);
    push();
    return *this;
  }
  subview& operator=(const subview& o) {
    for (uword i = 0; i < map.size(); i++) this->mem[i] = o.mem[i];
    push();
    return *this;
  }
  void fill(T v) { std::fill(this->mem.begin(), this->mem.end(), v); push(); }
  void zeros() { fill(T(0)); }
  void ones() { fill(T(1)); }
};

template <class T> static std::vector<uword> hyg_range(uword a, uword b) {
  std::vector<uword> m;
  for (uword i = a; i <= b && b != static_cast<uword>(-1); i++) m.push_back(i);
  return m;
}
template <class T> subview<T> Mat<T>::operator()(const span& s) {
  std::vector<uword> m = hyg_range<T>(s.a, s.b);
  return subview<T>(this, m, m.size(), 1);
}
template <class T> Mat<T> Mat<T>::operator()(const span& s) const {
  std::vector<uword> m = hyg_range<T>(s.a, s.b);
  Mat<T> out(m.size(), 1);
  for (uword i = 0; i < m.size(); i++) out.mem[i] = mem[m[i]];
  return out;
}
template <class T> subview<T> Mat<T>::subvec(uword a, uword b) { return (*this)(span(a, b)); }
template <class T> Mat<T> Mat<T>::subvec(uword a, uword b) const { return (*this)(span(a, b)); }
template <class T> subview<T> Mat<T>::col(uword j) {
  std::vector<uword> m = hyg_range<T>(j * n_rows, (j + 1) * n_rows - 1);
  return subview<T>(this, m, n_rows, 1);
}
template <class T> Mat<T> Mat<T>::col(uword j) const {
  Mat<T> out(n_rows, 1);
  for (uword i = 0; i < n_rows; i++) out.mem[i] = mem[i + j * n_rows];
  return out;
}
template <class T> subview<T> Mat<T>::elem(const Mat<uword>& idx) {
  std::vector<uword> m(idx.mem.begin(), idx.mem.end());
  return subview<T>(this, m, m.size(), 1);
}
template <class T> Mat<T> Mat<T>::elem(const Mat<uword>& idx) const {
  Mat<T> out(idx.n_elem, 1);
  for (uword i = 0; i < idx.n_elem; i++) out.mem[i] = mem[idx.mem[i]];
  return out;
}
template <class T> subview<T> Mat<T>::diag() {
  std::vector<uword> m;
  for (uword i = 0; i < std::min(n_rows, n_cols); i++) m.push_back(i + i * n_rows);
  return subview<T>(this, m, m.size(), 1);
}
template <class T> Mat<T> Mat<T>::diag() const {
  uword n = std::min(n_rows, n_cols);
  Mat<T> out(n, 1);
  for (uword i = 0; i < n; i++) out.mem[i] = mem[i + i * n_rows];
  return out;
}

// ---------------------------------------------------------------------------
// element-wise operators (mixed element types promote like C++ arithmetic)
// ---------------------------------------------------------------------------
template <class A, class B> struct hyg_ct { typedef typename std::common_type<A, B>::type type; };

#define HYG_ELEMWISE(OP)                                                                          \
  template <class A, class B>                                                                     \
  Mat<typename hyg_ct<A, B>::type> operator OP(const Mat<A>& x, const Mat<B>& y) {                \
    typedef typename hyg_ct<A, B>::type R;                                                        \
    Mat<R> out(x.n_rows, x.n_cols);                                                               \
    for (uword i = 0; i < x.n_elem; i++) out.mem[i] = static_cast<R>(x.mem[i]) OP static_cast<R>(y.mem[i]); \
    return out;                                                                                   \
  }                                                                                               \
  template <class A, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type> \
  Mat<typename hyg_ct<A, S>::type> operator OP(const Mat<A>& x, S s) {                            \
    typedef typename hyg_ct<A, S>::type R;                                                        \
    Mat<R> out(x.n_rows, x.n_cols);                                                               \
    for (uword i = 0; i < x.n_elem; i++) out.mem[i] = static_cast<R>(x.mem[i]) OP static_cast<R>(s); \
    return out;                                                                                   \
  }                                                                                               \
  template <class A, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type> \
  Mat<typename hyg_ct<A, S>::type> operator OP(S s, const Mat<A>& x) {                            \
    typedef typename hyg_ct<A, S>::type R;                                                        \
    Mat<R> out(x.n_rows, x.n_cols);                                                               \
    for (uword i = 0; i < x.n_elem; i++) out.mem[i] = static_cast<R>(s) OP static_cast<R>(x.mem[i]); \
    return out;                                                                                   \
  }
HYG_ELEMWISE(+)
HYG_ELEMWISE(-)
HYG_ELEMWISE(/)
#undef HYG_ELEMWISE

// Schur product
template <class A, class B> Mat<typename hyg_ct<A, B>::type> operator%(const Mat<A>& x, const Mat<B>& y) {
  typedef typename hyg_ct<A, B>::type R;
  Mat<R> out(x.n_rows, x.n_cols);
  for (uword i = 0; i < x.n_elem; i++) out.mem[i] = static_cast<R>(x.mem[i]) * static_cast<R>(y.mem[i]);
  return out;
}
// scalar scaling
template <class A, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Mat<typename hyg_ct<A, S>::type> operator*(const Mat<A>& x, S s) {
  typedef typename hyg_ct<A, S>::type R;
  Mat<R> out(x.n_rows, x.n_cols);
  for (uword i = 0; i < x.n_elem; i++) out.mem[i] = static_cast<R>(x.mem[i]) * static_cast<R>(s);
  return out;
}
template <class A, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Mat<typename hyg_ct<A, S>::type> operator*(S s, const Mat<A>& x) { return x * s; }
// true matrix product
template <class A, class B> Mat<typename hyg_ct<A, B>::type> operator*(const Mat<A>& x, const Mat<B>& y) {
  typedef typename hyg_ct<A, B>::type R;
  Mat<R> out(x.n_rows, y.n_cols, fill::zeros);
  for (uword i = 0; i < x.n_rows; i++)
    for (uword j = 0; j < y.n_cols; j++) {
      R acc = R(0);
      for (uword k = 0; k < x.n_cols; k++) acc += static_cast<R>(x(i, k)) * static_cast<R>(y(k, j));
      out(i, j) = acc;
    }
  return out;
}
template <class A> Mat<A> operator-(const Mat<A>& x) {
  Mat<A> out(x.n_rows, x.n_cols);
  for (uword i = 0; i < x.n_elem; i++) out.mem[i] = -x.mem[i];
  return out;
}

#define HYG_COMPARE(OP)                                                                    \
  template <class A, class B> umat operator OP(const Mat<A>& x, const Mat<B>& y) {         \
    umat out(x.n_rows, x.n_cols);                                                          \
    for (uword i = 0; i < x.n_elem; i++) out.mem[i] = (x.mem[i] OP y.mem[i]) ? 1 : 0;      \
    return out;                                                                            \
  }                                                                                        \
  template <class A, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type> \
  umat operator OP(const Mat<A>& x, S s) {                                                 \
    umat out(x.n_rows, x.n_cols);                                                          \
    for (uword i = 0; i < x.n_elem; i++) out.mem[i] = (x.mem[i] OP s) ? 1 : 0;             \
    return out;                                                                            \
  }
HYG_COMPARE(>)
HYG_COMPARE(>=)
HYG_COMPARE(<)
HYG_COMPARE(<=)
HYG_COMPARE(==)
HYG_COMPARE(!=)
#undef HYG_COMPARE

template <class T> std::ostream& operator<<(std::ostream& os, const Mat<T>& m) {
  for (uword i = 0; i < m.n_rows; i++) {
    for (uword j = 0; j < m.n_cols; j++) os << "  " << m(i, j);
    os << "\n";
  }
  return os;
}

// ---------------------------------------------------------------------------
// free functions
// ---------------------------------------------------------------------------
#define HYG_MAP(NAME, EXPR)                                        \
  template <class T> Mat<double> NAME(const Mat<T>& x) {           \
    Mat<double> out(x.n_rows, x.n_cols);                           \
    for (uword i = 0; i < x.n_elem; i++) {                         \
      double v = static_cast<double>(x.mem[i]);                    \
      out.mem[i] = (EXPR);                                         \
    }                                                              \
    return out;                                                    \
  }
HYG_MAP(exp, std::exp(v))
HYG_MAP(log, std::log(v))
HYG_MAP(sqrt, std::sqrt(v))
HYG_MAP(ceil, std::ceil(v))
HYG_MAP(lgamma, std::lgamma(v))
#undef HYG_MAP
template <class T> Mat<double> pow(const Mat<T>& x, double p) {
  Mat<double> out(x.n_rows, x.n_cols);
  for (uword i = 0; i < x.n_elem; i++) out.mem[i] = std::pow(static_cast<double>(x.mem[i]), p);
  return out;
}
template <class T> T accu(const Mat<T>& x) {
  T acc = T(0);
  for (uword i = 0; i < x.n_elem; i++) acc += x.mem[i];
  return acc;
}
template <class T> T sum(const Mat<T>& x) { return accu(x); }  // vectors only on this path
template <class T> T max(const Mat<T>& x) {
  T m = x.mem[0];
  for (uword i = 1; i < x.n_elem; i++) if (x.mem[i] > m) m = x.mem[i];
  return m;
}
template <class T> Mat<T> cumsum(const Mat<T>& x) {
  Mat<T> out(x.n_rows, x.n_cols);
  T acc = T(0);
  for (uword i = 0; i < x.n_elem; i++) { acc += x.mem[i]; out.mem[i] = acc; }
  return out;
}
template <class T> Mat<T> reverse(const Mat<T>& x) {
  Mat<T> out(x.n_rows, x.n_cols);
  for (uword i = 0; i < x.n_elem; i++) out.mem[i] = x.mem[x.n_elem - 1 - i];
  return out;
}
template <class T> Mat<T> trans(const Mat<T>& x) { return x.t(); }
template <class T> Mat<T> normalise(const Mat<T>& x, int p) {
  double nrm = 0.0;
  for (uword i = 0; i < x.n_elem; i++) nrm += (p == 1) ? std::fabs(x.mem[i]) : x.mem[i] * x.mem[i];
  if (p != 1) nrm = std::sqrt(nrm);
  Mat<T> out(x.n_rows, x.n_cols);
  for (uword i = 0; i < x.n_elem; i++) out.mem[i] = (nrm > 0) ? x.mem[i] / nrm : x.mem[i];
  return out;
}
template <class T> uvec find(const Mat<T>& x) {
  std::vector<uword> v;
  for (uword i = 0; i < x.n_elem; i++) if (x.mem[i] != T(0)) v.push_back(i);
  uvec out(v.size());
  for (uword i = 0; i < v.size(); i++) out.mem[i] = v[i];
  return out;
}
template <class T> uvec find(const Mat<T>& x, uword k, const char* /*"first"*/) {
  std::vector<uword> v;
  for (uword i = 0; i < x.n_elem && v.size() < k; i++) if (x.mem[i] != T(0)) v.push_back(i);
  uvec out(v.size());
  for (uword i = 0; i < v.size(); i++) out.mem[i] = v[i];
  return out;
}
template <class T> uvec find_finite(const Mat<T>& x) {
  std::vector<uword> v;
  for (uword i = 0; i < x.n_elem; i++) if (std::isfinite(static_cast<double>(x.mem[i]))) v.push_back(i);
  uvec out(v.size());
  for (uword i = 0; i < v.size(); i++) out.mem[i] = v[i];
  return out;
}
template <class T> uvec sort_index(const Mat<T>& x, const char* dir) {
  std::vector<std::pair<T, uword> > v(x.n_elem);
  for (uword i = 0; i < x.n_elem; i++) v[i] = std::make_pair(x.mem[i], i);
  if (dir[0] == 'd')
    std::sort(v.begin(), v.end(), [](const std::pair<T, uword>& a, const std::pair<T, uword>& b) { return a.first > b.first; });
  else
    std::sort(v.begin(), v.end(), [](const std::pair<T, uword>& a, const std::pair<T, uword>& b) { return a.first < b.first; });
  uvec out(x.n_elem);
  for (uword i = 0; i < x.n_elem; i++) out.mem[i] = v[i].second;
  return out;
}
template <class Out> struct conv_to {
  template <class T> static Out from(const Mat<T>& x) { return hyg_conv(x, static_cast<Out*>(0)); }
};
template <class T> unsigned int hyg_conv(const Mat<T>& x, unsigned int*) { return static_cast<unsigned int>(x.mem[0]); }
template <class T> double hyg_conv(const Mat<T>& x, double*) { return static_cast<double>(x.mem[0]); }
template <class T, class U> Col<U> hyg_conv(const Mat<T>& x, Col<U>*) { return Col<U>(x); }
template <class T, class U> Mat<U> hyg_conv(const Mat<T>& x, Mat<U>*) { return Mat<U>(x); }

template <class V = colvec> V linspace(double a, double b, uword n) {
  // Armadillo: delta = (end-start)/(N-1); x[i] = start + i*delta.  Unit-step grids (the only ones on this path)
  // are produced with exact integer steps so that -ffast-math (reciprocal-math) cannot turn i into i - eps.
  V out(n);
  const bool unit = (n > 1) && ((b - a) == static_cast<double>(n - 1));
  const double delta = (n > 1) ? (b - a) / static_cast<double>(n - 1) : 0.0;
  for (uword i = 0; i < n; i++) {
    double v = unit ? a + static_cast<double>(i) : ((n > 1) ? a + static_cast<double>(i) * delta : b);
    out.mem[i] = static_cast<typename V::elem_type>(v);
  }
  return out;
}
template <class V = colvec> V ones(uword n) { V out(n); out.ones(); return out; }
template <class V = colvec> V zeros(uword n) { V out(n); out.zeros(); return out; }
inline mat ones(const SizeMat& s) { return mat(s.n_rows, s.n_cols, fill::ones); }
template <class T> SizeMat size(const Mat<T>& x) { return SizeMat(x.n_rows, x.n_cols); }
inline SizeMat size(uword r, uword c) { return SizeMat(r, c); }
inline double randu() { return g_randu(); }
template <class V = colvec> V randn(uword n) { V out(n); for (uword i = 0; i < n; i++) out.mem[i] = g_randn(); return out; }
template <class T> Mat<T> repmat(const Mat<T>& x, uword r, uword c) {
  Mat<T> out(x.n_rows * r, x.n_cols * c);
  for (uword i = 0; i < out.n_rows; i++) for (uword j = 0; j < out.n_cols; j++) out(i, j) = x(i % x.n_rows, j % x.n_cols);
  return out;
}
template <class T> Mat<T> diagmat(const Mat<T>& x) {
  Mat<T> out(x.n_elem, x.n_elem, fill::zeros);
  for (uword i = 0; i < x.n_elem; i++) out(i, i) = x.mem[i];
  return out;
}
template <class T> Mat<T> reshape(const Mat<T>& x, uword r, uword c) {
  Mat<T> out(r, c, fill::zeros);
  for (uword i = 0; i < std::min(out.n_elem, x.n_elem); i++) out.mem[i] = x.mem[i];
  return out;
}
// dead code on the hot path that must still type-check
inline mat hyg_dead(const char* what) { std::cerr << "oracle shim: " << what << " is not implemented" << std::endl; std::abort(); return mat(); }
template <class T> mat chol(const Mat<T>&) { return hyg_dead("chol"); }
template <class T> mat inv(const Mat<T>&) { return hyg_dead("inv"); }
template <class T> Mat<T> trimatu(const Mat<T>& x) { return x; }
inline uvec sub2ind(const SizeMat& s, const umat& sub) {
  uvec out(sub.n_cols);
  for (uword k = 0; k < sub.n_cols; k++) out.mem[k] = sub(0, k) + sub(1, k) * s.n_rows;
  return out;
}
struct arma_rng { static void set_seed(int) {} };

}  // namespace arma

namespace R {
inline double digamma(double x) {  // asymptotic series after upward recurrence
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  double f = 1.0 / (x * x);
  return r + std::log(x) - 0.5 / x - f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f * (1.0 / 132)))));
}
inline double hyg_dead(const char* what) { std::cerr << "oracle shim: R::" << what << " is not implemented" << std::endl; std::abort(); return 0.0; }
inline double pnorm(double, double, double, bool, bool) { return hyg_dead("pnorm"); }
inline double dnorm(double, double, double, bool) { return hyg_dead("dnorm"); }
inline double qnorm(double, double, double, bool, bool) { return hyg_dead("qnorm"); }
}  // namespace R

#endif
