// Minimal stand-in for the parts of <Rcpp.h> that r_shim/bvar_cpp_shim.cpp uses, so that the shim can be compiled
// and RUN in an image without R (tests/test_shim_cpp.py).  Test infrastructure only - nothing here ships.
//
// Same spellings and semantics as Rcpp for: NumericVector / IntegerVector / LogicalVector (size, begin, end,
// operator[], attr("dim"), ::create), NumericMatrix / IntegerMatrix (nrow, ncol, begin, column-major), List (named,
// operator[](name / index), ::create(Named(..) = ..)), Nullable<T>, Named, as<T>, stop, Rprintf, checkUserInterrupt,
// R::unif_rand / R::norm_rand, R_NilValue, R_xlen_t.  Vectors share storage on copy, as SEXP-backed Rcpp vectors do.
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

typedef long R_xlen_t;
struct SEXPREC_stub {};
static SEXPREC_stub* const R_NilValue = nullptr;

namespace Rcpp {

struct exception : std::runtime_error { using std::runtime_error::runtime_error; };
inline void stop(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw exception(buf);
}
inline void warning(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "Warning: ");
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
}
inline void checkUserInterrupt() {}

struct RObj;
typedef std::shared_ptr<RObj> Ptr;
struct RObj {
  enum Kind { NIL, REAL, INT, LGL, STR, LIST } kind = NIL;
  std::vector<double> d;
  std::vector<int> i;           // INT and LGL
  std::vector<std::string> s;
  std::vector<Ptr> l;
  std::vector<std::string> names;
  std::map<std::string, Ptr> attrs;
};
inline Ptr mk(RObj::Kind k) { auto p = std::make_shared<RObj>(); p->kind = k; return p; }

template <RObj::Kind KIND, typename T>
class Vec {
 public:
  Ptr p;
  Vec() : p(mk(KIND)) {}
  explicit Vec(Ptr q) : p(q ? q : mk(KIND)) {
    if (p->kind == RObj::NIL) p = mk(KIND);
    if (KIND == RObj::REAL && p->kind == RObj::INT) {        // R coerces integer -> double on the way in
      auto c = mk(RObj::REAL); c->d.assign(p->i.begin(), p->i.end()); c->attrs = p->attrs; p = c;
    }
    if ((KIND == RObj::INT || KIND == RObj::LGL) && (p->kind == RObj::INT || p->kind == RObj::LGL)) return;
    if (p->kind != KIND) throw exception("Rcpp stub: vector type mismatch");
  }
  Vec(size_t n) : p(mk(KIND)) { store().assign(n, T()); }
  Vec(int n) : Vec((size_t)(n < 0 ? 0 : n)) {}
  Vec(long n) : Vec((size_t)(n < 0 ? 0 : n)) {}
  template <typename It>
  Vec(It a, It b) : p(mk(KIND)) { store().assign(a, b); }
  std::vector<T>& store() const { return storage(p.get(), (T*)nullptr); }
  static std::vector<double>& storage(RObj* o, double*) { return o->d; }
  static std::vector<int>& storage(RObj* o, int*) { return o->i; }
  R_xlen_t size() const { return (R_xlen_t)store().size(); }
  R_xlen_t length() const { return size(); }
  T* begin() const { return store().data(); }
  T* end() const { return store().data() + store().size(); }
  T& operator[](R_xlen_t k) const { return store()[(size_t)k]; }
  struct AttrProxy {
    Ptr owner; std::string name;
    template <typename V> AttrProxy& operator=(const V& v) { owner->attrs[name] = v.p; return *this; }
    template <typename V> operator V() const {
      auto it = owner->attrs.find(name);
      return V(it == owner->attrs.end() ? Ptr() : it->second);
    }
  };
  AttrProxy attr(const std::string& n) const { return AttrProxy{p, n}; }
  static Vec create(T a) { Vec v((size_t)1); v[0] = a; return v; }
  static Vec create(T a, T b) { Vec v((size_t)2); v[0] = a; v[1] = b; return v; }
  static Vec create(T a, T b, T c) { Vec v((size_t)3); v[0] = a; v[1] = b; v[2] = c; return v; }
};
typedef Vec<RObj::REAL, double> NumericVector;
typedef Vec<RObj::INT, int> IntegerVector;
typedef Vec<RObj::LGL, int> LogicalVector;

template <RObj::Kind KIND, typename T>
class Mat : public Vec<KIND, T> {
 public:
  typedef Vec<KIND, T> Base;
  Mat() : Base() { setdim(0, 0); }
  explicit Mat(Ptr q) : Base(q) {
    if (!this->p->attrs.count("dim")) setdim((int)this->size(), this->size() ? 1 : 0);
  }
  Mat(int nr, int nc) : Base((size_t)((long)nr * nc)) { setdim(nr, nc); }
  void setdim(int nr, int nc) {
    auto d = mk(RObj::INT); d->i = {nr, nc};
    this->p->attrs["dim"] = d;
  }
  int nrow() const { return this->p->attrs.at("dim")->i[0]; }
  int ncol() const { return this->p->attrs.at("dim")->i[1]; }
  T& operator()(int r, int c) const { return this->store()[(size_t)c * nrow() + r]; }
};
typedef Mat<RObj::REAL, double> NumericMatrix;
typedef Mat<RObj::INT, int> IntegerMatrix;

struct Scalar {   // what List["key"] can turn into besides vectors
  Ptr p;
};

class List;
struct Named {
  std::string name; Ptr value;
  explicit Named(const std::string& n) : name(n) {}
  template <typename V> Named& operator=(const V& v) { value = v.p; return *this; }
};

class List {
 public:
  Ptr p;
  List() : p(mk(RObj::LIST)) {}
  explicit List(Ptr q) : p(q ? q : mk(RObj::LIST)) { if (p->kind != RObj::LIST) throw exception("Rcpp stub: not a list"); }
  explicit List(int n) : p(mk(RObj::LIST)) { p->l.assign((size_t)n, Ptr()); }
  R_xlen_t size() const { return (R_xlen_t)p->l.size(); }
  struct Proxy {
    Ptr owner; long idx; std::string name;
    Ptr get() const {
      if (idx >= 0) return owner->l.at((size_t)idx);
      for (size_t k = 0; k < owner->names.size(); ++k) if (owner->names[k] == name) return owner->l[k];
      throw exception(("Rcpp stub: no list element named '" + name + "'").c_str());
    }
    template <typename V> Proxy& operator=(const V& v) { owner->l.at((size_t)idx) = v.p; return *this; }
    operator NumericVector() const { return NumericVector(get()); }
    operator IntegerVector() const { return IntegerVector(get()); }
    operator LogicalVector() const { return LogicalVector(get()); }
    operator NumericMatrix() const { return NumericMatrix(get()); }
    operator IntegerMatrix() const { return IntegerMatrix(get()); }
    operator List() const { return List(get()); }
    operator int() const { Ptr q = get(); return q->kind == RObj::REAL ? (int)q->d.at(0) : q->i.at(0); }
    operator double() const { Ptr q = get(); return q->kind == RObj::REAL ? q->d.at(0) : (double)q->i.at(0); }
    operator bool() const { Ptr q = get(); return q->kind == RObj::REAL ? q->d.at(0) != 0 : q->i.at(0) != 0; }
    operator std::string() const { return get()->s.at(0); }
  };
  Proxy operator[](const char* n) const { return Proxy{p, -1, n}; }
  Proxy operator[](const std::string& n) const { return Proxy{p, -1, n}; }
  Proxy operator[](int k) const { return Proxy{p, k, ""}; }
  void push(const Named& a) { p->l.push_back(a.value); p->names.push_back(a.name); }
  template <typename... A>
  static List create(const A&... a) { List L; (L.push(a), ...); return L; }
};

template <typename T> T as(const List::Proxy& x) { return (T)x; }

template <typename T>
class Nullable {
  bool has = false; T v;
 public:
  Nullable() {}
  Nullable(SEXPREC_stub*) {}
  Nullable(const T& x) : has(true), v(x) {}
  bool isNotNull() const { return has; }
  bool isNull() const { return !has; }
  T get() const { return v; }
};

}  // namespace Rcpp

inline void Rprintf(const char* fmt, ...) {
  if (!std::getenv("RSTUB_VERBOSE")) return;
  va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap);
}

namespace R {
inline std::mt19937_64& stub_rng() { static std::mt19937_64 g(537); return g; }
inline void stub_set_seed(uint64_t s) { stub_rng().seed(s); }
inline double unif_rand() { return std::uniform_real_distribution<double>(0.0, 1.0)(stub_rng()); }
inline double norm_rand() { return std::normal_distribution<double>(0.0, 1.0)(stub_rng()); }
}  // namespace R
