// TEST INFRASTRUCTURE ONLY -- not part of the shipped engine.
//
// Stand-in for <Rcpp.h> so the reference's hot-path headers under
// /root/reference/src/{util,core,solver} compile UNMODIFIED into the
// `oracle/_ref` checker library (see oracle/Makefile, oracle/ref_driver.cpp).
// Only the surface those headers touch is provided; the List machinery only
// has to compile (Model::save_model/load_model, Tracker::save/load,
// SMatrix(List)) -- the driver never executes it.
//
// RNG seams (R's libR and glibc are outside the reference tree):
//   Rf_rnorm / Rf_rgamma  -> injectable streams of standard normals /
//                            unit-scale gammas (fmwr_ref_set_streams)
//   rand()                -> injectable stream of ints (macro redirect below)
#ifndef FMWR_ORACLE_RCPP_SHIM_H_
#define FMWR_ORACLE_RCPP_SHIM_H_

#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <algorithm>
#include <iostream>
#include <stdexcept>
#include <limits>
#include <initializer_list>

typedef unsigned int uint;

#define R_PosInf (std::numeric_limits<double>::infinity())
#define R_NegInf (-std::numeric_limits<double>::infinity())
#define NA_REAL (std::numeric_limits<double>::quiet_NaN())

inline int R_IsNaN(double x) { return std::isnan(x) ? 1 : 0; }

// ---- injectable RNG streams -------------------------------------------------
struct OracleStreams {
  const double* normals; long n_normals; long i_normal;   // standard normals
  const double* gammas;  long n_gammas;  long i_gamma;    // unit-scale gamma(shape) draws
  const int*    rands;   long n_rands;   long i_rand;     // glibc rand() values
  long overrun;                                           // draws past the end of a stream
};
extern OracleStreams g_oracle_streams;

inline double Rf_rnorm(double mean, double sd)
{
  OracleStreams& s = g_oracle_streams;
  double z = 0.0;
  if (s.normals && s.i_normal < s.n_normals) z = s.normals[s.i_normal];
  else s.overrun++;
  s.i_normal++;
  return mean + sd * z;
}

inline double Rf_rgamma(double shape, double scale)
{
  (void)shape;
  OracleStreams& s = g_oracle_streams;
  double g = 1.0;
  if (s.gammas && s.i_gamma < s.n_gammas) g = s.gammas[s.i_gamma];
  else s.overrun++;
  s.i_gamma++;
  return scale * g;
}

inline int oracle_rand()
{
  OracleStreams& s = g_oracle_streams;
  int r = 0;
  if (s.rands && s.i_rand < s.n_rands) r = s.rands[s.i_rand];
  else if (!s.rands) r = std::rand();
  else s.overrun++;
  s.i_rand++;
  return r;
}
#define rand oracle_rand

inline void Rf_warning(const char* msg) { std::fprintf(stderr, "[ref warning] %s\n", msg); }
typedef void* SEXP;
inline bool Rf_isNull(SEXP p) { return p == NULL; }

namespace Rcpp {

inline void stop(const char* msg) { throw std::runtime_error(msg); }
inline void stop(const std::string& msg) { throw std::runtime_error(msg); }

class NumericVector {
  std::vector<double> d_;
public:
  typedef double* iterator;
  NumericVector() {}
  NumericVector(int n) : d_(n, 0.0) {}
  NumericVector(uint n) : d_(n, 0.0) {}
  unsigned int size() const { return (unsigned int)d_.size(); }
  double* begin() { return d_.data(); }
  double* end() { return d_.data() + d_.size(); }
  double& operator[](long i) { return d_[i]; }
  double operator[](long i) const { return d_[i]; }
  static NumericVector create(double a, double b) { NumericVector v(2); v[0] = a; v[1] = b; return v; }
};

class IntegerVector {
  std::vector<int> d_;
public:
  typedef int* iterator;
  IntegerVector() {}
  IntegerVector(int n) : d_(n, 0) {}
  unsigned int size() const { return (unsigned int)d_.size(); }   // unsigned: compared with uint
  int* begin() { return d_.data(); }
  int* end() { return d_.data() + d_.size(); }
  int& operator[](long i) { return d_[i]; }
  int operator[](long i) const { return d_[i]; }
};

class NumericMatrix {
  int nr_, nc_;
  std::vector<double> d_;
public:
  NumericMatrix() : nr_(0), nc_(0) {}
  NumericMatrix(int r, int c) : nr_(r), nc_(c), d_((size_t)r * c, 0.0) {}
  int nrow() const { return nr_; }
  int ncol() const { return nc_; }
  double& operator()(int i, int j) { return d_[(size_t)j * nr_ + i]; }   // column-major like R
  double operator()(int i, int j) const { return d_[(size_t)j * nr_ + i]; }
  double* begin() { return d_.data(); }
};

class CharacterVector { public: CharacterVector() {} };
class String { std::string s_; public: String() {} String(const char* s) : s_(s) {} operator std::string() const { return s_; } };

class List;

// One dynamically-typed slot of a List (compile-only machinery).
class Slot {
public:
  double num; NumericVector nv; IntegerVector iv; NumericMatrix nm; std::string str; List* lst;
  Slot() : num(0.0), lst(NULL) {}
  Slot& operator=(double v) { num = v; return *this; }
  Slot& operator=(int v) { num = v; return *this; }
  Slot& operator=(bool v) { num = v; return *this; }
  Slot& operator=(const NumericVector& v) { nv = v; return *this; }
  Slot& operator=(const IntegerVector& v) { iv = v; return *this; }
  Slot& operator=(const NumericMatrix& v) { nm = v; return *this; }
  Slot& operator=(const CharacterVector&) { return *this; }
  Slot& operator=(const List& v);
  operator double() const { return num; }
  operator int() const { return (int)num; }
  operator uint() const { return (uint)num; }
  operator bool() const { return num != 0.0; }
  operator NumericVector() const { return nv; }
  operator IntegerVector() const { return iv; }
  operator NumericMatrix() const { return nm; }
  operator std::string() const { return str; }
  operator List() const;
};

struct NamedSlot { std::string name; Slot slot; };

class NamedProxy {
public:
  std::string name;
  NamedProxy(const char* n) : name(n) {}
  template <typename T> NamedSlot operator=(const T& v) const { NamedSlot s; s.name = name; s.slot = v; return s; }
};
struct NamedMaker { NamedProxy operator[](const char* n) const { return NamedProxy(n); } };
static const NamedMaker _ = NamedMaker();

class List {
  std::vector<NamedSlot> slots_;
  std::map<std::string, Slot> attrs_;
public:
  List() {}
  List(int n) : slots_(n) {}
  int size() const { return (int)slots_.size(); }
  Slot& operator[](const char* name)
  {
    for (size_t i = 0; i < slots_.size(); ++i) if (slots_[i].name == name) return slots_[i].slot;
    NamedSlot s; s.name = name; slots_.push_back(s); return slots_.back().slot;
  }
  Slot& operator[](const std::string& name) { return (*this)[name.c_str()]; }
  Slot& operator[](int i) { return slots_[i].slot; }
  Slot& attr(const char* name) { return attrs_[name]; }
  static List create() { return List(); }
  template <typename... A> static List create(const A&... a)
  {
    List l; NamedSlot arr[] = { a... };
    for (size_t i = 0; i < sizeof...(A); ++i) l.slots_.push_back(arr[i]);
    return l;
  }
};

inline Slot& Slot::operator=(const List& v) { if (lst) delete lst; lst = new List(v); return *this; }
inline Slot::operator List() const { return lst ? *lst : List(); }

template <typename T> T as(const Slot& s) { return s.operator T(); }
template <typename T> T as(const NumericVector& v);
template <> inline NumericVector as<NumericVector>(const NumericVector& v) { return v; }

}  // namespace Rcpp

#endif
