// TEST INFRASTRUCTURE -- CPU oracle for sosie_b200 (parity unpinned by the reference: see DESIGN.md).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (sosie_b200/) never links, imports or calls it.
//
// Common helpers: column-major 1-based views (Fortran layout), the libm switch, Fortran
// intrinsics with their exact semantics.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../sosie_b200/csrc/pmath.h"

namespace orc {

// 0 = glibc (the platform libm gfortran would bind to), 1 = pmath (portable, == device bits)
extern int g_libm;
extern int g_error;  // set when the reference would have hit a STOP

inline double m_sin(double x) { return g_libm ? pm_sin(x) : std::sin(x); }
inline double m_cos(double x) { return g_libm ? pm_cos(x) : std::cos(x); }
inline double m_tan(double x) { return g_libm ? pm_tan(x) : std::tan(x); }
inline double m_acos(double x) { return g_libm ? pm_acos(x) : std::acos(x); }
inline double m_log(double x) { return g_libm ? pm_log(x) : std::log(x); }
inline double m_atan2(double y, double x) { return g_libm ? pm_atan2(y, x) : std::atan2(y, x); }
inline float m_expf(float x) { return g_libm ? pm_expf(x) : ::expf(x); }

// Fortran (ni,nj) column-major, 1-based.  Strides (si,sj) default to (1,ni); a 1-D axis can be viewed as
// its 2-D expansion without materialising it (si=1,sj=0 for lon(i); si=0,sj=1 for lat(j)): same values
// as the reference's X1(:,jj)=X10(:,1) copies (src/mod_bilin_2d.f90:103-112), none of the memory.
template <class T>
struct A2 {
  T* p;
  int64_t ni, nj, si, sj;
  A2() : p(nullptr), ni(0), nj(0), si(1), sj(0) {}
  A2(T* p_, int64_t ni_, int64_t nj_) : p(p_), ni(ni_), nj(nj_), si(1), sj(ni_) {}
  A2(T* p_, int64_t ni_, int64_t nj_, int64_t si_, int64_t sj_) : p(p_), ni(ni_), nj(nj_), si(si_), sj(sj_) {}
  inline T& operator()(int64_t i, int64_t j) const { return p[(i - 1) * si + (j - 1) * sj]; }
  inline bool dense() const { return si == 1 && sj == ni; }
};
template <class T>
struct A3 {
  T* p;
  int64_t ni, nj, nk;
  A3(T* p_, int64_t ni_, int64_t nj_, int64_t nk_) : p(p_), ni(ni_), nj(nj_), nk(nk_) {}
  inline T& operator()(int64_t i, int64_t j, int64_t k) const { return p[(i - 1) + (j - 1) * ni + (k - 1) * ni * nj]; }
};

// Fortran intrinsics
inline double f_sign(double a, double b) { return std::copysign(std::fabs(a), b); }  // SIGN(a,b)
inline double f_mod(double a, double p) { return std::fmod(a, p); }                  // MOD for reals
inline int f_nint(double x) { return (int)std::lround(x); }                          // NINT: half away from zero

// REAL(4) literals widened to REAL(8) exactly as gfortran does
static const double kPi = 0x1.921fb54442d18p+1;         // ACOS(-1._8)
static const double kRepsilon = (double)1.E-9f;         // REAL(8), PARAMETER :: repsilon = 1.E-9
static const double kRmv = -9999.0;                     // REAL(rmissval,8), rmissval = -9999.
inline double earth_radius_km() {                       // zr = (6378.137 + 6356.7523)/2.0  (all REAL(4))
  float a = 6378.137f, b = 6356.7523f;
  float s = a + b;
  float r = s / 2.0f;
  return (double)r;
}

// src/mod_manip.f90:1727-1732  degE_to_degWE_scal
inline double degE_to_degWE(double rlong) {
  return f_sign(1.0, 180.0 - rlong) * std::min(rlong, std::fabs(rlong - 360.0));
}

}  // namespace orc
