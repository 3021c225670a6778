// host build of geom.cuh (test infrastructure): the device functions compiled by g++ with the same IEEE flags
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
using std::signbit;
#ifndef __noinline__
#define __noinline__
#endif
#include "geom.cuh"
extern "C" {
int h_l_inpoly(const double* vplon, const double* vplat, double px, double py) { return d_l_inpoly(vplon, vplat, px, py); }
int h_l_inpoly_rect(double xa, double xb, double ya, double yb, int pattern_a, double px, double py) {
  return d_l_inpoly_rect(xa, xb, ya, yb, pattern_a != 0, px, py);
}
double h_heading(double lona, double lonb, double lata, double latb) { return d_heading_y(lona, lonb, d_merc(lata), d_merc(latb)); }
double h_degE_to_degWE(double x) { return d_degE_to_degWE(x); }
void h_local_coord(const double* xlam, const double* xphi, double* xa, double* xb, int* ipb) { d_local_coord(xlam, xphi, *xa, *xb, *ipb); }
}
