// host build of geom.cuh (test infrastructure): the device functions compiled by g++ with the same IEEE flags
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
using std::signbit;
#include <algorithm>
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
#ifndef __noinline__
#define __noinline__
#endif
#include "geom.cuh"
#include "pack.cuh"
extern "C" {
int h_l_inpoly(const double* vplon, const double* vplat, double px, double py) { return d_l_inpoly(vplon, vplat, px, py); }
int h_l_inpoly_rect(double xa, double xb, double ya, double yb, int pattern_a, double px, double py) {
  return d_l_inpoly_rect(xa, xb, ya, yb, pattern_a != 0, px, py);
}
double h_heading(double lona, double lonb, double lata, double latb) { return d_heading_y(lona, lonb, d_merc(lata), d_merc(latb)); }
double h_degE_to_degWE(double x) { return d_degE_to_degWE(x); }
// pack_record (pack.cuh): the 32-byte apply record of one target from its mapping entry
void h_pack_record(int nx, int nyw, int separable, const double* lon1, const double* lat1, const double* X, const double* Y, int k_ew,
                   int iP, int jP, int iqd, double xa, double xb, double lon_t, double lat_t, int* id4, float* w4) {
  SrcGeom sg;
  sg.nx = nx; sg.ny = nyw; sg.nyw = nyw; sg.separable = separable; sg.lon1 = lon1; sg.lat1 = lat1; sg.X = X; sg.Y = Y;
  int4 id = make_int4(0, 0, 0, 0);
  float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
  pack_record(sg, k_ew, iP, jP, iqd, xa, xb, lon_t, lat_t, id, w);
  id4[0] = id.x; id4[1] = id.y; id4[2] = id.z; id4[3] = id.w;
  w4[0] = w.x; w4[1] = w.y; w4[2] = w.z; w4[3] = w.w;
}
void h_local_coord(const double* xlam, const double* xphi, double* xa, double* xb, int* ipb) { d_local_coord(xlam, xphi, *xa, *xb, *ipb); }
}
