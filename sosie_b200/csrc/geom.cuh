// geom.cuh -- device-side geometry of the bilinear mapping.  Every function is the GPU counterpart
// of a reference routine (cited) and is written so that its IEEE operations, their order and the
// REAL(4)/REAL(8) kinds are those of the Fortran source; libm calls go through pmath.h.
#pragma once
#include "common.cuh"

#define G_PI 0x1.921fb54442d18p+1
#define G_ZCONV (G_PI / 180.0)
// REAL(4) literals widened to REAL(8)
#define G_REPS ((double)1.E-9f)
#define G_RMV (-9999.0)
// zr = (6378.137 + 6356.7523)/2.0 evaluated in REAL(4), src/mod_manip.f90:1476,1544
#define G_ZR 6367.44482421875

__device__ __forceinline__ double d_sign(double a, double b) { return copysign(fabs(a), b); }

// src/mod_manip.f90:1727-1732
__device__ __forceinline__ double d_degE_to_degWE(double rlong) {
  return d_sign(1.0, 180.0 - rlong) * fmin(rlong, fabs(rlong - 360.0));
}

// ---- source accessors (1-based i,j) ----------------------------------------------------------------
__device__ __forceinline__ double src_lon(const SrcGeom& g, int i, int j) {
  return g.separable ? g.lon1[i - 1] : g.X[(size_t)(i - 1) + (size_t)(j - 1) * g.nx];
}
__device__ __forceinline__ double src_lat(const SrcGeom& g, int i, int j) {
  return g.separable ? g.lat1[j - 1] : g.Y[(size_t)(i - 1) + (size_t)(j - 1) * g.nx];
}
// unit vector of source point, as DISTANCE_2D forms it (src/mod_manip.f90:1558-1562)
__device__ __forceinline__ void src_vec(const SrcGeom& g, int i, int j, double& vx, double& vy, double& vz) {
  if (g.separable) {
    double cl = g.clat[j - 1];
    vx = g.clon[i - 1] * cl;
    vy = g.slon[i - 1] * cl;
    vz = g.slat[j - 1];
  } else {
    size_t k = (size_t)(i - 1) + (size_t)(j - 1) * g.nx;
    vx = g.vx[k];
    vy = g.vy[k];
    vz = g.vz[k];
  }
}
__device__ __forceinline__ double trg_lon(const TrgGeom& t, int i, int j) {
  return t.is_1d ? t.X[i - 1] : t.X[(size_t)(i - 1) + (size_t)(j - 1) * t.nx];
}
__device__ __forceinline__ double trg_lat(const TrgGeom& t, int i, int j) {
  return t.is_1d ? t.Y[j - 1] : t.Y[(size_t)(i - 1) + (size_t)(j - 1) * t.nx];
}

__device__ __forceinline__ void unit_vec(double lon, double lat, double& ux, double& uy, double& uz) {
  double zlatr = lat * G_ZCONV;
  double zlonr = lon * G_ZCONV;
  double sl, cl, sp, cp;
  pm_sincos(zlonr, &sl, &cl);
  pm_sincos(zlatr, &sp, &cp);
  ux = cl * cp;
  uy = sl * cp;
  uz = sp;
}

// the same, also returning cos(lat): |ux*c + uy*s| <= |cp| * (1 + a few ulp) for any (c, s) on the unit circle -- the row
// bound of the EASY window scan
__device__ __forceinline__ void unit_vec_cp(double lon, double lat, double& ux, double& uy, double& uz, double& cp_out) {
  double zlatr = lat * G_ZCONV;
  double zlonr = lon * G_ZCONV;
  double sl, cl, sp, cp;
  pm_sincos(zlonr, &sl, &cl);
  pm_sincos(zlatr, &sp, &cp);
  ux = cl * cp;
  uy = sl * cp;
  uz = sp;
  cp_out = cp;
}

// DISTANCE, src/mod_manip.f90:1443-1504
__device__ __forceinline__ double d_distance(double plona, double plonb, double plata, double platb) {
  double ux, uy, uz, vx, vy, vz;
  unit_vec(plona, plata, ux, uy, uz);
  unit_vec(plonb, platb, vx, vy, vz);
  double zpds = ux * vx + uy * vy + uz * vz;
  if (zpds >= 1.0) return 0.0;
  return G_ZR * pm_acos(zpds);
}

// the Mercator ordinate used by heading(): -LOG(MAX(ABS(TAN(zpi/4.-zconv*plat/2.)), repsilon))
__device__ __forceinline__ double d_merc(double plat) {
  double rr = fmax(fabs(pm_tan(G_PI / 4.0 - G_ZCONV * plat / 2.0)), G_REPS);
  return -pm_log(rr);
}

// the same ordinate for a source point, from the table built once per grid (identical bits: pure function)
__device__ __forceinline__ double src_merc(const SrcGeom& g, int i, int j) {
  return g.separable ? g.merc[j - 1] : g.merc[(size_t)(i - 1) + (size_t)(j - 1) * g.nx];
}
// MOD(a, p) for reals; exact shortcut when |a| < p (the common case: longitudes already in [0,360))
static __device__ __noinline__ double d_fmod_slow(double a, double p) { return fmod(a, p); }  // out of line: code size
__device__ __forceinline__ double d_mod(double a, double p) {
  if (fabs(a) < p) return a;
  return d_fmod_slow(a, p);
}

// heading, src/mod_bilin_2d.f90:818-870 (ya passed in: it is shared by the five calls at :511-515)
__device__ __forceinline__ double d_heading_y(double plona, double plonb, double ya, double yb) {
  double xa = plona * G_ZCONV;
  double xb = plonb * G_ZCONV;
  double xb_xa = d_mod((xb - xa), 2 * G_PI);
  if (xb_xa >= G_PI) xb_xa = xb_xa - 2 * G_PI;
  if (xb_xa <= -G_PI) xb_xa = xb_xa + 2 * G_PI;
  double angled = pm_atan2(xb_xa, yb - ya);
  double h = angled * 180.0 / G_PI;
  if (h < 0) h = h + 360.0;
  return h;
}

// PointSlope, src/mod_poly.f90:215-266
__device__ __forceinline__ void d_point_slope(double& pslup, float xa, float xb, float ya, float yb, double& pax,
                                              double& pcnstnt) {
  double zvertxa = xa, zvertxb = xb, zvertya = ya, zvertyb = yb;
  double zrise = zvertyb - zvertya;
  double zrun = zvertxb - zvertxa;
  // zrise == 0 (every horizontal edge of a regular grid's cells): the quotient is +-0 and the next statement makes it
  // +0 -- skip the division (CUDA's fp64 divide takes its ~70-instruction slow path on a zero numerator)
  if (zrun == 0) pslup = 9999;
  else if (zrise == 0) pslup = 0.0;
  else pslup = zrise / zrun;
  if (fabs(pslup) <= (double)0.001f) pslup = 0.0;
  if (pslup == 0) { pax = pslup; pcnstnt = zvertya; }
  else if (pslup == 9999) { pax = 1; pcnstnt = zvertxa; }
  else { pax = pslup; pcnstnt = (pslup * zvertxa - zvertya); }
}

// L_InPoly, src/mod_poly.f90:54-212.  Returns 0/1, or -1 where the reference STOPs (:79).
__device__ __forceinline__ int d_l_inpoly(const double* vplon, const double* vplat, double pxpoint, double pypoint) {
  float vertx[5], verty[5];
  float rmaxx = (float)fmax(fmax(vplon[0], vplon[1]), fmax(vplon[2], vplon[3]));
  float rmaxy = (float)fmax(fmax(vplat[0], vplat[1]), fmax(vplat[2], vplat[3]));
  float rminx = (float)fmin(fmin(vplon[0], vplon[1]), fmin(vplon[2], vplon[3]));
  float rminy = (float)fmin(fmin(vplat[0], vplat[1]), fmin(vplat[2], vplat[3]));
  if ((rminx < 0.f) || (pxpoint < 0.0)) return -1;
  double zx, zy;
  if ((rminx < 10.f) && (rmaxx > 350.f)) {
    zx = d_sign(1.0, 180.0 - pxpoint) * fmin(pxpoint, fabs(pxpoint - 360.0));
#pragma unroll
    for (int k = 0; k < 4; k++)
      vertx[k] = (float)(d_sign(1.0, 180.0 - vplon[k]) * fmin(vplon[k], fabs(vplon[k] - 360.0)));
    rmaxx = fmaxf(fmaxf(vertx[0], vertx[1]), fmaxf(vertx[2], vertx[3]));
    rminx = fminf(fminf(vertx[0], vertx[1]), fminf(vertx[2], vertx[3]));
  } else {
    zx = pxpoint;
#pragma unroll
    for (int k = 0; k < 4; k++) vertx[k] = (float)vplon[k];
  }
  zy = pypoint;
#pragma unroll
  for (int k = 0; k < 4; k++) verty[k] = (float)vplat[k];
  vertx[4] = vertx[0];
  verty[4] = verty[0];
#pragma unroll
  for (int jj = 0; jj < 5; jj++) {
    if ((vertx[jj] - (float)(int)vertx[jj]) == 0) vertx[jj] = __fadd_rn(vertx[jj], 0.001f);
    if ((verty[jj] - (float)(int)verty[jj]) == 0) verty[jj] = __fadd_rn(verty[jj], 0.001f);
  }
  if (!(zx <= (double)rmaxx)) return 0;
  if (!(zx >= (double)rminx)) return 0;
  if (!(zy <= (double)rmaxy)) return 0;
  if (!(zy >= (double)rminy)) return 0;
  int icross = 0;
#pragma unroll
  for (int ji = 0; ji < 4; ji++) {
    int jn = (ji == 3) ? 0 : ji + 1;
    double slope, ra, rc;
    d_point_slope(slope, vertx[ji], vertx[jn], verty[ji], verty[jn], ra, rc);
    double vxi = vertx[ji], vxn = vertx[jn], vyi = verty[ji], vyn = verty[jn];
    if (slope == 9999) {
      if (zx >= vxi) {
        if (((zy <= vyi) && (zy > vyn)) || ((zy >= vyi) && (zy < vyn))) icross++;
      }
    } else if (slope != 0) {
      double zxpt = (rc + zy) / ra;
      if (((zxpt <= vxi) && (zxpt > vxn)) || ((zxpt >= vxi) && (zxpt < vxn))) {
        if (zx >= zxpt) icross++;
      }
    }
  }
  return (icross & 1);
}

// L_InPoly for the cells of a regular source: the quad is (xa,ya) -> (xb,ya) -> (xb,yb) -> (xa,yb) (pattern A: quadrants 1
// and 3 of MAPPING_BL) or (xa,ya) -> (xa,yb) -> (xb,yb) -> (xb,ya) (pattern B: quadrants 2 and 4).  Same statements as
// d_l_inpoly on those vertices, with what is known about them folded in: equal REAL(8) values stay equal through the
// re-framing, the REAL(4) cast and the +0.001 nudge, so two edges have zrise == 0 -- PointSlope gives slope 0 (ignored),
// or 9999 when the cast also merged the two abscissae, and then the crossing test needs zy strictly between two equal
// ordinates: never counted either way -- and the other two have zrun == 0: slope 9999, the "vertical" branch.
__device__ __forceinline__ int d_l_inpoly_rect(double xa, double xb, double ya, double yb, bool pattern_a, double pxpoint,
                                               double pypoint) {
  float fxa = (float)xa, fxb = (float)xb;
  const float fya0 = (float)ya, fyb0 = (float)yb;
  float rmaxx = fmaxf(fxa, fxb), rminx = fminf(fxa, fxb);
  const float rmaxy = fmaxf(fya0, fyb0), rminy = fminf(fya0, fyb0);
  if ((rminx < 0.f) || (pxpoint < 0.0)) return -1;
  double zx;
  if ((rminx < 10.f) && (rmaxx > 350.f)) {
    zx = d_sign(1.0, 180.0 - pxpoint) * fmin(pxpoint, fabs(pxpoint - 360.0));
    fxa = (float)(d_sign(1.0, 180.0 - xa) * fmin(xa, fabs(xa - 360.0)));
    fxb = (float)(d_sign(1.0, 180.0 - xb) * fmin(xb, fabs(xb - 360.0)));
    rmaxx = fmaxf(fxa, fxb);
    rminx = fminf(fxa, fxb);
  } else {
    zx = pxpoint;
  }
  const double zy = pypoint;
  float fya = fya0, fyb = fyb0;
  if ((fxa - (float)(int)fxa) == 0) fxa = __fadd_rn(fxa, 0.001f);
  if ((fxb - (float)(int)fxb) == 0) fxb = __fadd_rn(fxb, 0.001f);
  if ((fya - (float)(int)fya) == 0) fya = __fadd_rn(fya, 0.001f);
  if ((fyb - (float)(int)fyb) == 0) fyb = __fadd_rn(fyb, 0.001f);
  if (!(zx <= (double)rmaxx)) return 0;
  if (!(zx >= (double)rminx)) return 0;
  if (!(zy <= (double)rmaxy)) return 0;
  if (!(zy >= (double)rminy)) return 0;
  const double Xa = fxa, Xb = fxb, Ya = fya, Yb = fyb;
  // an edge running from Ya to Yb counts zy in the interval that includes the Ya end; the one running back includes Yb
  const bool in_from_a = ((zy <= Ya) && (zy > Yb)) || ((zy >= Ya) && (zy < Yb));
  const bool in_from_b = ((zy <= Yb) && (zy > Ya)) || ((zy >= Yb) && (zy < Ya));
  // pattern A: edge (xb: ya -> yb) and edge (xa: yb -> ya); pattern B: edge (xa: ya -> yb) and edge (xb: yb -> ya)
  const double x_from_a = pattern_a ? Xb : Xa, x_from_b = pattern_a ? Xa : Xb;
  const int icross = (int)((zx >= x_from_a) && in_from_a) + (int)((zx >= x_from_b) && in_from_b);
  return (icross & 1);
}

// x / y for a finite non-zero y: identical to the IEEE quotient, without the divide's slow path when x == 0
__device__ __forceinline__ double d_div_nz(double x, double y) {
  if (x == 0.0) return (signbit(x) != signbit(y)) ? -0.0 : 0.0;
  return x / y;
}
__device__ __forceinline__ double d_det(double p1, double p2, double p3, double p4) { return p1 * p4 - p2 * p3; }

// LOCAL_COORD, src/mod_bilin_2d.f90:697-799
__device__ __forceinline__ void d_local_coord(const double* xlam, const double* xphi, double& xa, double& xb, int& ipb) {
  const int itermax = 200;
  double zxlam[5];
  ipb = 0;
#pragma unroll
  for (int k = 0; k < 5; k++) zxlam[k] = xlam[k];
  if ((fabs(zxlam[1] - zxlam[4]) >= 180.0) || (fabs(zxlam[1] - zxlam[2]) >= 180.0) ||
      (fabs(zxlam[1] - zxlam[3]) >= 180.0)) {
#pragma unroll
    for (int k = 0; k < 5; k++) zxlam[k] = d_degE_to_degWE(zxlam[k]);
  }
  // "zres > zresmax" with zres = SQRT(s) is decided on s itself: SQRT is correctly rounded and monotone, and
  // 0x1.47ae151eb8520p-7 is the largest REAL(8) whose square root rounds to <= REAL(0.1,4) (tests/test_pmath.py), so
  // s > that bound <=> SQRT(s) > zresmax (both false for NaN, both true for +Inf).  Saves one SQRT per iteration.
  const double zres2max = 0x1.47ae151eb8520p-7;
  double zres2 = 1.0e6, zdlam = 0.5, zdphi = 0.5, zalpha = 0.0, zbeta = 0.0;
  int niter = 0;
  const double z1 = (zxlam[2] - zxlam[1]);
  const double z2 = (zxlam[1] - zxlam[4]);
  const double z3 = (zxlam[3] - zxlam[2]);
  const double p21 = xphi[2] - xphi[1];
  const double p41 = xphi[4] - xphi[1];
  const double pc = (xphi[1] - xphi[4] + xphi[3] - xphi[2]);
  while ((zres2 > zres2max) && (niter < itermax)) {
    double za11 = z1 + (z2 + z3) * zbeta;
    double za12 = -z2 + (z2 + z3) * zalpha;
    double za21 = p21 + pc * zbeta;
    double za22 = p41 + pc * zalpha;
    double zdeta = d_det(za11, za12, za21, za22);
    zdeta = (d_sign(1.0, zdeta) * fmax(fabs(zdeta), G_REPS));
    double zdalp = d_div_nz(d_det(zdlam, za12, zdphi, za22), zdeta);  // |zdeta| >= repsilon: finite, non-zero
    double zdbet = d_div_nz(d_det(za11, zdlam, za21, zdphi), zdeta);
    zres2 = zdalp * zdalp + zdbet * zdbet;
    zalpha = zalpha + zdalp;
    zbeta = zbeta + zdbet;
    zdlam = zxlam[0] - ((1.0 - zalpha) * (1 - zbeta) * zxlam[1] + zalpha * (1 - zbeta) * zxlam[2] +
                        zalpha * zbeta * zxlam[3] + (1 - zalpha) * zbeta * zxlam[4]);
    zdphi = xphi[0] - ((1.0 - zalpha) * (1 - zbeta) * xphi[1] + zalpha * (1 - zbeta) * xphi[2] +
                       zalpha * zbeta * xphi[3] + (1 - zalpha) * zbeta * xphi[4]);
    niter = niter + 1;
  }
  if (niter >= itermax) { zalpha = 0.5; zbeta = 0.5; ipb = 11; }
  xa = zalpha;
  xb = zbeta;
  if ((xphi[1] == xphi[2]) && (xphi[2] == xphi[3]) && (xphi[3] == xphi[4])) { xa = 0.5; xb = 0.5; ipb = 12; }
}

// ---- lexicographic (value, index) min used by the parallel MINLOCs (first index wins ties, Q12) -----
struct MinIdx {
  double v;
  long long k;
};
__device__ __forceinline__ MinIdx minidx_better(MinIdx a, MinIdx b) {
  if (b.k < 0) return a;
  if (a.k < 0) return b;
  if (b.v < a.v || (b.v == a.v && b.k < a.k)) return b;
  return a;
}
__device__ __forceinline__ MinIdx warp_minidx(MinIdx m) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MinIdx t;
    t.v = __shfl_xor_sync(0xffffffffu, m.v, o);
    t.k = __shfl_xor_sync(0xffffffffu, m.k, o);
    m = minidx_better(m, t);
  }
  return m;
}
