// TEST INFRASTRUCTURE -- CPU oracle (see orc_common.hpp header).  "parity unpinned".
//
// Restatement of:
//   FILL_EXTRA_BANDS, extra_2_east/west      src/mod_manip.f90:69-305, 534-630
//   AKIMA_2D, build_pol, pol_val, SLOPES_AKIMA, find_nearest_akima   src/mod_akima_2d.f90:59-745
//   BDROWN, SMOOTHER                         src/mod_bdrown.f90:18-608
//   DROWN                                    src/mod_drown.f90:18-198
//   vector rotation                          src/corr_vect.f90:622-624, 963, 970
#include "orc_common.hpp"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// src/mod_manip.f90:534-569
static void extra_2_east(double x1, double x2, double x3, double x4, double x5, double y1, double y2, double y3,
                         double* y4, double* y5) {
  double A = x2 - x1, B = x3 - x2, C = x4 - x3, D = x5 - x4;
  double ALF = y2 - y1, BET = y3 - y2;
  if ((A == 0.) || (B == 0.) || (C == 0.)) { *y4 = y3; *y5 = y3; }
  else {
    *y4 = C * (2 * BET / B - ALF / A) + y3;
    *y5 = *y4 + *y4 * D / C + BET * D / B - ALF * D / A - y3 * D / C;
  }
}
// src/mod_manip.f90:592-630
static void extra_2_west(double x5, double x4, double x3, double x2, double x1, double y5, double y4, double y3,
                         double* y2, double* y1) {
  double A = x4 - x5, B = x3 - x4, C = x2 - x3, D = x1 - x2;
  double ALF = y4 - y5, BET = y3 - y4;
  if ((A == 0.) || (B == 0.) || (C == 0.)) { *y2 = y3; *y1 = y3; }
  else {
    *y2 = C * (2 * BET / B - ALF / A) + y3;
    *y1 = *y2 + *y2 * D / C + BET * D / B - ALF * D / A - y3 * D / C;
  }
}

// section copy dst(l0 + k*ls, jl) = src(r0 + k*rs, jr), k=0..n-1, RHS evaluated first.  Non-conforming
// ORCA-fold sections (iorca==4): the left-hand extent governs (see orc_bilin.cpp).
static void sect_copy(A2<double>& P, int l0, int ls, int n, int jl, int r0, int rs, int jr) {
  std::vector<double> tmp(n > 0 ? n : 0);
  for (int k = 0; k < n; k++) {
    int ir = r0 + k * rs;
    tmp[k] = (ir >= 1 && ir <= P.ni) ? P(ir, jr) : 0.0;
  }
  for (int k = 0; k < n; k++) P(l0 + k * ls, jl) = tmp[k];
}
static void orca_fold_top(A2<double>& P, int iorca) {
  int nxp4 = (int)P.ni, nyp4 = (int)P.nj;
  if (iorca == 4) {
    int nA = nxp4 / 2 - 1, nB = nxp4 / 2 + 3;
    sect_copy(P, 2, 1, nA, nyp4 - 1, nxp4, -1, nyp4 - 5);
    sect_copy(P, nxp4, -1, nB, nyp4 - 1, 2, 1, nyp4 - 5);
    sect_copy(P, 2, 1, nA, nyp4, nxp4, -1, nyp4 - 6);
    sect_copy(P, nxp4, -1, nB, nyp4, 2, 1, nyp4 - 6);
  } else {  // 6
    int nA = nxp4 / 2 - 1;
    sect_copy(P, 2, 1, nA, nyp4 - 1, nxp4 - 1, -1, nyp4 - 4);
    sect_copy(P, nxp4 - 1, -1, nA, nyp4 - 1, 2, 1, nyp4 - 4);
    sect_copy(P, 2, 1, nA, nyp4, nxp4 - 1, -1, nyp4 - 5);
    sect_copy(P, nxp4 - 1, -1, nA, nyp4, 2, 1, nyp4 - 5);
  }
}

// src/mod_manip.f90:69-305
void fill_extra_bands(int k_ew, const A2<const double>& XX, const A2<const double>& YY, const A2<const double>& XF,
                      A2<double>& XP4, A2<double>& YP4, A2<double>& FP4, int iorca) {
  int nx = (int)XX.ni, ny = (int)XX.nj;
  int nxp4 = nx + 4, nyp4 = ny + 4;
  for (int64_t k = 0; k < (int64_t)nxp4 * nyp4; k++) { XP4.p[k] = 0.; YP4.p[k] = 0.; FP4.p[k] = 0.; }
  for (int jj = 1; jj <= ny; jj++)
    for (int ji = 1; ji <= nx; ji++) {
      XP4(ji + 2, jj + 2) = XX(ji, jj);
      YP4(ji + 2, jj + 2) = YY(ji, jj);
      FP4(ji + 2, jj + 2) = XF(ji, jj);
    }
  // X array
  if (k_ew > -1) {
    for (int jj = 1; jj <= ny; jj++) {
      XP4(1, jj + 2) = XX(nx - 1 - k_ew, jj) - 360.;
      XP4(2, jj + 2) = XX(nx - k_ew, jj) - 360.;
      XP4(nxp4, jj + 2) = XX(2 + k_ew, jj) + 360.;
      XP4(nxp4 - 1, jj + 2) = XX(1 + k_ew, jj) + 360.;
    }
  } else {
    for (int jj = 1; jj <= ny; jj++) {
      XP4(2, jj + 2) = XX(2, jj) - (XX(3, jj) - XX(1, jj));
      XP4(1, jj + 2) = XX(1, jj) - (XX(3, jj) - XX(1, jj));
      XP4(nxp4 - 1, jj + 2) = XX(nx - 1, jj) + XX(nx, jj) - XX(nx - 2, jj);
      XP4(nxp4, jj + 2) = XX(nx, jj) + XX(nx, jj) - XX(nx - 2, jj);
    }
  }
  for (int ji = 1; ji <= nxp4; ji++) {
    double a2 = XP4(ji, 4) - (XP4(ji, 5) - XP4(ji, 3));
    double a1 = XP4(ji, 3) - (XP4(ji, 5) - XP4(ji, 3));
    XP4(ji, 2) = a2;
    XP4(ji, 1) = a1;
  }
  if (iorca == 4 || iorca == 6) orca_fold_top(XP4, iorca);
  else
    for (int ji = 1; ji <= nxp4; ji++) {
      double b1 = XP4(ji, nyp4 - 3) + XP4(ji, nyp4 - 2) - XP4(ji, nyp4 - 4);
      double b2 = XP4(ji, nyp4 - 2) + XP4(ji, nyp4 - 2) - XP4(ji, nyp4 - 4);
      XP4(ji, nyp4 - 1) = b1;
      XP4(ji, nyp4) = b2;
    }
  // Y array
  if (iorca == 4 || iorca == 6) orca_fold_top(YP4, iorca);
  else
    for (int ji = 1; ji <= nx; ji++) {
      YP4(ji + 2, nyp4 - 1) = YY(ji, ny - 1) + YY(ji, ny) - YY(ji, ny - 2);
      YP4(ji + 2, nyp4) = YY(ji, ny) + YY(ji, ny) - YY(ji, ny - 2);
    }
  for (int ji = 1; ji <= nx; ji++) {
    YP4(ji + 2, 2) = YY(ji, 2) - (YY(ji, 3) - YY(ji, 1));
    YP4(ji + 2, 1) = YY(ji, 1) - (YY(ji, 3) - YY(ji, 1));
  }
  if (k_ew > -1) {
    for (int jj = 1; jj <= nyp4; jj++) {
      double v1 = YP4(nx - 1 - k_ew + 2, jj), v2 = YP4(nx - k_ew + 2, jj);
      double v3 = YP4(2 + k_ew + 2, jj), v4 = YP4(1 + k_ew + 2, jj);
      YP4(1, jj) = v1;
      YP4(2, jj) = v2;
      YP4(nxp4, jj) = v3;
      YP4(nxp4 - 1, jj) = v4;
    }
  } else {
    for (int jj = 1; jj <= nyp4; jj++) {
      YP4(2, jj) = YP4(4, jj) - (YP4(5, jj) - YP4(3, jj));
      YP4(1, jj) = YP4(3, jj) - (YP4(5, jj) - YP4(3, jj));
      YP4(nxp4 - 1, jj) = YP4(nxp4 - 3, jj) + YP4(nxp4 - 2, jj) - YP4(nxp4 - 4, jj);
      YP4(nxp4, jj) = YP4(nxp4 - 2, jj) + YP4(nxp4 - 2, jj) - YP4(nxp4 - 4, jj);
    }
  }
  // Data array
  if (k_ew > -1) {
    for (int jj = 1; jj <= ny; jj++) {
      FP4(1, jj + 2) = XF(nx - 1 - k_ew, jj);
      FP4(2, jj + 2) = XF(nx - k_ew, jj);
      FP4(nxp4, jj + 2) = XF(2 + k_ew, jj);
      FP4(nxp4 - 1, jj + 2) = XF(1 + k_ew, jj);
    }
  } else {
    for (int jj = 3; jj <= nyp4 - 2; jj++)
      extra_2_east(XP4(nxp4 - 4, jj), XP4(nxp4 - 3, jj), XP4(nxp4 - 2, jj), XP4(nxp4 - 1, jj), XP4(nxp4, jj),
                   FP4(nxp4 - 4, jj), FP4(nxp4 - 3, jj), FP4(nxp4 - 2, jj), &FP4(nxp4 - 1, jj), &FP4(nxp4, jj));
    for (int jj = 3; jj <= nyp4 - 2; jj++)
      extra_2_west(XP4(5, jj), XP4(4, jj), XP4(3, jj), XP4(2, jj), XP4(1, jj), FP4(5, jj), FP4(4, jj), FP4(3, jj),
                   &FP4(2, jj), &FP4(1, jj));
  }
  if (iorca == 4 || iorca == 6) orca_fold_top(FP4, iorca);
  else
    for (int ji = 1; ji <= nxp4; ji++)
      extra_2_east(YP4(ji, nyp4 - 4), YP4(ji, nyp4 - 3), YP4(ji, nyp4 - 2), YP4(ji, nyp4 - 1), YP4(ji, nyp4),
                   FP4(ji, nyp4 - 4), FP4(ji, nyp4 - 3), FP4(ji, nyp4 - 2), &FP4(ji, nyp4 - 1), &FP4(ji, nyp4));
  for (int ji = 1; ji <= nxp4; ji++)
    extra_2_west(YP4(ji, 5), YP4(ji, 4), YP4(ji, 3), YP4(ji, 2), YP4(ji, 1), FP4(ji, 5), FP4(ji, 4), FP4(ji, 3),
                 &FP4(ji, 2), &FP4(ji, 1));
}

static inline double guard(double rx) { return f_sign(1.0, rx) * std::max(std::fabs(rx), kRepsilon); }

// src/mod_akima_2d.f90:483-664
void slopes_akima(int k_ew, const A2<const double>& XX, const A2<const double>& XY, const A2<const double>& XF,
                  A2<double>& dFdX, A2<double>& dFdY, A2<double>& d2FdXdY) {
  int nx = (int)XF.ni, ny = (int)XF.nj;
  int64_t ne = (int64_t)(nx + 4) * (ny + 4);
  std::vector<double> zx(ne), zy(ne), zf(ne);
  A2<double> ZX(zx.data(), nx + 4, ny + 4), ZY(zy.data(), nx + 4, ny + 4), ZF(zf.data(), nx + 4, ny + 4);
  fill_extra_bands(k_ew, XX, XY, XF, ZX, ZY, ZF, 0);
  for (int jj = 1; jj <= ny; jj++)
    for (int ji = 1; ji <= nx; ji++) {
      double m1, m2, m3, m4, Wx2 = 0., Wx3 = 0., Wy2 = 0., Wy3 = 0.;
      int ip1 = ji + 1, ip2 = ji + 2, jp1 = jj + 1, jp2 = jj + 2;
      m1 = (ZF(ip1, jp2) - ZF(ji, jp2)) / guard(ZX(ip1, jp2) - ZX(ji, jp2));
      m2 = (ZF(ip2, jp2) - ZF(ip1, jp2)) / guard(ZX(ip2, jp2) - ZX(ip1, jp2));
      m3 = (ZF(ji + 3, jp2) - ZF(ip2, jp2)) / guard(ZX(ji + 3, jp2) - ZX(ip2, jp2));
      m4 = (ZF(ji + 4, jp2) - ZF(ji + 3, jp2)) / guard(ZX(ji + 4, jp2) - ZX(ji + 3, jp2));
      if ((m1 == m2) && (m3 == m4)) dFdX(ji, jj) = 0.5 * (m2 + m3);
      else {
        Wx2 = std::fabs(m4 - m3);
        Wx3 = std::fabs(m2 - m1);
        dFdX(ji, jj) = (Wx2 * m2 + Wx3 * m3) / (Wx2 + Wx3);
      }
      double mx2 = m2, mx3 = m3;  // (m2,m3 are overwritten below; the cross term uses the Y ones, as the reference)
      (void)mx2; (void)mx3;
      m1 = (ZF(ip2, jp1) - ZF(ip2, jj)) / guard(ZY(ip2, jp1) - ZY(ip2, jj));
      m2 = (ZF(ip2, jp2) - ZF(ip2, jp1)) / guard(ZY(ip2, jp2) - ZY(ip2, jp1));
      m3 = (ZF(ip2, jj + 3) - ZF(ip2, jp2)) / guard(ZY(ip2, jj + 3) - ZY(ip2, jp2));
      m4 = (ZF(ip2, jj + 4) - ZF(ip2, jj + 3)) / guard(ZY(ip2, jj + 4) - ZY(ip2, jj + 3));
      if ((m1 == m2) && (m3 == m4)) dFdY(ji, jj) = 0.5 * (m2 + m3);
      else {
        Wy2 = std::fabs(m4 - m3);
        Wy3 = std::fabs(m2 - m1);
        dFdY(ji, jj) = (Wy2 * m2 + Wy3 * m3) / (Wy2 + Wy3);
      }
      double d22 = (ZF(ip1, jp2) - ZF(ip1, jp1)) / guard(ZY(ip1, jp2) - ZY(ip1, jp1));
      double d23 = (ZF(ip1, jj + 3) - ZF(ip1, jp2)) / guard(ZY(ip1, jj + 3) - ZY(ip1, jp2));
      double d42 = (ZF(ji + 3, jp2) - ZF(ji + 3, jp1)) / guard(ZY(ji + 3, jp2) - ZY(ji + 3, jp1));
      double d43 = (ZF(ji + 3, jj + 3) - ZF(ji + 3, jp2)) / guard(ZY(ji + 3, jj + 3) - ZY(ji + 3, jp2));
      double e22 = (m2 - d22) / guard(ZX(ip2, jp1) - ZX(ip1, jp1));
      double e23 = (m3 - d23) / guard(ZX(ip2, jp2) - ZX(ip1, jp2));
      double e32 = (d42 - m2) / guard(ZX(ji + 3, jp2) - ZX(ip2, jp2));
      double e33 = (d43 - m3) / guard(ZX(ji + 3, jp2) - ZX(ip2, jp2));
      if (((Wx2 == 0) && (Wx3 == 0)) || ((Wy2 == 0) && (Wy3 == 0))) {
        if ((Wx2 == 0) && (Wx3 == 0)) { Wx2 = 1.; Wx3 = 1.; }
        if ((Wy2 == 0) && (Wy3 == 0)) { Wy2 = 1.; Wy3 = 1.; }
      }
      d2FdXdY(ji, jj) = (Wx2 * (Wy2 * e22 + Wy3 * e23) + Wx3 * (Wy2 * e32 + Wy3 * e33)) / ((Wx2 + Wx3) * (Wy2 + Wy3));
    }
}

// src/mod_akima_2d.f90:249-441, one cell
static void build_pol_cell(const A2<double>& ZX, const A2<double>& ZY, const A2<double>& ZF, const A2<double>& sx,
                           const A2<double>& sy, const A2<double>& sxy, int ji, int jj, double VX[16]) {
  for (int k = 0; k < 16; k++) VX[k] = 0.;
  double x = ZX(ji + 1, jj) - ZX(ji, jj);
  double y = ZY(ji, jj + 1) - ZY(ji, jj);
  double x2 = x * x, x3 = x2 * x, y2 = y * y, y3 = y2 * y, xy = x * y;
  double b1 = ZF(ji, jj), b2 = ZF(ji + 1, jj), b3 = ZF(ji + 1, jj + 1), b4 = ZF(ji, jj + 1);
  double b5 = sx(ji, jj), b6 = sx(ji + 1, jj), b7 = sx(ji + 1, jj + 1), b8 = sx(ji, jj + 1);
  double b9 = sy(ji, jj), b10 = sy(ji + 1, jj), b11 = sy(ji + 1, jj + 1), b12 = sy(ji, jj + 1);
  double b13 = sxy(ji, jj), b14 = sxy(ji + 1, jj), b15 = sxy(ji + 1, jj + 1), b16 = sxy(ji, jj + 1);
  VX[10] = b13; VX[11] = b5; VX[14] = b9; VX[15] = b1;
  b5 = x * b5; b6 = x * b6; b7 = x * b7; b8 = x * b8;
  b9 = y * b9; b10 = y * b10; b11 = y * b11; b12 = y * b12;
  b13 = xy * b13; b14 = xy * b14; b15 = xy * b15; b16 = xy * b16;
  double c1 = b1 - b2, c2 = b3 - b4, c3 = b5 + b6, c4 = b7 + b8, c5 = b9 - b10, c6 = b11 - b12;
  double c7 = b13 + b14, c8 = b15 + b16, c9 = 2 * b5 + b6, c10 = b7 + 2 * b8, c11 = 2 * b13 + b14;
  double c12 = b15 + 2 * b16, c13 = b5 - b8, c14 = b1 - b4, c15 = b13 + b16, c16 = 2 * b13 + b16;
  double c17 = b9 + b12, c18 = 2 * b9 + b12;
  double d1 = c1 + c2, d2 = c3 - c4, d3 = c5 - c6, d4 = c7 + c8, d5 = c9 - c10, d6 = 2 * c5 - c6;
  double d7 = 2 * c7 + c8, d8 = c11 + c12, d9 = 2 * c11 + c12;
  double f1 = 2 * d1 + d2, f2 = 2 * d3 + d4, f3 = 2 * d6 + d7, f4 = 3 * d1 + d5, f5 = 3 * d3 + d8, f6 = 3 * d6 + d9;
  VX[0] = 2 * f1 + f2; VX[1] = -(3 * f1 + f3); VX[2] = 2 * c5 + c7;
  VX[3] = 2 * c1 + c3; VX[4] = -(2 * f4 + f5); VX[5] = 3 * f4 + f6;
  VX[6] = -(3 * c5 + c11); VX[7] = -(3 * c1 + c9); VX[8] = 2 * c13 + c15;
  VX[9] = -(3 * c13 + c16); VX[12] = 2 * c14 + c17; VX[13] = -(3 * c14 + c18);
  VX[0] = VX[0] / (x3 * y3); VX[1] = VX[1] / (x3 * y2); VX[2] = VX[2] / (x3 * y); VX[3] = VX[3] / x3;
  VX[4] = VX[4] / (x2 * y3); VX[5] = VX[5] / (x2 * y2); VX[6] = VX[6] / (x2 * y); VX[7] = VX[7] / x2;
  VX[8] = VX[8] / (x * y3); VX[9] = VX[9] / (x * y2); VX[12] = VX[12] / y3; VX[13] = VX[13] / y2;
}

// src/mod_akima_2d.f90:445-480
static inline double pol_val(double x, double y, const double V[16]) {
  double p1 = ((V[0] * y + V[1]) * y + V[2]) * y + V[3];
  double p2 = ((V[4] * y + V[5]) * y + V[6]) * y + V[7];
  double p3 = ((V[8] * y + V[9]) * y + V[10]) * y + V[11];
  double p4 = ((V[12] * y + V[13]) * y + V[14]) * y + V[15];
  return ((p1 * x + p2) * x + p3) * x + p4;
}

// src/mod_akima_2d.f90:670-745 (literal scan)
static void find_nearest_akima(const A2<double>& plon_src, const A2<double>& plat_src, const double pxyr[4],
                               const A2<const double>& plon_trg, const A2<const double>& plat_trg, A3<int32_t>& ixyp) {
  int nxs = (int)plon_src.ni, nys = (int)plon_src.nj, nxt = (int)ixyp.ni, nyt = (int)ixyp.nj;
  for (int jjt = 1; jjt <= nyt; jjt++)
    for (int jit = 1; jit <= nxt; jit++) {
      int jis = 1, jjs = 1;
      bool l_x_found = false, l_y_found = false;
      double pxt = plon_trg(jit, jjt), pyt = plat_trg(jit, jjt);
      if (((pxt >= pxyr[0]) && (pxt <= pxyr[1])) && ((pyt >= pxyr[2]) && (pyt <= pxyr[3]))) {
        while (!(l_x_found && l_y_found)) {
          l_x_found = false;
          while (!l_x_found) {
            if (jis < nxs) {
              if ((plon_src(jis, jjs) <= pxt) && (plon_src(jis + 1, jjs) > pxt)) l_x_found = true;
              else jis = jis + 1;
            } else { jis = jis - 1; l_x_found = true; }
          }
          l_y_found = false;
          while (!l_y_found) {
            if (jjs < nys) {
              if ((plat_src(jis, jjs) <= pyt) && (plat_src(jis, jjs + 1) > pyt)) l_y_found = true;
              else { jjs = jjs + 1; l_x_found = false; l_y_found = true; }
            } else { jjs = nys - 1; l_y_found = true; }
          }
        }
        ixyp(jit, jjt, 1) = jis;
        ixyp(jit, jjt, 2) = jjs;
      }
    }
}

// ------------------------------------------------------------------------------------------
// src/mod_bdrown.f90:444-608
static const float ris2 = (float)(1.f / std::sqrt(2.f));  // REAL( 1./SQRT(2.) , 4 )

void smoother(int k_ew, A2<float>& X, int nsmooth, const int32_t* msk_p, bool l_emp) {
  const float w0 = 0.35f;
  int ni = (int)X.ni, nj = (int)X.nj;
  int64_t n = (int64_t)ni * nj;
  bool l_mask = (msk_p != nullptr);
  if (l_emp && !l_mask) { g_error = 30; return; }
  A2<const int32_t> msk(msk_p, ni, nj);
  std::vector<float> xtmp_v(n), xorig_v, rdnm_v;
  std::vector<int32_t> nbp_v;
  A2<float> xtmp(xtmp_v.data(), ni, nj);
  if (l_emp) { rdnm_v.assign(n, 0.25f); nbp_v.assign(n, 0); }
  if (l_mask) xorig_v.assign(X.p, X.p + n);
  A2<float> rdnm(rdnm_v.data(), ni, nj), xorig(xorig_v.data(), ni, nj);
  A2<int32_t> nbp(nbp_v.data(), ni, nj);
  int ivi[2] = {1, ni}, vim_per[2] = {0, 0}, vip_per[2] = {0, 0};
  if (k_ew >= 0) { vim_per[0] = ni - k_ew; vim_per[1] = ni - 1; vip_per[0] = 2; vip_per[1] = 1 + k_ew; }
  const float c4 = 4.f * (1.f + ris2);
  const float omw = 1.f - w0;
  for (int js = 1; js <= nsmooth; js++) {
    std::memcpy(xtmp_v.data(), X.p, sizeof(float) * n);
    if (l_emp) for (int64_t k = 0; k < n; k++) xtmp_v[k] = xtmp_v[k] * (float)msk_p[k];
    auto stencil = [&](int ji, int jim, int jip, int jj) {
      float s4 = xtmp(jip, jj) + xtmp(ji, jj + 1) + xtmp(jim, jj) + xtmp(ji, jj - 1);
      float sd = xtmp(jip, jj + 1) + xtmp(jip, jj - 1) + xtmp(jim, jj - 1) + xtmp(jim, jj + 1);
      if (l_emp) {
        nbp(ji, jj) = msk(jip, jj) + msk(ji, jj + 1) + msk(jim, jj) + msk(ji, jj - 1) + msk(jip, jj + 1) +
                      msk(jip, jj - 1) + msk(jim, jj - 1) + msk(jim, jj + 1);
        int i4 = msk(jip, jj) + msk(ji, jj + 1) + msk(jim, jj) + msk(ji, jj - 1);
        int id = msk(jip, jj + 1) + msk(jip, jj - 1) + msk(jim, jj - 1) + msk(jim, jj + 1);
        rdnm(ji, jj) = 1.f / std::max((float)((float)i4 + ris2 * (float)id), 0.01f);
        X(ji, jj) = w0 * xtmp(ji, jj) + omw * (s4 + ris2 * sd) * rdnm(ji, jj);
        if (nbp(ji, jj) == 0) X(ji, jj) = xorig(ji, jj);
      } else {
        X(ji, jj) = w0 * xtmp(ji, jj) + omw * (s4 + ris2 * sd) / c4;
      }
    };
    for (int jj = 2; jj <= nj - 1; jj++)
      for (int ji = 2; ji <= ni - 1; ji++) stencil(ji, ji - 1, ji + 1, jj);
    if (k_ew >= 0)
      for (int jci = 0; jci < 2; jci++)
        for (int jj = 2; jj <= nj - 1; jj++) stencil(ivi[jci], vim_per[jci], vip_per[jci], jj);
    if (l_mask)
      for (int jj = 2; jj <= nj - 1; jj++)
        for (int ji = 1; ji <= ni; ji++)
          X(ji, jj) = (float)msk(ji, jj) * X(ji, jj) - (float)(msk(ji, jj) - 1) * xorig(ji, jj);
  }
}

// src/mod_bdrown.f90:18-439.  zmin/zmax/zval_land (implicit SAVE, Q16) are always passed here.
int bdrown(int k_ew, A2<float>& X, const int32_t* mask_p, int ninc_max, int nsmooth, float zmin, float zmax,
           float zval_land) {
  int ni = (int)X.ni, nj = (int)X.nj;
  int64_t n = (int64_t)ni * nj;
  A2<const int32_t> mask(mask_p, ni, nj);
  std::vector<int32_t> maskv_v(mask_p, mask_p + n), mc_v(n), mtmp_v(n);
  std::vector<float> dold_v(n);
  A2<int32_t> maskv(maskv_v.data(), ni, nj), mask_coast(mc_v.data(), ni, nj);
  A2<float> dold(dold_v.data(), ni, nj);
  int ivi[2] = {1, ni}, vim_per[2] = {0, 0}, vip_per[2] = {0, 0};
  if (k_ew >= 0) { vim_per[0] = ni - k_ew; vim_per[1] = ni - 1; vip_per[0] = 2; vip_per[1] = 1 + k_ew; }

  for (int jj = 1; jj <= nj; jj++) {  // :109-120
    int nbp = 0;
    for (int ji = 1; ji <= ni; ji++) nbp += mask(ji, jj);
    if (nbp > 0) {
      float s = 0.f;
      for (int ji = 1; ji <= ni; ji++) s += X(ji, jj) * (float)mask(ji, jj);
      float zmv = s / (float)nbp;
      for (int ji = 1; ji <= ni; ji++) if (mask(ji, jj) == 0) X(ji, jj) = zmv;
    }
  }
  auto T = [&](int i, int j) -> float { return (float)maskv(i, j) * dold(i, j); };
  auto TD = [&](int i, int j) -> float { return ris2 * (float)maskv(i, j) * dold(i, j); };
  auto MD = [&](int i, int j) -> float { return ris2 * (float)maskv(i, j); };
  int nsweeps = 0;
  for (int jinc = 1; jinc <= ninc_max; jinc++) {
    bool any0 = false;
    for (int64_t k = 0; k < n; k++) if (maskv_v[k] == 0) { any0 = true; break; }
    if (!any0) break;
    nsweeps++;
    std::memcpy(dold_v.data(), X.p, sizeof(float) * n);
    std::fill(mc_v.begin(), mc_v.end(), 0);
    for (int jj = 2; jj <= nj - 1; jj++)
      for (int ji = 2; ji <= ni - 1; ji++)
        mask_coast(ji, jj) = (maskv(ji + 1, jj) + maskv(ji, jj + 1) + maskv(ji - 1, jj) + maskv(ji, jj - 1)) * (-(maskv(ji, jj) - 1));
    if (k_ew >= 0) {
      for (int jci = 0; jci < 2; jci++) {
        int jim = vim_per[jci], ji = ivi[jci], jip = vip_per[jci];
        for (int jj = 2; jj <= nj - 1; jj++)
          mask_coast(ji, jj) = (maskv(jip, jj) + maskv(ji, jj + 1) + maskv(jim, jj) + maskv(ji, jj - 1)) * (-(maskv(ji, jj) - 1));
      }
    } else {
      for (int jj = 2; jj <= nj - 1; jj++) {
        mask_coast(1, jj) = (maskv(2, jj) + maskv(1, jj + 1) + maskv(1, jj - 1)) * (-(maskv(1, jj) - 1));
        mask_coast(ni, jj) = (maskv(ni, jj + 1) + maskv(ni - 1, jj) + maskv(ni, jj - 1)) * (-(maskv(ni, jj) - 1));
      }
    }
    for (int ji = 2; ji <= ni - 1; ji++)
      mask_coast(ji, 1) = (maskv(ji + 1, 1) + maskv(ji, 2) + maskv(ji - 1, 1)) * (-(maskv(ji, 1) - 1));
    if (k_ew >= 0) mask_coast(1, 1) = (maskv(2, 1) + maskv(1, 2) + maskv(ni - k_ew, 1)) * (-(maskv(1, 1) - 1));
    else mask_coast(1, 1) = (maskv(2, 1) + maskv(1, 2)) * (-(maskv(1, 1) - 1));
    if (k_ew >= 0) mask_coast(ni, 1) = (maskv(1 + k_ew, 1) + maskv(ni, 2) + maskv(ni, 1)) * (-(maskv(ni, 1) - 1));
    else mask_coast(ni, 1) = (maskv(ni, 2) + maskv(ni, 1)) * (-(maskv(ni, 1) - 1));
    for (int ji = 2; ji <= ni - 1; ji++)
      mask_coast(ji, nj) = (maskv(ji + 1, nj) + maskv(ji - 1, nj) + maskv(ji, nj - 1)) * (-(maskv(ji, nj) - 1));
    if (k_ew >= 0) mask_coast(1, nj) = (maskv(2, nj) + maskv(ni - k_ew, nj) + maskv(1, nj - 1)) * (-(maskv(1, nj) - 1));
    else mask_coast(1, nj) = (maskv(2, nj) + maskv(1, nj - 1)) * (-(maskv(1, nj) - 1));
    if (k_ew >= 0) mask_coast(ni, nj) = (maskv(k_ew + 1, nj) + maskv(ni - 1, nj) + maskv(ni - 1, nj - 1)) * (-(maskv(ni, nj) - 1));
    else mask_coast(ni, nj) = (maskv(ni - 1, nj) + maskv(ni - 1, nj - 1)) * (-(maskv(ni, nj) - 1));
    for (int64_t k = 0; k < n; k++) mc_v[k] = (mc_v[k] > 0) ? 1 : 0;

    // centre (:224-235)
    for (int jj = 2; jj <= nj - 1; jj++)
      for (int ji = 2; ji <= ni - 1; ji++)
        if (mask_coast(ji, jj) == 1)
          X(ji, jj) = 1.f / ((float)(maskv(ji + 1, jj) + maskv(ji, jj + 1) + maskv(ji - 1, jj) + maskv(ji, jj - 1)) +
                             ris2 * (float)(maskv(ji + 1, jj + 1) + maskv(ji - 1, jj + 1) + maskv(ji - 1, jj - 1) + maskv(ji + 1, jj - 1))) *
                     (T(ji + 1, jj) + T(ji, jj + 1) + T(ji - 1, jj) + T(ji, jj - 1) + TD(ji + 1, jj + 1) + TD(ji - 1, jj + 1) +
                      TD(ji - 1, jj - 1) + TD(ji + 1, jj - 1));
    for (int jci = 0; jci < 2; jci++) {
      int ji = ivi[jci];
      if (k_ew >= 0) {
        int jim = vim_per[jci], jip = vip_per[jci];
        for (int jj = 2; jj <= nj - 1; jj++)
          if (mask_coast(ji, jj) == 1)
            X(ji, jj) = 1.f / ((float)(maskv(jip, jj) + maskv(ji, jj + 1) + maskv(jim, jj) + maskv(ji, jj - 1)) +
                               ris2 * (float)(maskv(jip, jj + 1) + maskv(jim, jj + 1) + maskv(jim, jj - 1) + maskv(jip, jj - 1))) *
                       (T(jip, jj) + T(ji, jj + 1) + T(jim, jj) + T(ji, jj - 1) + TD(jip, jj + 1) + TD(jim, jj + 1) +
                        TD(jim, jj - 1) + TD(jip, jj - 1));
      } else {
        if (ji == 1)
          for (int jj = 2; jj <= nj - 1; jj++)
            if (mask_coast(ji, jj) == 1)
              X(ji, jj) = 1.f / ((float)(maskv(2, jj) + maskv(ji, jj + 1) + maskv(ji, jj - 1)) + MD(2, jj + 1) + MD(2, jj - 1)) *
                         (T(2, jj) + T(ji, jj + 1) + T(ji, jj - 1) + TD(2, jj + 1) + TD(2, jj - 1));
        if (ji == ni)
          for (int jj = 2; jj <= nj - 1; jj++)
            if (mask_coast(ji, jj) == 1)
              X(ji, jj) = 1.f / ((float)(maskv(ji, jj + 1) + maskv(ni - 1, jj) + maskv(ji, jj - 1)) + MD(ni - 1, jj + 1) + MD(ni - 1, jj - 1)) *
                         (T(ji, jj + 1) + T(ni - 1, jj) + T(ji, jj - 1) + TD(ni - 1, jj + 1) + TD(ni - 1, jj - 1));
      }
    }
    {  // top row (:298-347)
      int jj = nj;
      for (int ji = 2; ji <= ni - 1; ji++)
        if (mask_coast(ji, jj) == 1)
          X(ji, jj) = 1.f / ((float)(maskv(ji + 1, jj) + maskv(ji - 1, jj) + maskv(ji, jj - 1)) + MD(ji - 1, jj - 1) + MD(ji + 1, jj - 1)) *
                     (T(ji + 1, jj) + T(ji - 1, jj) + T(ji, jj - 1) + TD(ji - 1, jj - 1) + TD(ji + 1, jj - 1));
      for (int jci = 0; jci < 2; jci++) {
        int ji = ivi[jci];
        if (k_ew >= 0) {
          int jim = vim_per[jci], jip = vip_per[jci];
          if (mask_coast(ji, jj) == 1)
            X(ji, jj) = 1.f / ((float)(maskv(jip, jj) + maskv(jim, jj) + maskv(ji, jj - 1)) + MD(jim, jj - 1) + MD(jip, jj - 1)) *
                       (T(jip, jj) + T(jim, jj) + T(ji, jj - 1) + TD(jim, jj - 1) + TD(jip, jj - 1));
        } else {
          if (ji == 1 && mask_coast(ji, jj) == 1)
            X(ji, jj) = 1.f / ((float)(maskv(2, jj) + maskv(ji, jj - 1)) + MD(2, jj - 1)) * (T(2, jj) + T(ji, jj - 1) + TD(2, jj - 1));
          if (ji == ni && mask_coast(ji, jj) == 1)
            X(ji, jj) = 1.f / ((float)(maskv(ni - 1, jj) + maskv(ji, jj - 1)) + MD(ni - 1, jj - 1)) *
                       (T(ni - 1, jj) + T(ji, jj - 1) + TD(ni - 1, jj - 1));
        }
      }
    }
    {  // bottom row (:352-400)
      int jj = 1;
      for (int ji = 2; ji <= ni - 1; ji++)
        if (mask_coast(ji, jj) == 1)
          X(ji, jj) = 1.f / ((float)(maskv(ji + 1, jj) + maskv(ji, jj + 1) + maskv(ji - 1, jj)) + MD(ji + 1, jj + 1) + MD(ji - 1, jj + 1)) *
                     (T(ji + 1, jj) + T(ji, jj + 1) + T(ji - 1, jj) + TD(ji + 1, jj + 1) + TD(ji - 1, jj + 1));
      for (int jci = 0; jci < 2; jci++) {
        int ji = ivi[jci];
        if (k_ew >= 0) {
          int jim = vim_per[jci], jip = vip_per[jci];
          if (mask_coast(ji, jj) == 1)
            X(ji, jj) = 1.f / ((float)(maskv(jip, jj) + maskv(ji, jj + 1) + maskv(jim, jj)) + MD(jip, jj + 1) + MD(jim, jj + 1)) *
                       (T(jip, jj) + T(ji, jj + 1) + T(jim, jj) + TD(jip, jj + 1) + TD(jim, jj + 1));
        } else {
          if (ji == 1 && mask_coast(ji, jj) == 1)
            X(ji, jj) = 1.f / ((float)(maskv(2, jj) + maskv(ji, jj + 1)) + MD(2, jj + 1)) * (T(2, jj) + T(ji, jj + 1) + TD(2, jj + 1));
          if (ji == ni && mask_coast(ji, jj) == 1)
            X(ji, jj) = 1.f / ((float)(maskv(ji, jj + 1) + maskv(ni - 1, jj)) + MD(ni - 1, jj + 1)) *
                       (T(ji, jj + 1) + T(ni - 1, jj) + TD(ni - 1, jj + 1));
        }
      }
    }
    for (int64_t k = 0; k < n; k++) maskv_v[k] = maskv_v[k] + mc_v[k];
    for (int64_t k = 0; k < n; k++) X.p[k] = std::max(X.p[k], zmin);
    for (int64_t k = 0; k < n; k++) X.p[k] = std::min(X.p[k], zmax);
  }
  if (nsmooth > 0) {
    for (int64_t k = 0; k < n; k++) mtmp_v[k] = 1 - mask_p[k];
    smoother(k_ew, X, nsmooth, mtmp_v.data(), false);
  }
  if (std::fabs(zval_land) > 1.e-12f)
    for (int64_t k = 0; k < n; k++) if (maskv_v[k] == 0) X.p[k] = zval_land;
  return nsweeps;
}

// src/mod_drown.f90:18-198
int drown(int k_ew, A2<float>& X, const int8_t* mask_p, int ninc_max) {
  int ni = (int)X.ni, nj = (int)X.nj;
  int64_t n = (int64_t)ni * nj;
  for (int64_t k = 0; k < n; k++) X.p[k] = X.p[k] * (float)mask_p[k];
  std::vector<int8_t> maskv_v(mask_p, mask_p + n), mc_v(n);
  std::vector<float> xtmp_v(n);
  A2<int8_t> maskv(maskv_v.data(), ni, nj), mask_coast(mc_v.data(), ni, nj);
  A2<float> xtmp(xtmp_v.data(), ni, nj);
  int nsweeps = 0;
  for (int jinc = 1; jinc <= ninc_max; jinc++) {
    bool any0 = false;
    for (int64_t k = 0; k < n; k++) if (maskv_v[k] == 0) { any0 = true; break; }
    if (!any0) break;
    nsweeps++;
    std::fill(mc_v.begin(), mc_v.end(), (int8_t)0);
    for (int jj = 1; jj <= nj; jj++)
      for (int ji = 1; ji <= ni; ji++) {
        int jjp1 = jj + 1, jjm1 = jj - 1;
        if (jjp1 > nj) jjp1 = nj;
        if (jjm1 < 1) jjm1 = 1;
        int ji0 = ji, jip1 = ji + 1, jim1 = ji - 1;
        if (k_ew >= 0) {
          if (ji0 > ni - k_ew) ji0 = ji0 - ni + 2 * k_ew;
          if (ji0 < 1 + k_ew) ji0 = ji0 + ni - 2 * k_ew;
          if (jip1 > ni - k_ew) jip1 = jip1 - ni + 2 * k_ew;
          if (jip1 < 1 + k_ew) jip1 = jip1 + ni - 2 * k_ew;
          if (jim1 > ni - k_ew) jim1 = jim1 - ni + 2 * k_ew;
          if (jim1 < 1 + k_ew) jim1 = jim1 + ni - 2 * k_ew;
        } else {
          if (jip1 > ni) jip1 = ni;
          if (jim1 < 1) jim1 = 1;
        }
        int s = (int)maskv(jim1, jj) + (int)maskv(jip1, jj) + (int)maskv(ji0, jjm1) + (int)maskv(ji0, jjp1);
        mask_coast(ji, jj) = (int8_t)(std::min(s, 1) * (1 - (int)maskv(ji0, jj)));
      }
    int ns = std::min(jinc, 50);
    std::memcpy(xtmp_v.data(), X.p, sizeof(float) * n);
    for (int jj = 1; jj <= nj; jj++)
      for (int ji = 1; ji <= ni; ji++) {
        if (mask_coast(ji, jj) != 1) continue;
        float summsk = 0.0f, datmsk = 0.0f;
        for (int jjm = -ns; jjm <= ns; jjm++)
          for (int jim = -ns; jim <= ns; jim++) {
            int jic = ji + jim, jjc = jj + jjm;
            if (k_ew >= 0) {
              if (jic > ni - k_ew) jic = jic - ni + 2 * k_ew;
              if (jic < 1 + k_ew) jic = jic + ni - 2 * k_ew;
            }
            if (jic >= 1 && jic <= ni && jjc >= 1 && jjc <= nj) {
              // EXP(-1.*(jjm**2+jim**2)/jinc**2)*maskv : -( (1.*REAL(int)) / REAL(int) ), REAL(4)
              float arg = -((1.f * (float)(jjm * jjm + jim * jim)) / (float)(jinc * jinc));
              float zweight = m_expf(arg) * (float)maskv(jic, jjc);
              summsk = summsk + zweight;
              datmsk = datmsk + zweight * X(jic, jjc);
            }
          }
        xtmp(ji, jj) = datmsk / summsk;
      }
    std::memcpy(X.p, xtmp_v.data(), sizeof(float) * n);
    for (int64_t k = 0; k < n; k++) maskv_v[k] = (int8_t)(maskv_v[k] + mc_v[k]);
  }
  return nsweeps;
}

}  // namespace orc

using namespace orc;

extern "C" {

void orc_fill_extra_bands(int k_ew, int nx, int ny, const double* XX, const double* YY, const double* XF, double* XP4,
                          double* YP4, double* FP4, int iorca) {
  A2<double> xp(XP4, nx + 4, ny + 4), yp(YP4, nx + 4, ny + 4), fp(FP4, nx + 4, ny + 4);
  fill_extra_bands(k_ew, A2<const double>(XX, nx, ny), A2<const double>(YY, nx, ny), A2<const double>(XF, nx, ny), xp, yp, fp, iorca);
}

void orc_slopes_akima(int k_ew, int nx, int ny, const double* XX, const double* XY, const double* XF, double* dx,
                      double* dy, double* dxy) {
  A2<double> a(dx, nx, ny), b(dy, nx, ny), c(dxy, nx, ny);
  slopes_akima(k_ew, A2<const double>(XX, nx, ny), A2<const double>(XY, nx, ny), A2<const double>(XF, nx, ny), a, b, c);
}

// AKIMA_2D (src/mod_akima_2d.f90:59-239), nslab calls in a row; ixy_pos is the SAVE'd table
// (built on the first slab when first_call != 0).
int orc_akima_2d(int k_ew_per, int nx1, int ny1, const double* X10, const double* Y10, int src_1d, int nslab,
                 const float* Z1, int nx2, int ny2, const double* X20, const double* Y20, int trg_1d, float* Z2,
                 int i_orca_src, int first_call, int32_t* ixy_pos, int nthreads) {
  int64_t n1 = (int64_t)nx1 * ny1, n2 = (int64_t)nx2 * ny2;
  std::vector<double> X1v(n1), Y1v(n1), X2v(n2), Y2v(n2);
  A2<double> X1(X1v.data(), nx1, ny1), Y1(Y1v.data(), nx1, ny1), X2(X2v.data(), nx2, ny2), Y2(Y2v.data(), nx2, ny2);
  if (src_1d) { for (int jj = 1; jj <= ny1; jj++) for (int ji = 1; ji <= nx1; ji++) { X1(ji, jj) = X10[ji - 1]; Y1(ji, jj) = Y10[jj - 1]; } }
  else { std::memcpy(X1v.data(), X10, 8 * n1); std::memcpy(Y1v.data(), Y10, 8 * n1); }
  if (trg_1d) { for (int jj = 1; jj <= ny2; jj++) for (int ji = 1; ji <= nx2; ji++) { X2(ji, jj) = X20[ji - 1]; Y2(ji, jj) = Y20[jj - 1]; } }
  else { std::memcpy(X2v.data(), X20, 8 * n2); std::memcpy(Y2v.data(), Y20, 8 * n2); }
  int ni1 = nx1 + 4, nj1 = ny1 + 4;
  int64_t ne = (int64_t)ni1 * nj1;
  A3<int32_t> ixy(ixy_pos, nx2, ny2, 2);
  if (nthreads < 1) nthreads = 1;
  bool first = first_call != 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) if (!first)
  for (int s = 0; s < nslab; s++) {
    std::vector<double> zf(ne), zx(ne), zy(ne), sx(ne), sy(ne), sxy(ne), z1d(n1);
    A2<double> zFs(zf.data(), ni1, nj1), zXs(zx.data(), ni1, nj1), zYs(zy.data(), ni1, nj1);
    A2<double> slpx(sx.data(), ni1, nj1), slpy(sy.data(), ni1, nj1), slpxy(sxy.data(), ni1, nj1);
    const float* Z1s = Z1 + (int64_t)s * n1;
    float* Z2s = Z2 + (int64_t)s * n2;
    for (int64_t k = 0; k < n1; k++) z1d[k] = (double)Z1s[k];
    fill_extra_bands(k_ew_per, A2<const double>(X1v.data(), nx1, ny1), A2<const double>(Y1v.data(), nx1, ny1),
                     A2<const double>(z1d.data(), nx1, ny1), zXs, zYs, zFs, i_orca_src);
    int kewp = -1;
    if (k_ew_per > -1) kewp = k_ew_per + 4;
    slopes_akima(kewp, A2<const double>(zx.data(), ni1, nj1), A2<const double>(zy.data(), ni1, nj1),
                 A2<const double>(zf.data(), ni1, nj1), slpx, slpy, slpxy);
    double min_lon1 = zx[0], max_lon1 = zx[0], min_lat1 = zy[0], max_lat1 = zy[0];
    for (int64_t k = 1; k < ne; k++) {
      min_lon1 = std::min(min_lon1, zx[k]); max_lon1 = std::max(max_lon1, zx[k]);
      min_lat1 = std::min(min_lat1, zy[k]); max_lat1 = std::max(max_lat1, zy[k]);
    }
    double xy_range[4] = {min_lon1, max_lon1, min_lat1, max_lat1};
    if (first && s == 0) {
      for (int64_t k = 0; k < 2 * n2; k++) ixy_pos[k] = 0;
      find_nearest_akima(zXs, zYs, xy_range, A2<const double>(X2v.data(), nx2, ny2), A2<const double>(Y2v.data(), nx2, ny2), ixy);
    }
    for (int jj2 = 1; jj2 <= ny2; jj2++)
      for (int ji2 = 1; ji2 <= nx2; ji2++) {
        double px2 = X2(ji2, jj2), py2 = Y2(ji2, jj2);
        float out;
        if (((px2 >= min_lon1) && (px2 <= max_lon1)) && ((py2 >= min_lat1) && (py2 <= max_lat1))) {
          int ji1 = ixy(ji2, jj2, 1), jj1 = ixy(ji2, jj2, 2);
          px2 = px2 - zXs(ji1, jj1);
          py2 = py2 - zYs(ji1, jj1);
          double vpl[16];
          build_pol_cell(zXs, zYs, zFs, slpx, slpy, slpxy, ji1, jj1, vpl);
          out = (float)pol_val(px2, py2, vpl);
        } else out = 0.f;
        Z2s[(ji2 - 1) + (int64_t)(jj2 - 1) * nx2] = out;
      }
  }
  return 0;
}

int orc_bdrown(int k_ew, int ni, int nj, int nslab, float* X, const int32_t* mask, int mask_per_slab, int nb_inc,
               int nb_smooth, float pmin, float pmax, float pval_land, int* nsweeps, int nthreads) {
  int64_t n = (int64_t)ni * nj;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int s = 0; s < nslab; s++) {
    A2<float> Xs(X + s * n, ni, nj);
    int ns = bdrown(k_ew, Xs, mask + (mask_per_slab ? s * n : 0), nb_inc, nb_smooth, pmin, pmax, pval_land);
    if (nsweeps) nsweeps[s] = ns;
  }
  return 0;
}

int orc_smoother(int k_ew, int ni, int nj, int nslab, float* X, int nb_smooth, const int32_t* msk, int l_emp) {
  g_error = 0;
  int64_t n = (int64_t)ni * nj;
  for (int s = 0; s < nslab; s++) {
    A2<float> Xs(X + s * n, ni, nj);
    smoother(k_ew, Xs, nb_smooth, msk, l_emp != 0);
  }
  return g_error;
}

int orc_drown(int k_ew, int ni, int nj, int nslab, float* X, const int8_t* mask, int nb_inc, int* nsweeps, int nthreads) {
  // src/mod_drown.f90:110-117: the folded indices stay inside 1..ni only when ni >= 2*k_ew + 1; below that the reference
  // reads maskv out of bounds (undefined behaviour): reported, nothing computed
  if (k_ew >= 0 && ni < 2 * k_ew + 1) return 40;
  int64_t n = (int64_t)ni * nj;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int s = 0; s < nslab; s++) {
    A2<float> Xs(X + s * n, ni, nj);
    int ns = drown(k_ew, Xs, mask, nb_inc);
    if (nsweeps) nsweeps[s] = ns;
  }
  return 0;
}

// corr_vect rotation: src/corr_vect.f90:622-624,630,632 (forward) and :963,970 (inverse)
void orc_rotate_uv(int64_t n, int nslab, const float* U, const float* V, const double* xcos, const double* xsin,
                   float* Uo, float* Vo, int inverse) {
  for (int s = 0; s < nslab; s++)
    for (int64_t k = 0; k < n; k++) {
      double zU = (double)U[s * n + k], zV = (double)V[s * n + k];
      double c = xcos[k], sn = xsin[k];
      if (!inverse) {
        Uo[s * n + k] = (float)(c * zU + sn * zV);
        Vo[s * n + k] = (float)(c * zV - sn * zU);
      } else {
        Uo[s * n + k] = (float)(c * zU - sn * zV);
        Vo[s * n + k] = (float)(c * zV + sn * zU);
      }
    }
}

}  // extern "C"
