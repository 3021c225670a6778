// TEST INFRASTRUCTURE -- CPU oracle (see orc_common.hpp header).  "parity unpinned": the reference
// ships no golden vectors for this path and cannot be compiled here.
//
// Restatement of SOSIE's bilinear path, statement by statement, keeping Fortran evaluation order
// and literal kinds:
//   DISTANCE / DISTANCE_2D        src/mod_manip.f90:1443-1573
//   L_IS_GRID_REGULAR             src/mod_manip.f90:1632-1663
//   IS_POLAR_PROJ                 src/mod_manip.f90:1843-1879
//   FIND_NEAREST_POINT/EASY/TWISTED  src/mod_manip.f90:834-1440
//   EXT_NORTH_TO_90_REGG          src/mod_manip.f90:1751-1839
//   L_InPoly / PointSlope         src/mod_poly.f90:54-266
//   heading / LOCAL_COORD / det   src/mod_bilin_2d.f90:697-870
//   MAPPING_BL                    src/mod_bilin_2d.f90:366-691
//   INTERP_BL / BILIN_2D          src/mod_bilin_2d.f90:44-361
#include "orc_common.hpp"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

int g_libm = 0;
int g_error = 0;

// ------------------------------------------------------------------------------------------
// src/mod_manip.f90:1443-1504  (the SAVE'd point-A cache only avoids recomputation)
double distance(double plona, double plonb, double plata, double platb) {
  const double zpi = kPi;
  const double zconv = zpi / 180.0;
  const double zr = earth_radius_km();
  double zlatar = plata * zconv;
  double zlonar = plona * zconv;
  double zux = m_cos(zlonar) * m_cos(zlatar);
  double zuy = m_sin(zlonar) * m_cos(zlatar);
  double zuz = m_sin(zlatar);
  double zlatbr = platb * zconv;
  double zlonbr = plonb * zconv;
  double zvx = m_cos(zlonbr) * m_cos(zlatbr);
  double zvy = m_sin(zlonbr) * m_cos(zlatbr);
  double zvz = m_sin(zlatbr);
  double zpds = zux * zvx + zuy * zvy + zuz * zvz;
  if (zpds >= 1.0) return 0.0;
  return zr * m_acos(zpds);
}

// src/mod_manip.f90:1509-1573 applied to the sub-block (i1:i2, j1:j2) of Xlonb/Xlatb, written
// into the same sub-block of Xdist (as the caller at :1039 / :1369 does).
static void distance_2d_block(double plona, const A2<const double>& Xlonb, double plata,
                              const A2<const double>& Xlatb, int i1, int i2, int j1, int j2,
                              A2<double>& Xdist) {
  const double zconv = kPi / 180.0;
  const double zr = earth_radius_km();
  double zlatar = plata * zconv;
  double zlonar = plona * zconv;
  double zux = m_cos(zlonar) * m_cos(zlatar);
  double zuy = m_sin(zlonar) * m_cos(zlatar);
  double zuz = m_sin(zlatar);
  for (int jj = j1; jj <= j2; jj++) {
    for (int ji = i1; ji <= i2; ji++) {
      double zlatbr = Xlatb(ji, jj) * zconv;
      double zlonbr = Xlonb(ji, jj) * zconv;
      double zvx = m_cos(zlonbr) * m_cos(zlatbr);
      double zvy = m_sin(zlonbr) * m_cos(zlatbr);
      double zvz = m_sin(zlatbr);
      double zpds = zux * zvx + zuy * zvy + zuz * zvz;
      Xdist(ji, jj) = zr * m_acos(std::min(zpds, 1.0));
    }
  }
}

// MINLOC over a sub-block, first minimum in column-major order (Q12)
static void minloc_block(const A2<double>& X, int i1, int i2, int j1, int j2, int* io, int* jo) {
  double best = 0.0;
  bool have = false;
  int bi = 0, bj = 0;  // Fortran returns 0 when nothing is found
  for (int jj = j1; jj <= j2; jj++)
    for (int ji = i1; ji <= i2; ji++) {
      double v = X(ji, jj);
      if (v != v) continue;
      if (!have || v < best) { best = v; bi = ji; bj = jj; have = true; }
    }
  *io = bi;
  *jo = bj;
}

// src/mod_manip.f90:1632-1663
bool l_is_grid_regular(const A2<const double>& Xlon, const A2<const double>& Xlat) {
  int nx = (int)Xlon.ni, ny = (int)Xlon.nj;
  if (Xlon.sj == 0 && Xlat.si == 0) return true;  // virtual expansion of 1-D axes: all the sums are exactly 0
  const double tol = (double)1.E-12f;
  bool reg = true;
  for (int jj = 2; jj <= ny; jj++) {
    double s = 0.0;
    for (int ji = 1; ji <= nx; ji++) s += std::fabs(Xlon(ji, jj) - Xlon(ji, 1));
    if (s > tol) { reg = false; break; }
  }
  if (reg) {
    for (int ji = 1; ji <= nx; ji++) {
      double s = 0.0;
      for (int jj = 1; jj <= ny; jj++) s += std::fabs(Xlat(ji, jj) - Xlat(1, jj));
      if (s > tol) { reg = false; break; }
    }
  }
  return reg;
}

// src/mod_manip.f90:1843-1879.  Q5: the reference reads YY(iNP,jNP+-3) unchecked; the oracle
// DEFINES an out-of-range probe as "not a polar projection".
void is_polar_proj(const A2<const double>& XX, const A2<const double>& YY, bool* lPP, int* jjNP) {
  int nx = (int)YY.ni, ny = (int)YY.nj;
  *jjNP = -1;
  bool lipp = false;
  for (int jj = 1; jj <= ny && !lipp; jj++)
    for (int ji = 1; ji <= nx; ji++)
      if (YY(ji, jj) > 88.0) { lipp = true; break; }
  int jNP = 0;
  if (lipp) {
    double best = 0.0;
    int iNP = 0;
    bool have = false;
    for (int jj = 1; jj <= ny; jj++)
      for (int ji = 1; ji <= nx; ji++) {
        double v = std::fabs(YY(ji, jj) - 90.0);
        if (!have || v < best) { best = v; iNP = ji; jNP = jj; have = true; }
      }
    if (jNP - 3 < 1 || jNP + 3 > ny) {
      lipp = false;
    } else {
      lipp = (YY(iNP, jNP - 3) < YY(iNP, jNP)) && (YY(iNP, jNP + 3) < YY(iNP, jNP));
      if (lipp) {
        double d = std::fabs(XX(iNP, jNP - 3) - XX(iNP, jNP + 3));
        lipp = (d > 100.0) && (d < 220.0);
      }
    }
  }
  *lPP = lipp;
  if (lipp) *jjNP = jNP;
}

// ------------------------------------------------------------------------------------------
// src/mod_manip.f90:967-1069
static void find_nearest_easy(const A2<const double>& Xtrg, const A2<const double>& Ytrg,
                              const A2<const double>& Xsrc, const A2<const double>& Ysrc,
                              A2<int32_t>& JIpos, A2<int32_t>& JJpos, const A2<int32_t>& mask_t,
                              int j_strt_t, int j_stop_t, int jlat_icr, int elide_fill, int nthreads) {
  const int nframe_scan = 4;
  int nx_s = (int)Xsrc.ni, ny_s = (int)Xsrc.nj, nx_t = (int)Xtrg.ni, ny_t = (int)Xtrg.nj;
  for (int64_t k = 0; k < (int64_t)nx_t * ny_t; k++) { JIpos.p[k] = -1; JJpos.p[k] = -1; }
  std::vector<double> vlon(nx_s), vlat(ny_s);
  for (int i = 1; i <= nx_s; i++) vlon[i - 1] = Xsrc(i, 3);
  for (int j = 1; j <= ny_s; j++) vlat[j - 1] = Ysrc(3, j);
  const double fill = (double)1.E12f;

  // Fortran DO jj_t = j_strt_t, j_stop_t, jlat_icr
  long ntrip = ((long)j_stop_t - (long)j_strt_t + jlat_icr) / jlat_icr;
  if (ntrip < 0) ntrip = 0;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
  {
    std::vector<double> xd_store;
    // the per-target whole-array store of the reference (:1038) needs the full array; when it is
    // elided only the 9x9 window is ever touched, so a (nx_s,ny_s)-shaped virtual view suffices.
    if (!elide_fill) xd_store.resize((size_t)nx_s * ny_s);
    else xd_store.resize((size_t)(2 * nframe_scan + 1) * (2 * nframe_scan + 1));
#pragma omp for schedule(dynamic, 1)
    for (long it = 0; it < ntrip; it++) {
      int jj_t = j_strt_t + (int)it * jlat_icr;
      for (int ji_t = 1; ji_t <= nx_t; ji_t++) {
        if (mask_t(ji_t, jj_t) != 1) continue;
        double rlon = Xtrg(ji_t, jj_t);
        double rlat = Ytrg(ji_t, jj_t);
        // ip = MINLOC(ABS(vlon(:)-rlon)), jp likewise: first minimum
        int ip = 1, jp = 1;
        {
          double best = std::fabs(vlon[0] - rlon);
          for (int i = 2; i <= nx_s; i++) {
            double v = std::fabs(vlon[i - 1] - rlon);
            if (v < best) { best = v; ip = i; }
          }
          best = std::fabs(vlat[0] - rlat);
          for (int j = 2; j <= ny_s; j++) {
            double v = std::fabs(vlat[j - 1] - rlat);
            if (v < best) { best = v; jp = j; }
          }
        }
        int i1s = std::max(ip - nframe_scan, 1);
        int i2s = std::min(ip + nframe_scan, nx_s);
        int j1s = std::max(jp - nframe_scan, 1);
        int j2s = std::min(jp + nframe_scan, ny_s);
        int ji_s, jj_s;
        if (!elide_fill) {
          A2<double> Xdist(xd_store.data(), nx_s, ny_s);
          std::fill(xd_store.begin(), xd_store.end(), fill);  // Xdist = 1.E12  (:1038)
          distance_2d_block(rlon, Xsrc, rlat, Ysrc, i1s, i2s, j1s, j2s, Xdist);
          minloc_block(Xdist, i1s, i2s, j1s, j2s, &ji_s, &jj_s);
        } else {
          // same arithmetic on a window-sized scratch: shift the view so (i1s,j1s) is element (1,1)
          int wi = i2s - i1s + 1, wj = j2s - j1s + 1;
          A2<double> W(xd_store.data(), wi, wj);
          A2<const double> Xs(&Xsrc(i1s, j1s), Xsrc.ni, Xsrc.nj, Xsrc.si, Xsrc.sj), Ys(&Ysrc(i1s, j1s), Ysrc.ni, Ysrc.nj, Ysrc.si, Ysrc.sj);
          // distance_2d_block indexes (ji,jj) on both; emulate with explicit loops
          const double zconv = kPi / 180.0;
          const double zr = earth_radius_km();
          double zlatar = rlat * zconv, zlonar = rlon * zconv;
          double zux = m_cos(zlonar) * m_cos(zlatar);
          double zuy = m_sin(zlonar) * m_cos(zlatar);
          double zuz = m_sin(zlatar);
          for (int jj = 1; jj <= wj; jj++)
            for (int ji = 1; ji <= wi; ji++) {
              double zlatbr = Ys(ji, jj) * zconv;
              double zlonbr = Xs(ji, jj) * zconv;
              double zvx = m_cos(zlonbr) * m_cos(zlatbr);
              double zvy = m_sin(zlonbr) * m_cos(zlatbr);
              double zvz = m_sin(zlatbr);
              double zpds = zux * zvx + zuy * zvy + zuz * zvz;
              W(ji, jj) = zr * m_acos(std::min(zpds, 1.0));
            }
          int li, lj;
          minloc_block(W, 1, wi, 1, wj, &li, &lj);
          ji_s = li + i1s - 1;
          jj_s = lj + j1s - 1;
        }
        JIpos(ji_t, jj_t) = ji_s;
        JJpos(ji_t, jj_t) = jj_s;
        if (ji_s == 0 || jj_s == 0) g_error = 1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// src/mod_manip.f90:1073-1440
static void find_nearest_twisted(const A2<const double>& Xtrg, const A2<const double>& Ytrg,
                                 const A2<const double>& Xsrc, const A2<const double>& Ysrc,
                                 A2<int32_t>& JIpos, A2<int32_t>& JJpos, A2<int32_t>& mask_t,
                                 int j_strt_t, int j_stop_t, int jlat_icr, int elide_fill) {
  const int Nlat_split = 40;
  const double frac_emax = (double)0.51f;
  int nx_s = (int)Xsrc.ni, ny_s = (int)Xsrc.nj, nx_t = (int)Xtrg.ni, ny_t = (int)Xtrg.nj;
  int64_t ns = (int64_t)nx_s * ny_s, nt = (int64_t)nx_t * ny_t;
  std::vector<double> xd_store(ns);
  A2<double> Xdist(xd_store.data(), nx_s, ny_s);

  double y_min_s = Ysrc(1, 1), y_max_s = Ysrc(1, 1);
  for (int jj = 1; jj <= ny_s; jj++) for (int ji = 1; ji <= nx_s; ji++) { y_min_s = std::min(y_min_s, Ysrc(ji, jj)); y_max_s = std::max(y_max_s, Ysrc(ji, jj)); }
  double y_min_t = Ytrg(1, 1), y_max_t = Ytrg(1, 1);
  for (int jj = 1; jj <= ny_t; jj++) for (int ji = 1; ji <= nx_t; ji++) { y_min_t = std::min(y_min_t, Ytrg(ji, jj)); y_max_t = std::max(y_max_t, Ytrg(ji, jj)); }
  if ((y_min_t >= y_max_s) || (y_max_t <= y_min_s)) { g_error = 2; return; }

  for (int64_t k = 0; k < nt; k++) { JIpos.p[k] = -1; JJpos.p[k] = -1; }

  std::vector<double> e1v(ns, (double)40000.f), e2v(ns, (double)40000.f);
  A2<double> e1_s(e1v.data(), nx_s, ny_s), e2_s(e2v.data(), nx_s, ny_s);
  for (int jj = 1; jj <= ny_s; jj++)
    for (int ji = 1; ji <= nx_s - 1; ji++)
      e1_s(ji, jj) = distance(Xsrc(ji, jj), Xsrc(ji + 1, jj), Ysrc(ji, jj), Ysrc(ji + 1, jj)) * 1000.0;
  for (int jj = 1; jj <= ny_s - 1; jj++)
    for (int ji = 1; ji <= nx_s; ji++)
      e2_s(ji, jj) = distance(Xsrc(ji, jj), Xsrc(ji, jj + 1), Ysrc(ji, jj), Ysrc(ji, jj + 1)) * 1000.0;
  if (nx_s > 1) for (int jj = 1; jj <= ny_s; jj++) e1_s(nx_s, jj) = e1_s(nx_s - 1, jj);
  if (ny_s > 1) for (int ji = 1; ji <= nx_s; ji++) e2_s(ji, ny_s) = e2_s(ji, ny_s - 1);

  bool lNPs, lNPt;
  int jNPs, jNPt;
  is_polar_proj(Xsrc, Ysrc, &lNPs, &jNPs);
  is_polar_proj(Xtrg, Ytrg, &lNPt, &jNPt);
  bool lStupido = lNPs || lNPt;

  int nsplit = 0;  // undefined in the reference when lStupido (never assigned): oracle DEFINES 0
  std::vector<double> VB;
  std::vector<int> JV;  // J_VLAT_S(nsplit,2) column-major; element (nsplit+1,1) aliases (1,2)  (Q4)
  if (!lStupido) {
    double y_max_bnd = std::min(y_max_t, y_max_s), y_max_bnd0 = y_max_bnd;
    double y_min_bnd = std::max(y_min_t, y_min_s), y_min_bnd0 = y_min_bnd;
    y_max_bnd = std::min((double)(int)(y_max_bnd + 2.0), 90.0);   // REAL(INT(y+2.),8)
    y_min_bnd = std::max((double)(int)(y_min_bnd - 2.0), -90.0);
    y_max_bnd = std::min(f_nint(y_max_bnd / 0.5) * 0.5, 90.0);
    y_min_bnd = std::max(f_nint(y_min_bnd / 0.5) * 0.5, -90.0);
    y_max_bnd = std::max(y_max_bnd, y_max_bnd0);
    y_min_bnd = std::min(y_min_bnd, y_min_bnd0);
    // nsplit = MAX( INT( REAL(Nlat_split) * (y_max_bnd - y_min_bnd)/180. ) , 1)
    nsplit = std::max((int)(((double)(float)Nlat_split * (y_max_bnd - y_min_bnd)) / 180.0), 1);
    VB.resize(nsplit + 1);
    JV.assign(2 * nsplit, 0);
    double dy = (y_max_bnd - y_min_bnd) / (double)(float)nsplit;
    for (int jlat = 1; jlat <= nsplit + 1; jlat++) VB[jlat - 1] = y_min_bnd + (double)(float)(jlat - 1) * dy;
    if ((y_min_bnd > y_min_bnd0) || (y_max_bnd < y_max_bnd0)) { g_error = 3; return; }
    for (int jlat = 1; jlat <= nsplit; jlat++) {
      double rlat_1 = VB[jlat - 1], rlat_2 = VB[jlat];
      // jmax_band = MAXVAL(MAXLOC(Ysrc, mask=(Ysrc<=rlat_2), dim=2))
      int jmax_band = 0;
      // i1dum = MINLOC(Ysrc, mask=(Ysrc > rlat_1), dim=2); jmin_band = MINVAL(i1dum, mask=i1dum>0)
      int jmin_band = 2147483647;
      for (int ji = 1; ji <= nx_s; ji++) {
        int jx = 0, jn = 0;
        double vmax = 0, vmin = 0;
        for (int jj = 1; jj <= ny_s; jj++) {
          double v = Ysrc(ji, jj);
          if (v <= rlat_2) { if (jx == 0 || v > vmax) { vmax = v; jx = jj; } }
          if (v > rlat_1)  { if (jn == 0 || v < vmin) { vmin = v; jn = jj; } }
        }
        jmax_band = std::max(jmax_band, jx);
        if (jn > 0) jmin_band = std::min(jmin_band, jn);
      }
      JV[(jlat - 1)] = std::max(jmin_band - 1, 1);
      JV[(jlat - 1) + nsplit] = std::min(jmax_band + 1, ny_s);
    }
    for (int jlat = 1; jlat <= nsplit; jlat++)
      if (JV[(jlat - 1) + nsplit] <= JV[jlat - 1]) { g_error = 4; return; }
  }
  auto JVLAT = [&](int a, int b) -> int { return JV[(size_t)(a - 1) + (size_t)(b - 1) * nsplit]; };

  const double fill = (double)1.E12f;
  const double sqrt2f = (double)std::sqrt(2.0f);  // SQRT(2.) in REAL(4)
  double rlat_old = kRmv;
  int ji_s = 0, jj_s = 0;  // Q3: uninitialised in the reference; oracle DEFINES 0
  int jlat = 0;
  long ntrip = ((long)j_stop_t - (long)j_strt_t + jlat_icr) / jlat_icr;
  if (ntrip < 0) ntrip = 0;
  for (long it = 0; it < ntrip; it++) {
    int jj_t = j_strt_t + (int)it * jlat_icr;
    for (int ji_t = 1; ji_t <= nx_t; ji_t++) {
      if (mask_t(ji_t, jj_t) != 1) continue;
      double rlon = Xtrg(ji_t, jj_t), rlat = Ytrg(ji_t, jj_t);
      if (!lStupido) {
        if (rlat != rlat_old) {
          for (jlat = 1; jlat <= nsplit; jlat++) {
            if (rlat == VB[jlat - 1]) break;
            if ((rlat > VB[jlat - 1]) && (rlat <= VB[jlat])) break;
          }
        }
      }
      bool lagain = true;
      int niter = -1;
      if (rlat > 60.0) niter = 0;
      while (lagain) {
        int i1s, i2s, j1s, j2s;
        if (niter == -1) {
          i1s = std::max(ji_s - 5, 1);
          i2s = std::min(ji_s + 5, nx_s);
          j1s = std::max(jj_s - 5, 1);
          j2s = std::min(jj_s + 5, ny_s);
        } else {
          i1s = 1;
          i2s = nx_s;
          if (lStupido) { j1s = 1; j2s = ny_s; }
          else {
            j1s = JVLAT(std::max(jlat - niter, 1), 1);
            j2s = JVLAT(std::min(jlat + niter, nsplit), 2);
          }
        }
        if (!elide_fill) std::fill(xd_store.begin(), xd_store.end(), fill);
        distance_2d_block(rlon, Xsrc, rlat, Ysrc, i1s, i2s, j1s, j2s, Xdist);
        minloc_block(Xdist, i1s, i2s, j1s, j2s, &ji_s, &jj_s);
        if (ji_s == 0 || jj_s == 0) { g_error = 5; return; }
        double emax = std::max(e1_s(ji_s, jj_s), e2_s(ji_s, jj_s)) / 1000.0 * sqrt2f;
        if (Xdist(ji_s, jj_s) <= frac_emax * emax) {
          lagain = false;
          JIpos(ji_t, jj_t) = ji_s;
          JJpos(ji_t, jj_t) = jj_s;
        } else {
          if (niter == 0) {
            double lo1 = std::max(rlon - 0.5, 0.0), hi1 = std::min(rlon + 0.5, 360.0);
            double lo2 = std::max(rlat - 0.5, -90.0), hi2 = std::min(rlat + 0.5, 90.0);
            bool any = false;
            for (int jj = 1; jj <= ny_s && !any; jj++)
              for (int ji = 1; ji <= nx_s; ji++) {
                double xs = Xsrc(ji, jj), ys = Ysrc(ji, jj);
                if ((xs > lo1) && (xs < hi1) && (ys > lo2) && (ys < hi2)) { any = true; break; }
              }
            if (!any) lagain = false;
          }
          if (niter > nsplit / 3) lagain = false;
        }
        niter = niter + 1;
      }
      rlat_old = rlat;
    }
  }
  for (int64_t k = 0; k < nt; k++) if (JIpos.p[k] == -1) mask_t.p[k] = -1;
  for (int64_t k = 0; k < nt; k++) if (JJpos.p[k] == -1) mask_t.p[k] = -2;
}

// ------------------------------------------------------------------------------------------
// src/mod_manip.f90:834-963.  mask (optional in the reference) is always passed here.
// returns 0 = EASY used, 1 = TWISTED used
int find_nearest_point(const A2<const double>& Xtrg, const A2<const double>& Ytrg,
                       const A2<const double>& Xsrc, const A2<const double>& Ysrc, A2<int32_t>& JIp,
                       A2<int32_t>& JJp, A2<int32_t>& mask_domain_trg, int elide_fill, int nthreads,
                       int* info /* [j_strt_t, j_stop_t, jlat_icr, l_is_reg_s, l_is_reg_t] */) {
  int nx_s = (int)Xsrc.ni, ny_s = (int)Xsrc.nj, nx_t = (int)Xtrg.ni, ny_t = (int)Xtrg.nj;
  bool l_is_reg_s = l_is_grid_regular(Xsrc, Ysrc);
  bool l_is_reg_t = l_is_grid_regular(Xtrg, Ytrg);
  std::vector<int32_t> mstore(mask_domain_trg.p, mask_domain_trg.p + (int64_t)nx_t * ny_t);
  A2<int32_t> mask_ignore_t(mstore.data(), nx_t, ny_t);

  double y_min_s = Ysrc(1, 1), y_max_s = Ysrc(1, 1);
  if (Ysrc.si == 0) {  // virtual expansion of a 1-D axis: every column holds the same values
    for (int jj = 1; jj <= ny_s; jj++) { y_min_s = std::min(y_min_s, Ysrc(1, jj)); y_max_s = std::max(y_max_s, Ysrc(1, jj)); }
  } else {
    for (int jj = 1; jj <= ny_s; jj++) for (int ji = 1; ji <= nx_s; ji++) { y_min_s = std::min(y_min_s, Ysrc(ji, jj)); y_max_s = std::max(y_max_s, Ysrc(ji, jj)); }
  }

  int j_strt_t = 1, j_stop_t = ny_t, jlat_icr = 1;
  double rmin_dlat_dj = (double)-100.f;
  if (nx_t > 2) {
    // ztmp_t(:,jj) = Ytrg(:,jj) - Ytrg(:,jj-1), jj>=2 ; rtmp = SUM(ztmp_t(:,ny_t/2))
    double rtmp = 0.0;
    int jm = ny_t / 2;
    if (jm >= 2)
      for (int ji = 1; ji <= nx_t; ji++) rtmp += (Ytrg(ji, jm) - Ytrg(ji, jm - 1));
    // (jm < 2: the reference sums an uninitialised column; oracle defines the sum as 0 -> sign +)
    double sg = f_sign(1.0, rtmp);
    bool have = false;
    for (int jj = 2; jj <= ny_t; jj++)
      for (int ji = 1; ji <= nx_t; ji++) {
        double v = sg * (Ytrg(ji, jj) - Ytrg(ji, jj - 1));
        if (!have || v < rmin_dlat_dj) { rmin_dlat_dj = v; have = true; }
      }
  }
  if ((rmin_dlat_dj > (double)-1.E-12f) || l_is_reg_t) {
    int jmin_all = 2147483647, jmax_all = 0;
    for (int ji = 1; ji <= nx_t; ji++) {
      int jn = 0, jx = 0;
      double vmin = 0, vmax = 0;
      for (int jj = 1; jj <= ny_t; jj++) {
        double v = Ytrg(ji, jj);
        if (v >= y_min_s) { if (jn == 0 || v < vmin) { vmin = v; jn = jj; } }
        if (v <= y_max_s) { if (jx == 0 || v > vmax) { vmax = v; jx = jj; } }
      }
      if (jn > 0) jmin_all = std::min(jmin_all, jn);
      jmax_all = std::max(jmax_all, jx);
    }
    j_strt_t = jmin_all;
    j_stop_t = jmax_all;
    if (j_strt_t > j_stop_t) jlat_icr = -1;
  }
  if (info) { info[0] = j_strt_t; info[1] = j_stop_t; info[2] = jlat_icr; info[3] = l_is_reg_s; info[4] = l_is_reg_t; }
  // MINVAL over an empty mask is HUGE and MAXLOC of an empty mask 0: the reference then runs its row loop outside the
  // arrays (undefined).  The oracle reports it (the GPU library returns SOSIE_ERR_REF_STOP for the same inputs).
  if (j_strt_t < 1 || j_strt_t > ny_t || j_stop_t < 1 || j_stop_t > ny_t) { g_error = 6; return 0; }
  int used;
  if (l_is_reg_s) {
    find_nearest_easy(Xtrg, Ytrg, Xsrc, Ysrc, JIp, JJp, mask_ignore_t, j_strt_t, j_stop_t, jlat_icr, elide_fill, nthreads);
    used = 0;
  } else {
    find_nearest_twisted(Xtrg, Ytrg, Xsrc, Ysrc, JIp, JJp, mask_ignore_t, j_strt_t, j_stop_t, jlat_icr, elide_fill);
    used = 1;
  }
  std::memcpy(mask_domain_trg.p, mstore.data(), sizeof(int32_t) * (size_t)nx_t * ny_t);
  return used;
}

// ------------------------------------------------------------------------------------------
// src/mod_manip.f90:1751-1839
void ext_north_to_90_regg(const A2<const double>& XX, const A2<const double>& YY, const A2<const float>& XF,
                          A2<double>& XP, A2<double>& YP, A2<float>& FP, int coords, int field) {
  int nx = (int)XX.ni, ny = (int)XX.nj;
  int nyp1 = ny + 1, nyp2 = ny + 2;
  if (coords) {
    for (int jj = 1; jj <= ny; jj++)
      for (int ji = 1; ji <= nx; ji++) { XP(ji, jj) = XX(ji, jj); YP(ji, jj) = YY(ji, jj); }
    double rr = YY(nx / 2, ny);
    if (rr == 90.0) g_error = 10;
    double s = 0.0;
    for (int ji = 1; ji <= nx; ji++) { double d = YY(ji, ny) - rr; s += d * d; }
    if (s > (double)1.E-12f) g_error = 11;
    for (int ji = 1; ji <= nx; ji++) {
      XP(ji, nyp1) = XX(ji, ny);
      XP(ji, nyp2) = XX(ji, ny);
      YP(ji, nyp1) = 90.0;
      YP(ji, nyp2) = 90.0 + (YY(nx / 2, ny) - YY(nx / 2, ny - 1));
    }
  }
  if (field) {
    for (int jj = 1; jj <= ny; jj++)
      for (int ji = 1; ji <= nx; ji++) FP(ji, jj) = XF(ji, jj);
    for (int ji = 1; ji <= nx; ji++) {
      double zlon = XX(ji, ny);
      double zlon_m = f_mod(zlon + 180.0, 360.0);
      int ji_m = 1;
      double best = std::fabs(XX(1, ny) - zlon_m);
      for (int i = 2; i <= nx; i++) {
        double v = std::fabs(XX(i, ny) - zlon_m);
        if (v < best) { best = v; ji_m = i; }
      }
      FP(ji, nyp1) = 0.5f * (XF(ji, ny) + XF(ji_m, ny));
      FP(ji, nyp2) = XF(ji_m, ny - 1);
    }
  }
}

// ------------------------------------------------------------------------------------------
// src/mod_poly.f90:215-266
static void point_slope(double* pslup, float pvertxa, float pvertxb, float pvertya, float pvertyb,
                        double* pax, double* pby, double* pcnstnt) {
  double zvertxa = pvertxa, zvertxb = pvertxb, zvertya = pvertya, zvertyb = pvertyb;
  double zrise = zvertyb - zvertya;
  double zrun = zvertxb - zvertxa;
  if (zrun == 0) *pslup = 9999;
  else *pslup = zrise / zrun;
  if (std::fabs(*pslup) <= (double)0.001f) *pslup = 0.0;
  if (*pslup == 0) { *pax = *pslup; *pby = 1; *pcnstnt = zvertya; }
  else if (*pslup == 9999) { *pax = 1; *pby = 0; *pcnstnt = zvertxa; }
  else { *pax = *pslup; *pby = -1; *pcnstnt = (*pslup * zvertxa - zvertya); }
}

// src/mod_poly.f90:54-212
bool l_inpoly(const double vplon[4], const double vplat[4], double pxpoint, double pypoint) {
  float vertx[5], verty[5];
  double slope[4], ra[4], rb[4], rc[4];
  float rmaxx = (float)std::max(std::max(vplon[0], vplon[1]), std::max(vplon[2], vplon[3]));
  float rmaxy = (float)std::max(std::max(vplat[0], vplat[1]), std::max(vplat[2], vplat[3]));
  float rminx = (float)std::min(std::min(vplon[0], vplon[1]), std::min(vplon[2], vplon[3]));
  float rminy = (float)std::min(std::min(vplat[0], vplat[1]), std::min(vplat[2], vplat[3]));
  if ((rminx < 0.f) || (pxpoint < 0.0)) { g_error = 20; return false; }
  double zx, zy;
  if ((rminx < 10.f) && (rmaxx > 350.f)) {
    zx = f_sign(1.0, 180.0 - pxpoint) * std::min(pxpoint, std::fabs(pxpoint - 360.0));
    for (int k = 0; k < 4; k++)
      vertx[k] = (float)(f_sign(1.0, 180.0 - vplon[k]) * std::min(vplon[k], std::fabs(vplon[k] - 360.0)));
    rmaxx = std::max(std::max(vertx[0], vertx[1]), std::max(vertx[2], vertx[3]));
    rminx = std::min(std::min(vertx[0], vertx[1]), std::min(vertx[2], vertx[3]));
  } else {
    zx = pxpoint;
    for (int k = 0; k < 4; k++) vertx[k] = (float)vplon[k];
  }
  zy = pypoint;
  for (int k = 0; k < 4; k++) verty[k] = (float)vplat[k];
  vertx[4] = vertx[0];
  verty[4] = verty[0];
  for (int jj = 0; jj < 5; jj++) {
    if ((vertx[jj] - (float)(int)vertx[jj]) == 0) vertx[jj] = vertx[jj] + 0.001f;
    if ((verty[jj] - (float)(int)verty[jj]) == 0) verty[jj] = verty[jj] + 0.001f;
  }
  const int inumvert = 4;
  for (int ji = 0; ji < inumvert - 1; ji++)
    point_slope(&slope[ji], vertx[ji], vertx[ji + 1], verty[ji], verty[ji + 1], &ra[ji], &rb[ji], &rc[ji]);
  point_slope(&slope[3], vertx[3], vertx[0], verty[3], verty[0], &ra[3], &rb[3], &rc[3]);
  int icross = 0;
  if (!(zx <= rmaxx)) return false;
  if (!(zx >= rminx)) return false;
  if (!(zy <= rmaxy)) return false;
  if (!(zy >= rminy)) return false;
  for (int ji = 0; ji < inumvert; ji++) {
    int jn = (ji == inumvert - 1) ? 0 : ji + 1;  // the ji==inumvert special case pairs with vertex 1
    if (slope[ji] == 9999) {
      if (zx >= vertx[ji]) {
        if (((zy <= verty[ji]) && (zy > verty[jn])) || ((zy >= verty[ji]) && (zy < verty[jn]))) icross++;
      }
    } else if (slope[ji] != 0) {
      double zxpt = (rc[ji] + zy) / ra[ji];
      if (((zxpt <= vertx[ji]) && (zxpt > vertx[jn])) || ((zxpt >= vertx[ji]) && (zxpt < vertx[jn]))) {
        if (zx >= zxpt) icross++;
      }
    }
  }
  return (icross % 2) != 0;
}

// ------------------------------------------------------------------------------------------
// src/mod_bilin_2d.f90:818-870
double heading(double plona, double plonb, double plata, double platb) {
  const double zpi = kPi;
  const double zconv = zpi / 180.0;
  double xa = plona * zconv;
  double xb = plonb * zconv;
  double rr = std::max(std::fabs(m_tan(zpi / 4.0 - zconv * plata / 2.0)), kRepsilon);
  double ya = -m_log(rr);
  rr = std::max(std::fabs(m_tan(zpi / 4.0 - zconv * platb / 2.0)), kRepsilon);
  double yb = -m_log(rr);
  double xb_xa = f_mod((xb - xa), 2 * zpi);
  if (xb_xa >= zpi) xb_xa = xb_xa - 2 * zpi;
  if (xb_xa <= -zpi) xb_xa = xb_xa + 2 * zpi;
  double angled = m_atan2(xb_xa, yb - ya);
  double h = angled * 180.0 / zpi;
  if (h < 0) h = h + 360.0;
  return h;
}

// src/mod_bilin_2d.f90:802-815
static inline double det(double p1, double p2, double p3, double p4) { return p1 * p4 - p2 * p3; }

// src/mod_bilin_2d.f90:697-799
void local_coord(const double xlam[5], const double xphi[5], double* xa, double* xb, int* ipb) {
  const double zresmax = (double)0.1f;
  const int itermax = 200;
  double zxlam[5];
  *ipb = 0;
  for (int k = 0; k < 5; k++) zxlam[k] = xlam[k];
  if ((std::fabs(zxlam[1] - zxlam[4]) >= 180.0) || (std::fabs(zxlam[1] - zxlam[2]) >= 180.0) ||
      (std::fabs(zxlam[1] - zxlam[3]) >= 180.0))
    for (int k = 0; k < 5; k++) zxlam[k] = degE_to_degWE(zxlam[k]);
  double zres = 1000.0, zdlam = 0.5, zdphi = 0.5, zalpha = 0.0, zbeta = 0.0;
  int niter = 0;
  while ((zres > zresmax) && (niter < itermax)) {
    double z1 = (zxlam[2] - zxlam[1]);
    double z2 = (zxlam[1] - zxlam[4]);
    double z3 = (zxlam[3] - zxlam[2]);
    double za11 = z1 + (z2 + z3) * zbeta;
    double za12 = -z2 + (z2 + z3) * zalpha;
    double za21 = xphi[2] - xphi[1] + (xphi[1] - xphi[4] + xphi[3] - xphi[2]) * zbeta;
    double za22 = xphi[4] - xphi[1] + (xphi[1] - xphi[4] + xphi[3] - xphi[2]) * zalpha;
    double zdeta = det(za11, za12, za21, za22);
    zdeta = (f_sign(1.0, zdeta) * std::max(std::fabs(zdeta), kRepsilon));
    double zdalp = det(zdlam, za12, zdphi, za22) / zdeta;
    double zdbet = det(za11, zdlam, za21, zdphi) / zdeta;
    zres = std::sqrt(zdalp * zdalp + zdbet * zdbet);
    zalpha = zalpha + zdalp;
    zbeta = zbeta + zdbet;
    zdlam = zxlam[0] - ((1.0 - zalpha) * (1 - zbeta) * zxlam[1] + zalpha * (1 - zbeta) * zxlam[2] +
                        zalpha * zbeta * zxlam[3] + (1 - zalpha) * zbeta * zxlam[4]);
    zdphi = xphi[0] - ((1.0 - zalpha) * (1 - zbeta) * xphi[1] + zalpha * (1 - zbeta) * xphi[2] +
                       zalpha * zbeta * xphi[3] + (1 - zalpha) * zbeta * xphi[4]);
    niter = niter + 1;
  }
  if (niter >= itermax) { zalpha = 0.5; zbeta = 0.5; *ipb = 11; }
  *xa = zalpha;
  *xb = zbeta;
  if ((xphi[1] == xphi[2]) && (xphi[2] == xphi[3]) && (xphi[3] == xphi[4])) { *xa = 0.5; *xb = 0.5; *ipb = 12; }
}

// ------------------------------------------------------------------------------------------
// src/mod_bilin_2d.f90:366-691 (the NetCDF write at :686 is replaced by returning the arrays)
void mapping_bl(int k_ew_per, const A2<const double>& X1, const A2<const double>& Y1,
                const A2<const double>& lon_trg, const A2<const double>& lat_trg, int32_t* mask_inout,
                int32_t* MTRCS_p, double* ZAB_p, int16_t* IDP_p, float* dnp_p, int elide_fill, int nthreads,
                int* info) {
  const double rmv = kRmv;
  int nxi = (int)X1.ni, nyi = (int)X1.nj, nxo = (int)lon_trg.ni, nyo = (int)lon_trg.nj;
  int64_t nt = (int64_t)nxo * nyo;
  A3<double> ZAB(ZAB_p, nxo, nyo, 2);
  A3<int32_t> MTRCS(MTRCS_p, nxo, nyo, 3);
  A2<int16_t> ID_problem(IDP_p, nxo, nyo);
  A2<float> distance_to_np(dnp_p, nxo, nyo);
  std::vector<int32_t> inr(nt), jnr(nt);
  A2<int32_t> i_nrst_in(inr.data(), nxo, nyo), j_nrst_in(jnr.data(), nxo, nyo);
  for (int64_t k = 0; k < 2 * nt; k++) ZAB_p[k] = 0.0;
  for (int64_t k = 0; k < 3 * nt; k++) MTRCS_p[k] = 0;
  for (int64_t k = 0; k < nt; k++) { IDP_p[k] = 0; dnp_p[k] = 0.f; }  // Q24: dist_np defined as 0 where unwritten
  A2<int32_t> mask_ignore_trg(mask_inout, nxo, nyo);

  find_nearest_point(lon_trg, lat_trg, X1, Y1, i_nrst_in, j_nrst_in, mask_ignore_trg, elide_fill, nthreads, info);
  if (g_error) return;
  const int iqdrn0 = 0;  // Q2: never assigned in the reference (static storage => 0)

  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
  for (int jj = 1; jj <= nyo; jj++) {
    for (int ji = 1; ji <= nxo; ji++) {
      if (mask_ignore_trg(ji, jj) != 1) continue;
      double xP = lon_trg(ji, jj), yP = lat_trg(ji, jj);
      int iP = i_nrst_in(ji, jj), jP = j_nrst_in(ji, jj);
      if (!((iP != (int)rmv) && (jP != (int)rmv) && (jP < nyi))) continue;
      int iPm1 = iP - 1, iPp1 = iP + 1, jPm1 = jP - 1, jPp1 = jP + 1;
      if (iPm1 == 0) { if (k_ew_per >= 0) iPm1 = nxi - k_ew_per; }
      if (iPp1 == nxi + 1) { if (k_ew_per >= 0) iPp1 = 1 + k_ew_per; }
      if ((iPm1 < 1) || (jPm1 < 1) || (iPp1 > nxi)) continue;  // "bound problem" warning
      double lonP = f_mod(X1(iP, jP), 360.0), latP = Y1(iP, jP);
      double lonN = f_mod(X1(iP, jPp1), 360.0), latN = Y1(iP, jPp1);
      double lonE = f_mod(X1(iPp1, jP), 360.0), latE = Y1(iPp1, jP);
      double lonS = f_mod(X1(iP, jPm1), 360.0), latS = Y1(iP, jPm1);
      double lonW = f_mod(X1(iPm1, jP), 360.0), latW = Y1(iPm1, jP);
      xP = f_mod(xP, 360.0);
      double hP = heading(lonP, xP, latP, yP);
      double hN = heading(lonP, lonN, latP, latN);
      double hE = heading(lonP, lonE, latP, latE);
      double hS = heading(lonP, lonS, latP, latS);
      double hW = heading(lonP, lonW, latP, latW);
      int iqdrn = 4;
      double hPp = degE_to_degWE(hP);
      if (hN > hE) hN = hN - 360.0;
      if (hPp > hN && hPp <= hE) iqdrn = 1;
      if (hP > hE && hP <= hS) iqdrn = 2;
      if (hP > hS && hP <= hW) iqdrn = 3;
      if (hP > hW && hPp <= hN) iqdrn = 4;
      double loni[5], lati[5];
      loni[0] = xP; lati[0] = yP;
      loni[1] = lonP; lati[1] = latP;
      distance_to_np(ji, jj) = (float)distance(xP, lonP, yP, latP);
      int icpt = 0;
      bool lagain = true;
      while (lagain) {
        switch (iqdrn) {
          case 1:
            loni[2] = lonE; lati[2] = latE;
            loni[3] = f_mod(X1(iPp1, jPp1), 360.0); lati[3] = Y1(iPp1, jPp1);
            loni[4] = lonN; lati[4] = latN;
            break;
          case 2:
            loni[2] = lonS; lati[2] = latS;
            loni[3] = f_mod(X1(iPp1, jPm1), 360.0); lati[3] = Y1(iPp1, jPm1);
            loni[4] = lonE; lati[4] = latE;
            break;
          case 3:
            loni[2] = lonW; lati[2] = latW;
            loni[3] = f_mod(X1(iPm1, jPm1), 360.0); lati[3] = Y1(iPm1, jPm1);
            loni[4] = lonS; lati[4] = latS;
            break;
          case 4:
            loni[2] = lonN; lati[2] = latN;
            loni[3] = f_mod(X1(iPm1, jPp1), 360.0); lati[3] = Y1(iPm1, jPp1);
            loni[4] = lonW; lati[4] = latW;
            break;
          default: break;
        }
        for (int k = 0; k < 5; k++) if (loni[k] <= 0.0) loni[k] = loni[k] + 360.0;
        bool l_ok = l_inpoly(&loni[1], &lati[1], xP, yP);  // Q8: un-moved xP
        if ((!l_ok) && (yP < 88.0)) { icpt = icpt + 1; iqdrn = icpt; }
        else lagain = false;
        if (icpt == 5) { lagain = false; iqdrn = iqdrn0; }
      }
      double alpha, beta;
      int iproblem;
      local_coord(loni, lati, &alpha, &beta, &iproblem);
      ID_problem(ji, jj) = (int16_t)iproblem;
      MTRCS(ji, jj, 3) = iqdrn;
      ZAB(ji, jj, 1) = alpha;
      ZAB(ji, jj, 2) = beta;
    }
  }
  // whole-array fix-ups :642-683
  for (int64_t k = 0; k < nt; k++) { MTRCS_p[k] = inr[k]; MTRCS_p[k + nt] = jnr[k]; }
  double* Z1 = ZAB_p;
  double* Z2 = ZAB_p + nt;
  int32_t* M3 = MTRCS_p + 2 * nt;
  for (int64_t k = 0; k < nt; k++)
    if ((inr[k] == (int)rmv) || (jnr[k] == (int)rmv)) { Z1[k] = rmv; Z2[k] = rmv; }
  for (int64_t k = 0; k < nt; k++) if ((Z1[k] < 0.0) && (Z1[k] > -kRepsilon)) Z1[k] = 0.0;
  for (int64_t k = 0; k < nt; k++) if ((Z2[k] < 0.0) && (Z2[k] > -kRepsilon)) Z2[k] = 0.0;
  for (int64_t k = 0; k < nt; k++) if ((Z1[k] > rmv) && (Z1[k] < 0.0)) { Z1[k] = 0.5; IDP_p[k] = 4; }
  for (int64_t k = 0; k < nt; k++) if (Z1[k] > 1.0) { Z1[k] = 0.5; IDP_p[k] = 5; }
  for (int64_t k = 0; k < nt; k++) if ((Z2[k] > rmv) && (Z2[k] < 0.0)) { Z2[k] = 0.5; IDP_p[k] = 6; }
  for (int64_t k = 0; k < nt; k++) if (Z2[k] > 1.0) { Z2[k] = 0.5; IDP_p[k] = 7; }
  for (int64_t k = 0; k < nt; k++) if (M3[k] < 1) { M3[k] = 1; IDP_p[k] = 1; }
  for (int64_t k = 0; k < nt; k++) if (mask_inout[k] <= -1) IDP_p[k] = -1;
  for (int64_t k = 0; k < nt; k++) if (mask_inout[k] == 0) IDP_p[k] = -2;
  for (int64_t k = 0; k < nt; k++) if (mask_inout[k] < -2) IDP_p[k] = -3;
}

// ------------------------------------------------------------------------------------------
// src/mod_bilin_2d.f90:275-361.  Q19: i1..j4 are SAVE'd; an iqd outside 1..4 reuses them.
struct InterpBLState { int i1 = 0, j1 = 0, i2 = 0, j2 = 0, i3 = 0, j3 = 0, i4 = 0, j4 = 0; };
template <class FZ>
float interp_bl_f(InterpBLState& st, int k_ew_per, int nxi, int nyi, int jiP, int jjP, int iqd, double xa, double xb, const FZ& Z_in) {
  int jiPm1 = jiP - 1, jiPp1 = jiP + 1;
  if ((jiPm1 == 0) && (k_ew_per >= 0)) jiPm1 = nxi - k_ew_per;
  if ((jiPp1 == nxi + 1) && (k_ew_per >= 0)) jiPp1 = 1 + k_ew_per;
  int &i1 = st.i1, &j1 = st.j1, &i2 = st.i2, &j2 = st.j2, &i3 = st.i3, &j3 = st.j3, &i4 = st.i4, &j4 = st.j4;
  switch (iqd) {
    case 1: i1 = jiP; j1 = jjP; i2 = jiPp1; j2 = jjP; i3 = jiPp1; j3 = jjP + 1; i4 = jiP; j4 = jjP + 1; break;
    case 2: i1 = jiP; j1 = jjP; i2 = jiP; j2 = jjP - 1; i3 = jiPp1; j3 = jjP - 1; i4 = jiPp1; j4 = jjP; break;
    case 3: i1 = jiP; j1 = jjP; i2 = jiPm1; j2 = jjP; i3 = jiPm1; j3 = jjP - 1; i4 = jiP; j4 = jjP - 1; break;
    case 4: i1 = jiP; j1 = jjP; i2 = jiP; j2 = jjP + 1; i3 = jiPm1; j3 = jjP + 1; i4 = jiPm1; j4 = jjP; break;
    default: break;
  }
  float w1 = (float)((1.0 - xa) * (1.0 - xb));
  float w2 = (float)(xa * (1.0 - xb));
  float w3 = (float)(xa * xb);
  float w4 = (float)((1.0 - xa) * xb);
  float wup = w1 + w2 + w3 + w4;
  if (wup == 0.f) return -9998.f;
  if ((i1 < 1) || (i2 < 1) || (i3 < 1) || (i4 < 1)) return -9997.f;
  if ((j1 < 1) || (j2 < 1) || (j3 < 1) || (j4 < 1)) return -9996.f;
  if ((j1 > nyi) || (j2 > nyi) || (j3 > nyi) || (j4 > nyi)) return -9995.f;
  if ((i1 > nxi) || (i2 > nxi) || (i3 > nxi) || (i4 > nxi)) return -9994.f;
  return (Z_in(i1, j1) * w1 + Z_in(i2, j2) * w2 + Z_in(i3, j3) * w3 + Z_in(i4, j4) * w4) / wup;
}
float interp_bl(InterpBLState& st, int k_ew_per, int jiP, int jjP, int iqd, double xa, double xb,
                const A2<const float>& Z_in) {
  return interp_bl_f(st, k_ew_per, (int)Z_in.ni, (int)Z_in.nj, jiP, jjP, iqd, xa, xb, Z_in);
}

}  // namespace orc

using namespace orc;

// ==========================================================================================
// C entry points (ctypes).  All arrays Fortran column-major.
extern "C" {

void orc_set_libm(int mode) { g_libm = mode ? 1 : 0; }
int orc_get_error() { int e = g_error; g_error = 0; return e; }
int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

double orc_distance(double lona, double lonb, double lata, double latb) { return distance(lona, lonb, lata, latb); }
double orc_heading(double lona, double lonb, double lata, double latb) { return heading(lona, lonb, lata, latb); }
int orc_l_inpoly(const double* vplon, const double* vplat, double px, double py) { return l_inpoly(vplon, vplat, px, py) ? 1 : 0; }
void orc_local_coord(const double* xlam, const double* xphi, double* xa, double* xb, int* ipb) { local_coord(xlam, xphi, xa, xb, ipb); }
double orc_earth_radius() { return earth_radius_km(); }
int orc_is_grid_regular(int nx, int ny, const double* X, const double* Y) {
  return l_is_grid_regular(A2<const double>(X, nx, ny), A2<const double>(Y, nx, ny)) ? 1 : 0;
}

// pmath probes for tests/test_pmath.py : fn 0 sin 1 cos 2 tan 3 acos 4 log 5 atan2(y,x) 6 expf
void orc_pmath_eval(int fn, int64_t n, const double* x, const double* y, double* out_pm, double* out_libm) {
  for (int64_t i = 0; i < n; i++) {
    switch (fn) {
      case 0: out_pm[i] = pm_sin(x[i]); out_libm[i] = std::sin(x[i]); break;
      case 1: out_pm[i] = pm_cos(x[i]); out_libm[i] = std::cos(x[i]); break;
      case 2: out_pm[i] = pm_tan(x[i]); out_libm[i] = std::tan(x[i]); break;
      case 3: out_pm[i] = pm_acos(x[i]); out_libm[i] = std::acos(x[i]); break;
      case 4: out_pm[i] = pm_log(x[i]); out_libm[i] = std::log(x[i]); break;
      case 5: out_pm[i] = pm_atan2(y[i], x[i]); out_libm[i] = std::atan2(y[i], x[i]); break;
      default: out_pm[i] = (double)pm_expf((float)x[i]); out_libm[i] = (double)::expf((float)x[i]); break;
    }
  }
}

int orc_find_nearest_point(int nx_t, int ny_t, const double* Xtrg, const double* Ytrg, int nx_s, int ny_s,
                           const double* Xsrc, const double* Ysrc, int32_t* JIp, int32_t* JJp, int32_t* mask_inout,
                           int elide_fill, int nthreads, int* info) {
  A2<int32_t> ji(JIp, nx_t, ny_t), jj(JJp, nx_t, ny_t), mk(mask_inout, nx_t, ny_t);
  return find_nearest_point(A2<const double>(Xtrg, nx_t, ny_t), A2<const double>(Ytrg, nx_t, ny_t),
                            A2<const double>(Xsrc, nx_s, ny_s), A2<const double>(Ysrc, nx_s, ny_s), ji, jj, mk,
                            elide_fill, nthreads, info);
}

void orc_mapping_bl(int k_ew_per, int nxi, int nyi, const double* X1, const double* Y1, int nxo, int nyo,
                    const double* lon_trg, const double* lat_trg, int32_t* mask_inout, int32_t* MTRCS, double* ZAB,
                    int16_t* ID_problem, float* dist_np, int elide_fill, int nthreads, int* info) {
  mapping_bl(k_ew_per, A2<const double>(X1, nxi, nyi), A2<const double>(Y1, nxi, nyi),
             A2<const double>(lon_trg, nxo, nyo), A2<const double>(lat_trg, nxo, nyo), mask_inout, MTRCS, ZAB,
             ID_problem, dist_np, elide_fill, nthreads, info);
}

void orc_ext_north_to_90_regg(int nx, int ny, const double* XX, const double* YY, const float* XF, double* XP,
                              double* YP, float* FP) {
  A2<double> xp(XP, nx, ny + 2), yp(YP, nx, ny + 2);
  A2<float> fp(FP, nx, ny + 2);
  ext_north_to_90_regg(A2<const double>(XX, nx, ny), A2<const double>(YY, nx, ny), A2<const float>(XF, nx, ny), xp, yp,
                       fp, 1, 1);
}

float orc_interp_bl(int k_ew_per, int nxi, int nyi, int jiP, int jjP, int iqd, double xa, double xb, const float* Z_in) {
  InterpBLState st;
  return interp_bl(st, k_ew_per, jiP, jjP, iqd, xa, xb, A2<const float>(Z_in, nxi, nyi));
}

// The per-point block of the trajectory front ends: src/interp_to_ground_track.f90:718-779 (nk = 1, linear interpolation in
// time between the two surrounding model records F1, F2) and src/interp_to_hydro_section.f90:567-607 (nk levels of one
// record, F2 == NULL).  The reference rebuilds the WHOLE field xvar = xvar1 + xdum_r4*(rt - vt_mod(jtm_1)) for every track
// point (:724) and reads five of its elements; the array statement is elementwise, so those elements are evaluated on
// demand here (tests/pyref.py keeps the literal whole-array version).  Outputs are written for valid points only, as in
// the reference (the caller pre-fills them with rmissval / 0).  f_bl, f_np, fmask: (nk, n), level fastest.
void orc_track_interp(int ni, int nj, int nk, const float* F1, const float* F2, double vt1, double vt2, const int8_t* imask, int64_t n,
                      const double* rt, const int32_t* jiP, const int32_t* jjP, const int32_t* iqdv, const double* xav, const double* xbv,
                      const double* r_obs_v, double obs_lo, double obs_hi, int clip_missing, double rmissval, double* f_bl, double* f_np,
                      int8_t* fmask) {
  const int64_t nij = (int64_t)ni * nj;
  std::vector<float> xdum_r4;
  if (F2) {   // :718  xdum_r4 = (xvar2 - xvar1) / (vt_mod(jtm_2) - vt_mod(jtm_1)), REAL(4) <- REAL(8)
    xdum_r4.resize((size_t)nij * nk);
    for (int64_t q = 0; q < nij * nk; q++) xdum_r4[q] = (float)((double)(F2[q] - F1[q]) / (vt2 - vt1));
  }
  for (int64_t p = 0; p < n; p++) {
    const int iP = jiP[p], jP = jjP[p], iquadran = iqdv[p];
    const double alpha = xav[p], beta = xbv[p];
    if (!((iP > 0) && (jP > 0))) continue;
    for (int jk = 1; jk <= nk; jk++) {
      const int64_t off = (int64_t)(jk - 1) * nij;
      auto M = [&](int i, int j) { return (int)imask[off + (i - 1) + (int64_t)(j - 1) * ni]; };
      auto xvar = [&](int64_t i, int64_t j) -> float {   // :724, REAL(4) <- REAL(4) + REAL(4)*REAL(8)
        const int64_t q = off + (i - 1) + (j - 1) * ni;
        if (!F2) return F1[q];
        return (float)((double)F1[q] + (double)xdum_r4[q] * (rt[p] - vt1));
      };
      if (M(iP, jP) == 1) {
        const double r_obs = r_obs_v[p];
        const bool l_obs_ok = (r_obs > obs_lo) && (r_obs < obs_hi);
        if (l_obs_ok) {
          const int ip1 = std::min(iP + 1, ni), jp1 = std::min(jP + 1, nj), im1 = std::max(iP - 1, 1), jm1 = std::max(jP - 1, 1);
          const int idot = M(ip1, jP) + M(ip1, jp1) + M(iP, jp1) + M(im1, jp1) + M(im1, jP) + M(im1, jm1) + M(iP, jm1) + M(ip1, jm1);
          if (idot == 8) {
            InterpBLState st;
            const int64_t o = (int64_t)(jk - 1) + p * nk;
            f_np[o] = (double)xvar(iP, jP);
            double v = (double)interp_bl_f(st, -1, ni, nj, iP, jP, iquadran, alpha, beta, xvar);
            if (clip_missing && (v < (double)-9990.f)) v = rmissval;   // :787
            f_bl[o] = v;
            fmask[o] = 1;
          }
        }
      }
    }
  }
}

// BILIN_2D (src/mod_bilin_2d.f90:44-269) with the module SAVE state (IMETRICS, RAB, IPB,
// l_last_y_row_missing) and the global l_first_call_interp_routine held by the caller:
//   first_call != 0 : build the mapping (MAPPING_BL) into metrics/alphabeta/ipb/dist_np, or, if
//                     have_mapping != 0, take the caller's arrays as "read from the mapping file".
//   x10/y10: src_1d ? (nx1),(ny1) vectors : (nx1,ny1) arrays; same for the target.
//   mask_domain_trg may be NULL.  Z1: nslab slabs of (nx1,ny1); Z2: nslab slabs of (nx2,ny2); every
//   slab after the first is a non-first call, exactly like the reference's record/level loop.
int orc_bilin_2d(int k_ew_per, int nx1, int ny1, const double* X10, const double* Y10, int src_1d, int nslab,
                 const float* Z1, int nx2, int ny2, const double* X20, const double* Y20, int trg_1d, float* Z2,
                 const int32_t* mask_domain_trg, int l_reg_src, int i_orca_trg, int first_call, int have_mapping,
                 int32_t* metrics, double* alphabeta, int16_t* ipb, float* dist_np, int* l_last_y_row_missing,
                 int elide_fill, int nthreads, int* info, int virtual_src) {
  // virtual_src != 0 (1-D source axes only): the 2-D expansions X1/Y1/X1w/Y1w of :100-148 are strided VIEWS of
  // the axes instead of copies (identical values; needed to run the 21600x10800 shape on a sample of targets
  // without 8 GB of coordinate copies per call).
  g_error = 0;   // a STOP reported by an earlier call must not leak into this one
  const double rmv = kRmv;
  const bool virt = (src_1d != 0) && (virtual_src != 0);
  int64_t n1 = (int64_t)nx1 * ny1, n2 = (int64_t)nx2 * ny2;
  std::vector<double> X1v, Y1v;
  A2<const double> X1c, Y1c;
  if (virt) {
    X1c = A2<const double>(X10, nx1, ny1, 1, 0);
    Y1c = A2<const double>(Y10, nx1, ny1, 0, 1);
  } else {
    X1v.resize(n1); Y1v.resize(n1);
    A2<double> X1(X1v.data(), nx1, ny1), Y1(Y1v.data(), nx1, ny1);
    if (src_1d) {
      for (int jj = 1; jj <= ny1; jj++) for (int ji = 1; ji <= nx1; ji++) { X1(ji, jj) = X10[ji - 1]; Y1(ji, jj) = Y10[jj - 1]; }
    } else {
      std::memcpy(X1v.data(), X10, sizeof(double) * n1);
      std::memcpy(Y1v.data(), Y10, sizeof(double) * n1);
    }
    X1c = A2<const double>(X1v.data(), nx1, ny1);
    Y1c = A2<const double>(Y1v.data(), nx1, ny1);
  }
  bool l_add_extra_j = false;
  int ny1w = ny1;
  if (l_reg_src && (ny1 > 20) && (nx1 > 20)) {
    double ymx = Y1c(nx1 / 2, ny1);
    if ((ymx < 90.0) && (2.0 * ymx - Y1c(nx1 / 2, ny1 - 1) >= 90.0)) { ny1w = ny1 + 2; l_add_extra_j = true; }
  }
  int64_t n1w = (int64_t)nx1 * ny1w;
  std::vector<double> X1wv, Y1wv, latw;
  std::vector<float> Z1wv(n1w);
  A2<float> Z1w(Z1wv.data(), nx1, ny1w);
  A2<double> X1w, Y1w;          // physical working arrays (dense mode)
  A2<const double> X1wc, Y1wc;  // what the rest of the routine reads
  if (virt) {
    latw.assign(Y10, Y10 + ny1);
    if (l_add_extra_j) { latw.push_back(90.0); latw.push_back(90.0 + (Y1c(nx1 / 2, ny1) - Y1c(nx1 / 2, ny1 - 1))); }
    X1wc = A2<const double>(X10, nx1, ny1w, 1, 0);
    Y1wc = A2<const double>(latw.data(), nx1, ny1w, 0, 1);
  } else {
    X1wv.resize(n1w); Y1wv.resize(n1w);
    X1w = A2<double>(X1wv.data(), nx1, ny1w);
    Y1w = A2<double>(Y1wv.data(), nx1, ny1w);
    X1wc = A2<const double>(X1wv.data(), nx1, ny1w);
    Y1wc = A2<const double>(Y1wv.data(), nx1, ny1w);
  }

  std::vector<double> X2v(n2), Y2v(n2);
  A2<double> X2(X2v.data(), nx2, ny2), Y2(Y2v.data(), nx2, ny2);
  if (trg_1d) {
    for (int jj = 1; jj <= ny2; jj++) for (int ji = 1; ji <= nx2; ji++) { X2(ji, jj) = X20[ji - 1]; Y2(ji, jj) = Y20[jj - 1]; }
  } else {
    std::memcpy(X2v.data(), X20, sizeof(double) * n2);
    std::memcpy(Y2v.data(), Y20, sizeof(double) * n2);
  }
  std::vector<int32_t> mask_ignore(n2, 1);
  if (mask_domain_trg) std::memcpy(mask_ignore.data(), mask_domain_trg, sizeof(int32_t) * n2);

  InterpBLState st;
  const double eps5 = (double)1.E-5f;
  for (int s = 0; s < nslab; s++) {
    const float* Z1s = Z1 + (int64_t)s * n1;
    float* Z2s = Z2 + (int64_t)s * n2;
    A2<const float> Z1c(Z1s, nx1, ny1);
    // the reference redoes the extension (coords + field) on every call (:142-148)
    if (l_add_extra_j) ext_north_to_90_regg(X1c, Y1c, Z1c, X1w, Y1w, Z1w, virt ? 0 : 1, 1);
    else {
      if (!virt) {
        std::memcpy(X1wv.data(), X1v.data(), sizeof(double) * n1);
        std::memcpy(Y1wv.data(), Y1v.data(), sizeof(double) * n1);
      }
      std::memcpy(Z1wv.data(), Z1s, sizeof(float) * n1);
    }
    bool l_first = (first_call != 0) && (s == 0);
    if (l_first) {
      *l_last_y_row_missing = 0;
      if (!have_mapping) {
        mapping_bl(k_ew_per, X1wc, Y1wc,
                   A2<const double>(X2v.data(), nx2, ny2), A2<const double>(Y2v.data(), nx2, ny2), mask_ignore.data(),
                   metrics, alphabeta, ipb, dist_np, elide_fill, nthreads, info);
        if (g_error) return g_error;
      }
    }
    A2<float> Z2a(Z2s, nx2, ny2);
    for (int64_t k = 0; k < n2; k++) Z2s[k] = (float)rmv;
    // IMETRICS(:,:,1:2) = MAX(IMETRICS(:,:,1:2), 1)   (:217) -- the reference clamps its SAVE'd copy (read back
    // from the mapping file); the caller's `metrics` keeps what MAPPING_BL wrote to the file.
    std::vector<int32_t> imv(metrics, metrics + 3 * n2);
    for (int64_t k = 0; k < 2 * n2; k++) imv[k] = std::max(imv[k], 1);
    A3<int32_t> IM(imv.data(), nx2, ny2, 3);
    A3<double> RAB(alphabeta, nx2, ny2, 2);
    A2<const float> Z1wc(Z1wv.data(), nx1, ny1w);
    for (int jj = 1; jj <= ny2; jj++)
      for (int ji = 1; ji <= nx2; ji++) {
        int iP = IM(ji, jj, 1), jP = IM(ji, jj, 2), iqdrn = IM(ji, jj, 3);
        double alpha = RAB(ji, jj, 1), beta = RAB(ji, jj, 2);
        if ((std::fabs(degE_to_degWE(X1wc(iP, jP)) - degE_to_degWE(X2(ji, jj))) < eps5) &&
            (std::fabs(Y1wc(iP, jP) - Y2(ji, jj)) < eps5))
          Z2a(ji, jj) = Z1w(iP, jP);
        else
          Z2a(ji, jj) = interp_bl(st, k_ew_per, iP, jP, iqdrn, alpha, beta, Z1wc);
      }
    // Z2 = Z2*REAL(mask,4) + REAL(1-mask,4)*(-9995.) with mask reset to 1 (:212,238)  (Q7)
    for (int64_t k = 0; k < n2; k++) Z2s[k] = Z2s[k] * 1.0f + 0.0f * (-9995.f);
    if (l_first) {
      if (i_orca_trg >= 4) {
        float sum = 0.f;
        for (int ji = 1; ji <= nx2; ji++) sum += Z2a(ji, ny2);
        double rmeanv = (double)(sum / (float)nx2);
        *l_last_y_row_missing = ((rmeanv < rmv + (double)0.1f) && (rmeanv > rmv - (double)0.1f)) ? 1 : 0;
      }
    }
    if (*l_last_y_row_missing) {
      // Array-section assignments (:256-261).  For i_orca_trg==4 the two sections do NOT conform
      // (nx2/2-1 vs nx2/2+3 elements): undefined in Fortran; the oracle DEFINES "the left-hand
      // extent governs, the right-hand section is read sequentially" (gfortran's scalariser).
      auto sect_copy = [&](int l0, int lstep, int ln, int jl, int r0, int rstep, int jr) {
        std::vector<float> tmp(ln);
        for (int k = 0; k < ln; k++) {
          int ir = r0 + k * rstep;
          tmp[k] = (ir >= 1 && ir <= nx2) ? Z2a(ir, jr) : 0.f;
        }
        for (int k = 0; k < ln; k++) Z2a(l0 + k * lstep, jl) = tmp[k];
      };
      if (i_orca_trg == 4) {
        int nA = nx2 / 2 - 1;                 // 2:nx2/2
        int nB = nx2 / 2 + 3;                 // nx2:nx2-nx2/2-2:-1
        sect_copy(2, 1, nA, ny2, nx2, -1, ny2 - 2);
        sect_copy(nx2, -1, nB, ny2, 2, 1, ny2 - 2);
      }
      if (i_orca_trg == 6) {
        int nA = nx2 / 2 - 1;                 // 2:nx2/2  and  nx2-1:nx2-nx2/2+1:-1 (same count)
        sect_copy(2, 1, nA, ny2, nx2 - 1, -1, ny2 - 1);
        sect_copy(nx2 - 1, -1, nA, ny2, 2, 1, ny2 - 1);
      }
    }
  }
  return 0;
}

}  // extern "C"
