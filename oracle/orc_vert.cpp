// TEST INFRASTRUCTURE -- CPU oracle for sosie_b200 (parity unpinned by the reference: see DESIGN.md).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
//
// Vertical stage of INTERP_3D, restated statement by statement:
//   AKIMA_1D_3D + slopes_3d + extra_2_top_3d / extra_2_bottom_3d   src/mod_akima_1d.f90:243-389, 461-515, 521-606
//   lbc_lnk_2d_r4 / lbc_lnk_2d_r8 / lbc_nfd_2d                      src/mod_nemotools.f90:65-417
//   post-interpolation block of INTERP_3D                           src/mod_interp.f90:447-511
// All arithmetic REAL(4) in source order (compiled -ffp-contract=off); whole-array statements are
// evaluated right-hand side first, as Fortran requires.
#include "orc_common.hpp"

namespace orc {

// extra_2_top_3d (:521-560): Xf4, Xf5 from Xf1..Xf3, per (i,j)
static void extra_2_top_3d(float x1, float x2, float x3, float x4, float x5, const float* Xf1, const float* Xf2, const float* Xf3,
                           float* Xf4, float* Xf5, int64_t n) {
  float A = x2 - x1, B = x3 - x2, C = x4 - x3, D = x5 - x4;
  for (int64_t k = 0; k < n; k++) {
    float ALF = Xf2[k] - Xf1[k], BET = Xf3[k] - Xf2[k];
    if ((A == 0.f) || (B == 0.f) || (C == 0.f)) { Xf4[k] = Xf3[k]; Xf5[k] = Xf3[k]; }
    else {
      float v4 = C * (2 * BET / B - ALF / A) + Xf3[k];
      Xf4[k] = v4;
      Xf5[k] = v4 + v4 * D / C + BET * D / B - ALF * D / A - Xf3[k] * D / C;
    }
  }
}
// extra_2_bottom_3d (:563-606): Xf2, Xf1 from Xf5, Xf4, Xf3
static void extra_2_bottom_3d(float x5, float x4, float x3, float x2, float x1, const float* Xf5, const float* Xf4, const float* Xf3,
                              float* Xf2, float* Xf1, int64_t n) {
  float A = x4 - x5, B = x3 - x4, C = x2 - x3, D = x1 - x2;
  for (int64_t k = 0; k < n; k++) {
    float ALF = Xf4[k] - Xf5[k], BET = Xf3[k] - Xf4[k];
    if ((A == 0.f) || (B == 0.f) || (C == 0.f)) { Xf2[k] = Xf3[k]; Xf1[k] = Xf3[k]; }
    else {
      float v2 = C * (2 * BET / B - ALF / A) + Xf3[k];
      Xf2[k] = v2;
      Xf1[k] = v2 + v2 * D / C + BET * D / B - ALF * D / A - Xf3[k] * D / C;
    }
  }
}

// slopes_3d (:461-515); vz(1..nz), Xf / slope (n, nz) with n = nx*ny
static void slopes_3d(const float* vz, const float* Xf, float* slope, int64_t n, int nz) {
  auto F = [&](int k) { return Xf + (int64_t)(k - 1) * n; };
  auto S = [&](int k) { return slope + (int64_t)(k - 1) * n; };
  auto Z = [&](int k) { return vz[k - 1]; };
  for (int k = 3; k <= nz - 2; k++) {
    for (int64_t p = 0; p < n; p++) {
      float m1 = (F(k - 1)[p] - F(k - 2)[p]) / (Z(k - 1) - Z(k - 2));
      float m2 = (F(k)[p] - F(k - 1)[p]) / (Z(k) - Z(k - 1));
      float m3 = (F(k + 1)[p] - F(k)[p]) / (Z(k + 1) - Z(k));
      float m4 = (F(k + 2)[p] - F(k + 1)[p]) / (Z(k + 2) - Z(k + 1));
      if ((m1 == m2) && (m3 == m4)) S(k)[p] = 0.5f * (m2 + m3);
      else S(k)[p] = (std::fabs(m4 - m3) * m2 + std::fabs(m2 - m1) * m3) / (std::fabs(m4 - m3) + std::fabs(m2 - m1));
    }
  }
  for (int64_t p = 0; p < n; p++) {
    float m2 = (F(2)[p] - F(1)[p]) / (Z(2) - Z(1));
    float m3 = (F(3)[p] - F(2)[p]) / (Z(3) - Z(2));
    S(2)[p] = 0.5f * (m2 + m3);
    m2 = (F(nz - 1)[p] - F(nz - 2)[p]) / (Z(nz - 1) - Z(nz - 2));
    m3 = (F(nz)[p] - F(nz - 1)[p]) / (Z(nz) - Z(nz - 1));
    S(nz - 1)[p] = 0.5f * (m2 + m3);
    m3 = (F(2)[p] - F(1)[p]) / (Z(2) - Z(1));
    S(1)[p] = m3;
    m2 = (F(nz)[p] - F(nz - 1)[p]) / (Z(nz) - Z(nz - 1));
    S(nz)[p] = m2;
  }
}

// AKIMA_1D_3D (:243-389)
static void akima_1d_3d(int nx, int ny, int nk1, const float* vz1, const float* xf1, int nk2, const float* vz2, float* xf2,
                        float rfill_val) {
  const int64_t n = (int64_t)nx * ny;
  const int nk1e = nk1 + 4;
  std::vector<float> vz1e(nk1e + 1), xf1e((size_t)n * nk1e), xslp((size_t)n * nk1e);
  auto VZ1 = [&](int k) { return vz1[k - 1]; };
  auto VZ2 = [&](int k) { return vz2[k - 1]; };
  auto E = [&](int k) -> float& { return vz1e[k]; };  // 1-based
  auto FE = [&](int k) { return xf1e.data() + (int64_t)(k - 1) * n; };
  auto SL = [&](int k) { return xslp.data() + (int64_t)(k - 1) * n; };
  auto F1 = [&](int k) { return xf1 + (int64_t)(k - 1) * n; };
  auto F2 = [&](int k) { return xf2 + (int64_t)(k - 1) * n; };
  for (int k = 1; k <= nk1; k++) { E(k + 2) = VZ1(k); std::memcpy(FE(k + 2), F1(k), sizeof(float) * n); }
  E(2) = E(4) - (E(5) - E(3));
  E(1) = E(3) - (E(5) - E(3));
  E(nk1e - 1) = E(nk1e - 3) + E(nk1e - 2) - E(nk1e - 4);
  E(nk1e) = E(nk1e - 2) + E(nk1e - 2) - E(nk1e - 4);
  extra_2_bottom_3d(E(5), E(4), E(3), E(2), E(1), FE(5), FE(4), FE(3), FE(2), FE(1), n);
  extra_2_top_3d(E(nk1e - 4), E(nk1e - 3), E(nk1e - 2), E(nk1e - 1), E(nk1e), FE(nk1e - 4), FE(nk1e - 3), FE(nk1e - 2), FE(nk1e - 1),
                 FE(nk1e), n);
  slopes_3d(vz1e.data() + 1, xf1e.data(), xslp.data(), n, nk1e);
  for (int jk2 = 1; jk2 <= nk2; jk2++) {
    bool l_do_interp = true, lfnd = false;
    int jk1 = 1, jp1 = 0, jp2 = 0;
    if (VZ2(jk2) < VZ1(1)) { std::memcpy(F2(jk2), F1(1), sizeof(float) * n); l_do_interp = false; }
    while ((!lfnd) && l_do_interp) {
      if (jk1 == nk1) break;
      if ((VZ2(jk2) > VZ1(jk1)) && (VZ2(jk2) <= VZ1(jk1 + 1))) {
        int jk1e = jk1 + 2;
        jp1 = jk1e; jp2 = jk1e + 1;
        lfnd = true;
      } else {
        if (VZ2(jk2) == VZ1(jk1)) { std::memcpy(F2(jk2), F1(jk1), sizeof(float) * n); l_do_interp = false; break; }
        else if (VZ2(jk2) == VZ1(jk1 + 1)) { std::memcpy(F2(jk2), F1(jk1 + 1), sizeof(float) * n); l_do_interp = false; break; }
      }
      jk1 = jk1 + 1;
    }
    if ((jk1 == nk1) && (!lfnd)) { std::memcpy(F2(jk2), F1(nk1), sizeof(float) * n); l_do_interp = false; }
    if (!lfnd) { for (int64_t p = 0; p < n; p++) F2(jk2)[p] = rfill_val; l_do_interp = false; }
    if (l_do_interp) {
      float dz = 1.f / (E(jp2) - E(jp1));
      float dz2 = dz * dz;
      float dzt = VZ2(jk2) - E(jp1);
      float dzt2 = dzt * dzt;
      const float dv = E(jp2) - E(jp1);
      for (int64_t p = 0; p < n; p++) {
        float a0 = FE(jp1)[p];
        float a1 = SL(jp1)[p];
        float a2 = (3 * (FE(jp2)[p] - FE(jp1)[p]) * dz - 2 * SL(jp1)[p] - SL(jp2)[p]) * dz;
        float a3 = (SL(jp1)[p] + SL(jp2)[p] - 2 * (FE(jp2)[p] - FE(jp1)[p]) / dv) * dz2;
        F2(jk2)[p] = a0 + a1 * dzt + a2 * dzt2 + a3 * dzt2 * dzt;
      }
    }
  }
  // SURFACE persistence (:372-377)
  int jk2 = 1;
  while ((VZ2(jk2) < VZ1(1)) && (jk2 < nk2)) { std::memcpy(F2(jk2), F1(1), sizeof(float) * n); jk2 = jk2 + 1; }
  // BOTTOM persistence (:381-385); the reference reads vz2(0) when every target level is deeper: DEFINED stop at 0
  jk2 = nk2;
  while ((jk2 > 0) && (VZ2(jk2) > VZ1(nk1))) { std::memcpy(F2(jk2), F1(nk1), sizeof(float) * n); jk2 = jk2 - 1; }
}

// lbc_nfd_2d (src/mod_nemotools.f90:199-417) on a REAL(8) work array, jpni = 1
static void lbc_nfd_2d(int npolj, A2<double> pt2d, char cd_type, double psgn) {
  const int jpiglo = (int)pt2d.ni, ijpj = (int)pt2d.nj, ipr2dj = 0, ijpjm1 = ijpj - 1;
  switch (npolj) {
    case 3: case 4:
      switch (cd_type) {
        case 'T': case 'W':
          for (int jl = 0; jl <= ipr2dj; jl++)
            for (int ji = 2; ji <= jpiglo; ji++) { int ijt = jpiglo - ji + 2; pt2d(ji, ijpj + jl) = psgn * pt2d(ijt, ijpj - 2 - jl); }
          pt2d(1, ijpj) = psgn * pt2d(3, ijpj - 2);
          for (int ji = jpiglo / 2 + 1; ji <= jpiglo; ji++) { int ijt = jpiglo - ji + 2; pt2d(ji, ijpj - 1) = psgn * pt2d(ijt, ijpj - 1); }
          break;
        case 'U':
          for (int jl = 0; jl <= ipr2dj; jl++)
            for (int ji = 1; ji <= jpiglo - 1; ji++) { int iju = jpiglo - ji + 1; pt2d(ji, ijpj + jl) = psgn * pt2d(iju, ijpj - 2 - jl); }
          pt2d(1, ijpj) = psgn * pt2d(2, ijpj - 2);
          pt2d(jpiglo, ijpj) = psgn * pt2d(jpiglo - 1, ijpj - 2);
          pt2d(1, ijpj - 1) = psgn * pt2d(jpiglo, ijpj - 1);
          for (int ji = jpiglo / 2; ji <= jpiglo - 1; ji++) { int iju = jpiglo - ji + 1; pt2d(ji, ijpjm1) = psgn * pt2d(iju, ijpjm1); }
          break;
        case 'V':
          for (int jl = -1; jl <= ipr2dj; jl++)
            for (int ji = 2; ji <= jpiglo; ji++) { int ijt = jpiglo - ji + 2; pt2d(ji, ijpj + jl) = psgn * pt2d(ijt, ijpj - 3 - jl); }
          pt2d(1, ijpj) = psgn * pt2d(3, ijpj - 3);
          break;
        case 'F':
          for (int jl = -1; jl <= ipr2dj; jl++)
            for (int ji = 1; ji <= jpiglo - 1; ji++) { int iju = jpiglo - ji + 1; pt2d(ji, ijpj + jl) = psgn * pt2d(iju, ijpj - 3 - jl); }
          pt2d(1, ijpj) = psgn * pt2d(2, ijpj - 3);
          pt2d(jpiglo, ijpj) = psgn * pt2d(jpiglo - 1, ijpj - 3);
          pt2d(jpiglo, ijpj - 1) = psgn * pt2d(jpiglo - 1, ijpj - 2);
          pt2d(1, ijpj - 1) = psgn * pt2d(2, ijpj - 2);
          break;
        default: g_error |= 64; break;  // ice points: not used by SOSIE's callers
      }
      break;
    case 5: case 6:
      switch (cd_type) {
        case 'T': case 'W':
          for (int jl = 0; jl <= ipr2dj; jl++)
            for (int ji = 1; ji <= jpiglo; ji++) { int ijt = jpiglo - ji + 1; pt2d(ji, ijpj + jl) = psgn * pt2d(ijt, ijpj - 1 - jl); }
          break;
        case 'U':
          for (int jl = 0; jl <= ipr2dj; jl++)
            for (int ji = 1; ji <= jpiglo - 1; ji++) { int iju = jpiglo - ji; pt2d(ji, ijpj + jl) = psgn * pt2d(iju, ijpj - 1 - jl); }
          pt2d(jpiglo, ijpj) = psgn * pt2d(1, ijpj - 1);
          break;
        case 'V':
          for (int jl = 0; jl <= ipr2dj; jl++)
            for (int ji = 1; ji <= jpiglo; ji++) { int ijt = jpiglo - ji + 1; pt2d(ji, ijpj + jl) = psgn * pt2d(ijt, ijpj - 2 - jl); }
          for (int ji = jpiglo / 2 + 1; ji <= jpiglo; ji++) { int ijt = jpiglo - ji + 1; pt2d(ji, ijpjm1) = psgn * pt2d(ijt, ijpjm1); }
          break;
        case 'F':
          for (int jl = 0; jl <= ipr2dj; jl++)
            for (int ji = 1; ji <= jpiglo - 1; ji++) { int iju = jpiglo - ji; pt2d(ji, ijpj + jl) = psgn * pt2d(iju, ijpj - 2 - jl); }
          pt2d(jpiglo, ijpj) = psgn * pt2d(1, ijpj - 2);
          for (int ji = jpiglo / 2 + 1; ji <= jpiglo - 1; ji++) { int iju = jpiglo - ji; pt2d(ji, ijpjm1) = psgn * pt2d(iju, ijpjm1); }
          break;
        default: g_error |= 64; break;
      }
      break;
    default:
      switch (cd_type) {
        case 'T': case 'U': case 'V': case 'W':
          for (int ji = 1; ji <= jpiglo; ji++) { pt2d(ji, 1) = 0.0; pt2d(ji, ijpj) = 0.0; }
          break;
        case 'F':
          for (int ji = 1; ji <= jpiglo; ji++) pt2d(ji, ijpj) = 0.0;
          break;
        default: g_error |= 64; break;
      }
      break;
  }
}

// lbc_lnk_2d_r8 (:65-150)
static void lbc_lnk_2d_r8(int nperio, A2<double> pt2d, char cd_type, double psgn, bool has_pval, double pval) {
  const double zland = has_pval ? pval : 0.0;
  const int jpi = (int)pt2d.ni, jpj = (int)pt2d.nj, jpim1 = jpi - 1;
  switch (nperio) {
    case 1: case 4: case 6:
      for (int j = 1; j <= jpj; j++) pt2d(1, j) = pt2d(jpim1, j);
      for (int j = 1; j <= jpj; j++) pt2d(jpi, j) = pt2d(2, j);
      break;
    default:
      switch (cd_type) {
        case 'T': case 'U': case 'V': case 'W':
          for (int j = 1; j <= jpj; j++) { pt2d(1, j) = zland; }
          for (int j = 1; j <= jpj; j++) { pt2d(jpi, j) = zland; }
          break;
        case 'F':
          for (int j = 1; j <= jpj; j++) pt2d(jpi, j) = zland;
          break;
        default: break;
      }
      break;
  }
  switch (nperio) {
    case 2:
      switch (cd_type) {
        case 'T': case 'U': case 'W':
          for (int i = 1; i <= jpi; i++) pt2d(i, 1) = pt2d(i, 3);
          for (int i = 1; i <= jpi; i++) pt2d(i, jpj) = zland;
          break;
        case 'V': case 'F':
          for (int i = 1; i <= jpi; i++) pt2d(i, 1) = psgn * pt2d(i, 2);
          for (int i = 1; i <= jpi; i++) pt2d(i, jpj) = zland;
          break;
        default: break;
      }
      break;
    case 3: case 4: case 5: case 6:
      switch (cd_type) {
        case 'T': case 'U': case 'V': case 'W': case 'I':
          for (int i = 1; i <= jpi; i++) pt2d(i, 1) = zland;
          break;
        default: break;
      }
      lbc_nfd_2d(nperio, pt2d, cd_type, psgn);
      break;
    default:
      switch (cd_type) {
        case 'T': case 'U': case 'V': case 'W':
          for (int i = 1; i <= jpi; i++) pt2d(i, 1) = zland;
          for (int i = 1; i <= jpi; i++) pt2d(i, jpj) = zland;
          break;
        case 'F':
          for (int i = 1; i <= jpi; i++) pt2d(i, jpj) = zland;
          break;
        default: break;
      }
      break;
  }
}

// lbc_lnk_2d_r4 (:154-190): through a REAL(8) copy
static void lbc_lnk_2d_r4(int nperio, float* X, int jpi, int jpj, char cd_type, double psgn, bool has_pval, double pval) {
  std::vector<double> w((size_t)jpi * jpj);
  for (size_t k = 0; k < w.size(); k++) w[k] = (double)X[k];
  lbc_lnk_2d_r8(nperio, A2<double>(w.data(), jpi, jpj), cd_type, psgn, has_pval, pval);
  for (size_t k = 0; k < w.size(); k++) X[k] = (float)w[k];
}

// extra_1_east_r8 / extra_1_west_r8 (src/mod_manip.f90:465-525)
static inline double extra_1(double xa, double xb, double xc, double xd, double ya, double yb, double yc) {
  const double A = xb - xa, B = xc - xb, C = xd - xc;
  const double ALF = yb - ya, BET = yc - yb;
  if ((A == 0.) || (B == 0.) || (C == 0.)) return yc;
  return C * (2 * BET / B - ALF / A) + yc;
}

// angle2 (src/mod_nemotools.f90:607-823).  rpi = 3.141592653 is a default-REAL literal (Q1) -> 3.1415927410125732.
// Cells the reference leaves unassigned (allocated, never written: column 1 before the boundary treatment, the south
// row of the F arrays on a folded grid) are DEFINED as 0.
static void angle2(int nperio, int nx, int ny, const double* plamt, const double* pphit, const double* plamu, const double* pphiu,
                   const double* plamv, const double* pphiv, const double* plamf, const double* pphif, double* gcost, double* gsint,
                   double* gcosu, double* gsinu, double* gcosv, double* gsinv, double* gcosf, double* gsinf) {
  const double rpi = (double)3.141592653f, rad = rpi / (double)180.0f;
  const size_t n = (size_t)nx * ny;
  double* outs[8] = {gcost, gsint, gcosu, gsinu, gcosv, gsinv, gcosf, gsinf};
  for (auto* o : outs) std::fill(o, o + n, 0.0);
  A2<const double> LT(plamt, nx, ny), PT(pphit, nx, ny), LU(plamu, nx, ny), PU(pphiu, nx, ny), LV(plamv, nx, ny), PV(pphiv, nx, ny),
      LF(plamf, nx, ny), PF(pphif, nx, ny);
  A2<double> CT(gcost, nx, ny), ST(gsint, nx, ny), CU(gcosu, nx, ny), SU(gsinu, nx, ny), CV(gcosv, nx, ny), SV(gsinv, nx, ny),
      CF(gcosf, nx, ny), SF(gsinf, nx, ny);
  double mx = plamt[0], mn = plamt[0];
  for (size_t k = 1; k < n; k++) { mx = std::max(mx, plamt[k]); mn = std::min(mn, plamt[k]); }
  const bool l_cut_180 = (mx <= 180.) && (mx > 175.) && (mn >= -180.) && (mn < -175.);
  auto T = [&](double zphi) { return m_tan(rpi / 4. - rad * zphi / 2.); };
  for (int jj = 2; jj <= ny - 1; jj++) {
    for (int ji = 2; ji <= nx; ji++) {
      double zlamt = LT(ji, jj), zlamu = LU(ji, jj), zlamv = LV(ji, jj), zlamf = LF(ji, jj);
      double zlamvjm1 = LV(ji, jj - 1), zlamfjm1 = LF(ji, jj - 1), zlamfim1 = LF(ji - 1, jj), zlamujp1 = LU(ji, jj + 1);
      if (l_cut_180) {
        if ((std::fabs(zlamt) > 170.) || (std::fabs(zlamu) > 170.) || (std::fabs(zlamv) > 170.) || (std::fabs(zlamf) > 170.)) {
          zlamt = f_mod(360. + LT(ji, jj), 360.); zlamu = f_mod(360. + LU(ji, jj), 360.);
          zlamv = f_mod(360. + LV(ji, jj), 360.); zlamf = f_mod(360. + LF(ji, jj), 360.);
          zlamvjm1 = f_mod(360. + LV(ji, jj - 1), 360.); zlamfjm1 = f_mod(360. + LF(ji, jj - 1), 360.);
          zlamfim1 = f_mod(360. + LF(ji - 1, jj), 360.); zlamujp1 = f_mod(360. + LU(ji, jj + 1), 360.);
        }
      }
      double zphi = PT(ji, jj);
      const double zxnpt = 0. - 2. * m_cos(rad * zlamt) * T(zphi), zynpt = 0. - 2. * m_sin(rad * zlamt) * T(zphi);
      const double znnpt = zxnpt * zxnpt + zynpt * zynpt;
      zphi = PU(ji, jj);
      const double zxnpu = 0. - 2. * m_cos(rad * zlamu) * T(zphi), zynpu = 0. - 2. * m_sin(rad * zlamu) * T(zphi);
      const double znnpu = zxnpu * zxnpu + zynpu * zynpu;
      zphi = PV(ji, jj);
      const double zxnpv = 0. - 2. * m_cos(rad * zlamv) * T(zphi), zynpv = 0. - 2. * m_sin(rad * zlamv) * T(zphi);
      const double znnpv = zxnpv * zxnpv + zynpv * zynpv;
      zphi = PF(ji, jj);
      const double zxnpf = 0. - 2. * m_cos(rad * zlamf) * T(zphi), zynpf = 0. - 2. * m_sin(rad * zlamf) * T(zphi);
      const double znnpf = zxnpf * zxnpf + zynpf * zynpf;
      zphi = PV(ji, jj); double zphh = PV(ji, jj - 1);
      const double zxvvt = 2. * m_cos(rad * zlamv) * T(zphi) - 2. * m_cos(rad * zlamvjm1) * T(zphh);
      const double zyvvt = 2. * m_sin(rad * zlamv) * T(zphi) - 2. * m_sin(rad * zlamvjm1) * T(zphh);
      double znvvt = std::sqrt(znnpt * (zxvvt * zxvvt + zyvvt * zyvvt));
      znvvt = std::max(znvvt, (double)1.e-14f);
      zphi = PF(ji, jj); zphh = PF(ji, jj - 1);
      const double zxffu = 2. * m_cos(rad * zlamf) * T(zphi) - 2. * m_cos(rad * zlamfjm1) * T(zphh);
      const double zyffu = 2. * m_sin(rad * zlamf) * T(zphi) - 2. * m_sin(rad * zlamfjm1) * T(zphh);
      double znffu = std::sqrt(znnpu * (zxffu * zxffu + zyffu * zyffu));
      znffu = std::max(znffu, (double)1.e-14f);
      zphi = PF(ji, jj); zphh = PF(ji - 1, jj);
      const double zxffv = 2. * m_cos(rad * zlamf) * T(zphi) - 2. * m_cos(rad * zlamfim1) * T(zphh);
      const double zyffv = 2. * m_sin(rad * zlamf) * T(zphi) - 2. * m_sin(rad * zlamfim1) * T(zphh);
      double znffv = std::sqrt(znnpv * (zxffv * zxffv + zyffv * zyffv));
      znffv = std::max(znffv, (double)1.e-14f);
      zphi = PU(ji, jj + 1); zphh = PU(ji, jj);
      const double zxuuf = 2. * m_cos(rad * zlamujp1) * T(zphi) - 2. * m_cos(rad * zlamu) * T(zphh);
      const double zyuuf = 2. * m_sin(rad * zlamujp1) * T(zphi) - 2. * m_sin(rad * zlamu) * T(zphh);
      double znuuf = std::sqrt(znnpf * (zxuuf * zxuuf + zyuuf * zyuuf));
      znuuf = std::max(znuuf, (double)1.e-14f);
      ST(ji, jj) = (zxnpt * zyvvt - zynpt * zxvvt) / znvvt;
      CT(ji, jj) = (zxnpt * zxvvt + zynpt * zyvvt) / znvvt;
      SU(ji, jj) = (zxnpu * zyffu - zynpu * zxffu) / znffu;
      CU(ji, jj) = (zxnpu * zxffu + zynpu * zyffu) / znffu;
      SF(ji, jj) = (zxnpf * zyuuf - zynpf * zxuuf) / znuuf;
      CF(ji, jj) = (zxnpf * zxuuf + zynpf * zyuuf) / znuuf;
      SV(ji, jj) = (zxnpv * zxffv + zynpv * zyffv) / znffv;
      CV(ji, jj) = -(zxnpv * zyffv - zynpv * zxffv) / znffv;
    }
  }
  if (nperio > 0) {
    lbc_lnk_2d_r8(nperio, CT, 'T', -1.0, false, 0.0); lbc_lnk_2d_r8(nperio, ST, 'T', -1.0, false, 0.0);
    lbc_lnk_2d_r8(nperio, CU, 'U', -1.0, false, 0.0); lbc_lnk_2d_r8(nperio, SU, 'U', -1.0, false, 0.0);
    lbc_lnk_2d_r8(nperio, CV, 'V', -1.0, false, 0.0); lbc_lnk_2d_r8(nperio, SV, 'V', -1.0, false, 0.0);
    lbc_lnk_2d_r8(nperio, CF, 'F', -1.0, false, 0.0); lbc_lnk_2d_r8(nperio, SF, 'F', -1.0, false, 0.0);
  } else {
    A2<const double>* PH[4] = {&PT, &PU, &PV, &PF};
    A2<const double>* LA[4] = {&LT, &LU, &LV, &LF};
    A2<double>* G[8] = {&CT, &ST, &CU, &SU, &CV, &SV, &CF, &SF};
    for (int ji = 1; ji <= nx; ji++)      // northernmost row
      for (int q = 0; q < 8; q++) {
        A2<const double>& P = *PH[q / 2]; A2<double>& g = *G[q];
        g(ji, ny) = extra_1(P(ji, ny - 3), P(ji, ny - 2), P(ji, ny - 1), P(ji, ny), g(ji, ny - 3), g(ji, ny - 2), g(ji, ny - 1));
      }
    for (int ji = 1; ji <= nx; ji++)      // southernmost row
      for (int q = 0; q < 8; q++) {
        A2<const double>& P = *PH[q / 2]; A2<double>& g = *G[q];
        g(ji, 1) = extra_1(P(ji, 4), P(ji, 3), P(ji, 2), P(ji, 1), g(ji, 4), g(ji, 3), g(ji, 2));
      }
    for (int jj = 1; jj <= ny; jj++)      // westernmost column
      for (int q = 0; q < 8; q++) {
        A2<const double>& L = *LA[q / 2]; A2<double>& g = *G[q];
        g(1, jj) = extra_1(L(4, jj), L(3, jj), L(2, jj), L(1, jj), g(4, jj), g(3, jj), g(2, jj));
      }
  }
}


// ---------------------------------------------------------------------------------------------------
// AKIMA_1D_1D (src/mod_akima_1d.f90:26-232) with slopes_1d (:392-455) and extra_2_west_r4 / extra_2_east_r4
// (src/mod_manip.f90:571-587, 634-649): one column.  vz1, vf1, vz2 are REAL(8); the extended arrays vz1e, vf1e and the
// slopes are REAL(4); the polynomial is evaluated in REAL(8) with the REAL(4) sub-expressions the source implies.
// Returns 0 = done, 1 = every source value is missing ("DOING NOTHING", vf2 stays -6666), < 0 = the reference's
// behaviour is undefined for this input (array-shape mismatch or reads of unset elements):
//   -1  missing values not confined to "a run at the start (optional) and a run at the end" (:88-95)
//   -2  fewer than 3 valid source points (vz1e(5) unset)
int akima_1d_1d(int nk1, const double* vz1_, const double* vf1_, int nk2, const double* vz2_, float* vf2_, bool has_rmv, double rmv,
                bool l_sp, bool l_bp) {
  auto vz1 = [&](int k) { return vz1_[k - 1]; };
  auto vf1 = [&](int k) { return vf1_[k - 1]; };
  auto vz2 = [&](int k) { return vz2_[k - 1]; };
  for (int k = 0; k < nk2; k++) vf2_[k] = -6666.0f;
  int nz1 = nk1, k1_1 = 1, k1_n = nk1, nmissv = 0;
  std::vector<int> imiss1;
  if (has_rmv) {
    imiss1.assign(nk1 + 1, 0);
    for (int k = 1; k <= nk1; k++) if (vf1(k) == rmv) { imiss1[k] = 1; nmissv++; }
  }
  if (!(nmissv < nk1)) return 1;
  if (nmissv > 0) {
    k1_1 = 0;
    for (int k = 1; k <= nk1; k++) if (imiss1[k] == 0) { k1_1 = k; break; }            // FINDLOC(imiss1, 0)
    int f = 0;
    for (int k = k1_1; k <= nk1; k++) if (imiss1[k] == 1) { f = k - k1_1 + 1; break; }  // FINDLOC(imiss1(k1_1:), 1)
    k1_n = f + k1_1 - 2;
    nz1 = nk1 - nmissv;
    if (k1_n - k1_1 + 1 != nz1) return -1;
  }
  if (nz1 < 3) return -2;
  const int nk1e = nz1 + 4;
  std::vector<float> vz1e(nk1e + 1), vf1e(nk1e + 1), vslp(nk1e + 1);
  for (int k = 3; k <= nz1 + 2; k++) { vz1e[k] = (float)vz1(k1_1 + k - 3); vf1e[k] = (float)vf1(k1_1 + k - 3); }
  vz1e[2] = vz1e[4] - (vz1e[5] - vz1e[3]);
  vz1e[1] = vz1e[3] - (vz1e[5] - vz1e[3]);
  vz1e[nk1e - 1] = vz1e[nk1e - 3] + vz1e[nk1e - 2] - vz1e[nk1e - 4];
  vz1e[nk1e] = vz1e[nk1e - 2] + vz1e[nk1e - 2] - vz1e[nk1e - 4];
  {  // extra_2_west_r4(x5, x4, x3, x2, x1, y5, y4, y3 -> y2, y1)
    float x5 = vz1e[5], x4 = vz1e[4], x3 = vz1e[3], x2 = vz1e[2], x1 = vz1e[1], y5 = vf1e[5], y4 = vf1e[4], y3 = vf1e[3];
    float A = x4 - x5, B = x3 - x4, C = x2 - x3, D = x1 - x2, ALF = y4 - y5, BET = y3 - y4;
    if ((A == 0.f) || (B == 0.f) || (C == 0.f)) { vf1e[2] = y3; vf1e[1] = y3; }
    else {
      float y2 = C * (2 * BET / B - ALF / A) + y3;
      vf1e[2] = y2;
      vf1e[1] = y2 + y2 * D / C + BET * D / B - ALF * D / A - y3 * D / C;
    }
  }
  {  // extra_2_east_r4(x1..x5, y1, y2, y3 -> y4, y5)
    float x1 = vz1e[nk1e - 4], x2 = vz1e[nk1e - 3], x3 = vz1e[nk1e - 2], x4 = vz1e[nk1e - 1], x5 = vz1e[nk1e];
    float y1 = vf1e[nk1e - 4], y2 = vf1e[nk1e - 3], y3 = vf1e[nk1e - 2];
    float A = x2 - x1, B = x3 - x2, C = x4 - x3, D = x5 - x4, ALF = y2 - y1, BET = y3 - y2;
    if ((A == 0.f) || (B == 0.f) || (C == 0.f)) { vf1e[nk1e - 1] = y3; vf1e[nk1e] = y3; }
    else {
      float y4 = C * (2 * BET / B - ALF / A) + y3;
      vf1e[nk1e - 1] = y4;
      vf1e[nk1e] = y4 + y4 * D / C + BET * D / B - ALF * D / A - y3 * D / C;
    }
  }
  {  // slopes_1d
    const int nz = nk1e;
    auto x = [&](int k) { return vz1e[k]; };
    auto y = [&](int k) { return vf1e[k]; };
    for (int k = 3; k <= nz - 2; k++) {
      float m1 = (y(k - 1) - y(k - 2)) / (x(k - 1) - x(k - 2));
      float m2 = (y(k) - y(k - 1)) / (x(k) - x(k - 1));
      float m3 = (y(k + 1) - y(k)) / (x(k + 1) - x(k));
      float m4 = (y(k + 2) - y(k + 1)) / (x(k + 2) - x(k + 1));
      if ((m1 == m2) && (m3 == m4)) vslp[k] = 0.5f * (m2 + m3);
      else vslp[k] = (std::fabs(m4 - m3) * m2 + std::fabs(m2 - m1) * m3) / (std::fabs(m4 - m3) + std::fabs(m2 - m1));
    }
    float m2 = (y(2) - y(1)) / (x(2) - x(1)), m3 = (y(3) - y(2)) / (x(3) - x(2));
    vslp[2] = 0.5f * (m2 + m3);
    m2 = (y(nz - 1) - y(nz - 2)) / (x(nz - 1) - x(nz - 2));
    m3 = (y(nz) - y(nz - 1)) / (x(nz) - x(nz - 1));
    vslp[nz - 1] = 0.5f * (m2 + m3);
    vslp[1] = (y(2) - y(1)) / (x(2) - x(1));
    vslp[nz] = (y(nz) - y(nz - 1)) / (x(nz) - x(nz - 1));
  }
  for (int jk2 = 1; jk2 <= nk2; jk2++) {
    bool l_do_interp = true, lfnd = false;
    int jk1 = k1_1, jp1 = 0, jp2 = 0;
    if ((vz2(jk2) < vz1(k1_1)) && l_sp) { vf2_[jk2 - 1] = (float)vf1(k1_1); l_do_interp = false; lfnd = true; }
    while ((!lfnd) && l_do_interp) {
      if (jk1 == k1_n) break;
      if ((vz2(jk2) > vz1(jk1)) && (vz2(jk2) <= vz1(jk1 + 1))) {
        const int jk1e = jk1 + 2 - k1_1 + 1;
        jp1 = jk1e; jp2 = jk1e + 1; lfnd = true;
      } else {
        if (vz2(jk2) == vz1(jk1)) { vf2_[jk2 - 1] = (float)vf1(jk1); l_do_interp = false; break; }
        else if (vz2(jk2) == vz1(jk1 + 1)) { vf2_[jk2 - 1] = (float)vf1(jk1 + 1); l_do_interp = false; break; }
      }
      jk1 = jk1 + 1;
    }
    if (!lfnd) {
      l_do_interp = false;
      if ((jk1 == k1_n) && l_bp) vf2_[jk2 - 1] = (float)vf1(k1_n);
    }
    if (l_do_interp) {
      double a0 = vf1e[jp1];
      double a1 = vslp[jp1];
      double dz = 1.f / (vz1e[jp2] - vz1e[jp1]);                       // REAL(4) quotient stored in a REAL(8)
      double dz2 = dz * dz;
      double a2 = ((double)(3 * (vf1e[jp2] - vf1e[jp1])) * dz - (double)(2 * vslp[jp1]) - (double)vslp[jp2]) * dz;
      double a3 = (double)(vslp[jp1] + vslp[jp2] - 2 * (vf1e[jp2] - vf1e[jp1]) / (vz1e[jp2] - vz1e[jp1])) * dz2;
      dz = vz2(jk2) - (double)vz1e[jp1];
      dz2 = dz * dz;
      vf2_[jk2 - 1] = (float)(a0 + a1 * dz + a2 * dz2 + a3 * dz2 * dz);
    }
  }
  return 0;
}

// the generic-coordinate branch of INTERP_3D (src/mod_interp.f90:374-438): per target column, AKIMA_1D_1D between the
// source depths carried to the target grid and the target depths, sigma columns flipped, persistence below the last
// level the source reaches.  The reference flips its work arrays in place (and leaves them flipped for masked
// columns); here the inputs are left untouched and only data3d_trg is produced.  Returns the worst column status.
int interp3d_sigma(int ni, int nj, int nk_src, const float* depth_src, const float* data3d_tmp, int nk_trg, const float* depth_trg,
                   const int32_t* mask_trg, bool lmout, bool src_sigma, bool trg_sigma, float* data3d_trg) {
  const int64_t n = (int64_t)ni * nj;
  int worst = 0;
  std::vector<double> zs(nk_src), fs(nk_src), zt(nk_trg);
  std::vector<float> t(nk_trg);
  for (int64_t p = 0; p < n; p++) {
    for (int k = 0; k < nk_src; k++) {
      const int kk = src_sigma ? nk_src - 1 - k : k;   // FLIP_UD
      zs[k] = (double)depth_src[p + (int64_t)kk * n];
      fs[k] = (double)data3d_tmp[p + (int64_t)kk * n];
    }
    for (int k = 0; k < nk_trg; k++) zt[k] = (double)depth_trg[p + (int64_t)(trg_sigma ? nk_trg - 1 - k : k) * n];
    float zmax_src = (float)zs[0], zmax_trg = (float)zt[0];
    for (int k = 1; k < nk_src; k++) if ((float)zs[k] > zmax_src) zmax_src = (float)zs[k];
    for (int k = 1; k < nk_trg; k++) if ((float)zt[k] > zmax_trg) zmax_trg = (float)zt[k];
    int jklast;
    if (zmax_trg > zmax_src) {
      jklast = 1;
      while (jklast < nk_trg) {
        if ((float)zt[jklast] > zmax_src) break;   // depth_trg(ji,jj,jklast+1)
        jklast = jklast + 1;
      }
    } else jklast = nk_trg;
    if ((!lmout) || (mask_trg[p] == 1)) {
      int nlev;
      if (lmout) {
        nlev = 1;
        for (int jk = 1; jk <= nk_trg; jk++) if (mask_trg[p + (int64_t)(jk - 1) * n] == 1) nlev = nlev + 1;
        nlev = nlev - 1;
      } else nlev = jklast;
      for (int k = 0; k < nk_trg; k++) t[k] = data3d_trg[p + (int64_t)k * n];
      int st = akima_1d_1d(nk_src, zs.data(), fs.data(), nlev, zt.data(), t.data(), false, 0.0, false, false);
      if (st < worst) worst = st;
      if ((jklast > 0) && (jklast < nk_trg))
        for (int jk = jklast + 1; jk <= nk_trg; jk++) t[jk - 1] = t[jklast - 1];
      if (trg_sigma) std::reverse(t.begin(), t.end());
      for (int k = 0; k < nk_trg; k++) data3d_trg[p + (int64_t)k * n] = t[k];
    }
  }
  return worst;
}

}  // namespace orc

extern "C" {

// AKIMA_1D_1D over ncol columns; element (column c, level k) of an array at base[c*cs + k*ks].  status (ncol) may be NULL.
int orc_akima_1d_1d(int64_t ncol, int nk1, const double* vz1, int64_t z1cs, int64_t z1ks, const double* vf1, int64_t f1cs, int64_t f1ks,
                    int nk2, const double* vz2, int64_t z2cs, int64_t z2ks, float* vf2, int64_t f2cs, int64_t f2ks, int has_rmask,
                    double rmask_val, int l_surf_persist, int l_botm_persist, int32_t* status) {
  int worst = 0;
  std::vector<double> a(nk1), b(nk1), c(nk2);
  std::vector<float> o(nk2);
  for (int64_t q = 0; q < ncol; q++) {
    for (int k = 0; k < nk1; k++) { a[k] = vz1[q * z1cs + k * z1ks]; b[k] = vf1[q * f1cs + k * f1ks]; }
    for (int k = 0; k < nk2; k++) c[k] = vz2[q * z2cs + k * z2ks];
    int st = orc::akima_1d_1d(nk1, a.data(), b.data(), nk2, c.data(), o.data(), has_rmask != 0, rmask_val, l_surf_persist != 0, l_botm_persist != 0);
    for (int k = 0; k < nk2; k++) vf2[q * f2cs + k * f2ks] = o[k];
    if (status) status[q] = st;
    if (st < worst) worst = st;
  }
  return worst;
}

int orc_interp3d_sigma(int ni, int nj, int nk_src, const float* depth_src, const float* data3d_tmp, int nk_trg, const float* depth_trg,
                       const int32_t* mask_trg, int lmout, int src_sigma, int trg_sigma, float* data3d_trg) {
  return orc::interp3d_sigma(ni, nj, nk_src, depth_src, data3d_tmp, nk_trg, depth_trg, mask_trg, lmout != 0, src_sigma != 0, trg_sigma != 0, data3d_trg);
}

void orc_akima_1d_3d(int nx, int ny, int nk1, const float* vz1, const float* xf1, int nk2, const float* vz2, float* xf2, float rfill_val) {
  orc::akima_1d_3d(nx, ny, nk1, vz1, xf1, nk2, vz2, xf2, rfill_val);
}

void orc_lbc_lnk(int nperio, int jpi, int jpj, int nslab, float* X, int cd_type, double psgn, int has_pval, double pval) {
  for (int s = 0; s < nslab; s++) orc::lbc_lnk_2d_r4(nperio, X + (size_t)s * jpi * jpj, jpi, jpj, (char)cd_type, psgn, has_pval != 0, pval);
}

// post-interpolation block of INTERP_3D (src/mod_interp.f90:447-511) without the optional SMOOTHER calls:
// bottom persistence (ixtrpl_bot == 1), bounds, lbc_lnk on ORCA targets, target mask
void orc_interp3d_post(int ni, int nj, int nk, float* X, const int32_t* mask_trg, int ixtrpl_bot, float vmin, float vmax, float rmiss_val,
                       int i_orca_trg, int c_orca_trg, int lmout) {
  const int64_t n = (int64_t)ni * nj;
  if (ixtrpl_bot == 1) {
    for (int64_t p = 0; p < n; p++) {
      if (mask_trg[p] == 1) {
        int jk_bot = 0;  // FINDLOC(mask_trg(ji,jj,:), 0, DIM=1): first bedrock point (0 when none)
        for (int k = 1; k <= nk; k++) if (mask_trg[p + (int64_t)(k - 1) * n] == 0) { jk_bot = k; break; }
        if (jk_bot > 1) {
          float zwet = (float)(double)X[p + (int64_t)(jk_bot - 2) * n];  // zwet is REAL(8): exact round trip
          for (int k = jk_bot; k <= nk; k++) X[p + (int64_t)(k - 1) * n] = zwet;
        }
      }
    }
  }
  for (int k = 1; k <= nk; k++) {
    float* Xk = X + (int64_t)(k - 1) * n;
    for (int64_t p = 0; p < n; p++) if ((Xk[p] > vmax) && (Xk[p] != rmiss_val)) Xk[p] = vmax;
    for (int64_t p = 0; p < n; p++) if ((Xk[p] < vmin) && (Xk[p] != rmiss_val)) Xk[p] = vmin;
    if (i_orca_trg > 0) orc::lbc_lnk_2d_r4(i_orca_trg, Xk, ni, nj, (char)c_orca_trg, 1.0, false, 0.0);
    if (lmout) {
      const int32_t* mk = mask_trg + (int64_t)(k - 1) * n;
      for (int64_t p = 0; p < n; p++) Xk[p] = Xk[p] * (float)mk[p] + (1.f - (float)mk[p]) * rmiss_val;
    }
  }
}

void orc_angle2(int nperio, int nx, int ny, const double* plamt, const double* pphit, const double* plamu, const double* pphiu,
                const double* plamv, const double* pphiv, const double* plamf, const double* pphif, double* gcost, double* gsint,
                double* gcosu, double* gsinu, double* gcosv, double* gsinv, double* gcosf, double* gsinf) {
  orc::angle2(nperio, nx, ny, plamt, pphit, plamu, pphiu, plamv, pphiv, plamf, pphif, gcost, gsint, gcosu, gsinu, gcosv, gsinv, gcosf, gsinf);
}

// extrp_hl (src/mod_interp.f90:521-542): persistence of the last valid row towards the top / bottom of a regular target
void orc_extrp_hl(int ni, int nj, int nslab, float* X, int jj_ex_top, int jj_ex_btm, int nlat_icr_trg) {
  const int icr = nlat_icr_trg;
  for (int s = 0; s < nslab; s++) {
    orc::A2<float> X2d(X + (size_t)s * ni * nj, ni, nj);
    if (jj_ex_top > 0) {
      const int jend = (icr + 1) / 2 * nj + (1 - icr) / 2;
      for (int jj = jj_ex_top; (icr > 0) ? (jj <= jend) : (jj >= jend); jj += icr)
        for (int i = 1; i <= ni; i++) X2d(i, jj) = X2d(i, jj_ex_top - icr);
    }
    if (jj_ex_btm > 0) {
      const int jbeg = (icr + 1) / 2 + (1 - icr) / 2 * nj;
      for (int jj = jbeg; (icr > 0) ? (jj <= jj_ex_btm) : (jj >= jj_ex_btm); jj += icr)
        for (int i = 1; i <= ni; i++) X2d(i, jj) = X2d(i, jj_ex_btm + icr);
    }
  }
}

}  // extern "C"
