// vert.cu -- the vertical stage of INTERP_3D on device: a 3D job keeps data3d_tmp in HBM between
// the per-level horizontal interpolation and the vertical one instead of going back to the host.
//
// GPU counterparts of
//   AKIMA_1D_3D, slopes_3d, extra_2_top_3d, extra_2_bottom_3d   src/mod_akima_1d.f90:243-389, 461-515, 521-606
//   lbc_lnk_2d_r4 / lbc_lnk_2d_r8 / lbc_nfd_2d                   src/mod_nemotools.f90:65-417
//   the post-interpolation block of INTERP_3D                    src/mod_interp.f90:447-511
// REAL(4) arithmetic in the reference's order, FMA contraction off -> bit-identical results.
//
// AKIMA_1D_3D materialises the n+4 extended cube and its slopes (2 x (nk1+4) planes written and read back) and then
// evaluates each target level.  The level search depends on the two depth vectors only, so it is done once on the host
// (LevelPlan); on the device one thread per (point, target level) rebuilds what the polynomial needs -- six source
// values around the bracketing interval, extrapolated on the fly where they fall in the 2-point frames -- and the
// two slopes from the five divided differences they share.  HBM traffic: the source cube once (planes are re-read
// from L2 by the neighbouring target levels) + the target cube once.
#include <algorithm>

#include "geom.cuh"

#define LV_FILL 0
#define LV_COPY 1
#define LV_INTERP 2

struct LevelPlan {  // per target level
  int mode;         // LV_*
  int k;            // LV_COPY: source level (1-based); LV_INTERP: jp1 in the n+4 extended numbering
};

// value of the extended cube xf1e(:,:,ke) at point p (ke = 1..nk1+4), :266-289
__device__ __forceinline__ float ak1_fe(const float* __restrict__ xf1, size_t n, size_t p, int nk1, const float* __restrict__ vze,
                                        int ke) {
  if (ke >= 3 && ke <= nk1 + 2) return xf1[(size_t)(ke - 3) * n + p];
  const int nk1e = nk1 + 4;
  if (ke <= 2) {  // extra_2_bottom_3d(vz1e(5), vz1e(4), vz1e(3), vz1e(2), vz1e(1), xf1e(5), xf1e(4), xf1e(3) -> xf1e(2), xf1e(1))
    const float x5 = vze[5], x4 = vze[4], x3 = vze[3], x2 = vze[2], x1 = vze[1];
    const float A = __fsub_rn(x4, x5), B = __fsub_rn(x3, x4), C = __fsub_rn(x2, x3), D = __fsub_rn(x1, x2);
    const float f5 = xf1[2 * n + p], f4 = xf1[n + p], f3 = xf1[p];
    if ((A == 0.f) || (B == 0.f) || (C == 0.f)) return f3;
    const float ALF = __fsub_rn(f4, f5), BET = __fsub_rn(f3, f4);
    const float v2 = __fadd_rn(__fmul_rn(C, __fsub_rn(__fdiv_rn(__fmul_rn(2.f, BET), B), __fdiv_rn(ALF, A))), f3);
    if (ke == 2) return v2;
    float r = __fadd_rn(v2, __fdiv_rn(__fmul_rn(v2, D), C));
    r = __fadd_rn(r, __fdiv_rn(__fmul_rn(BET, D), B));
    r = __fsub_rn(r, __fdiv_rn(__fmul_rn(ALF, D), A));
    r = __fsub_rn(r, __fdiv_rn(__fmul_rn(f3, D), C));
    return r;
  }
  // extra_2_top_3d(vz1e(nk1e-4..nk1e), xf1e(nk1e-4), xf1e(nk1e-3), xf1e(nk1e-2) -> xf1e(nk1e-1), xf1e(nk1e))
  const float x1 = vze[nk1e - 4], x2 = vze[nk1e - 3], x3 = vze[nk1e - 2], x4 = vze[nk1e - 1], x5 = vze[nk1e];
  const float A = __fsub_rn(x2, x1), B = __fsub_rn(x3, x2), C = __fsub_rn(x4, x3), D = __fsub_rn(x5, x4);
  const float f1 = xf1[(size_t)(nk1 - 3) * n + p], f2 = xf1[(size_t)(nk1 - 2) * n + p], f3 = xf1[(size_t)(nk1 - 1) * n + p];
  if ((A == 0.f) || (B == 0.f) || (C == 0.f)) return f3;
  const float ALF = __fsub_rn(f2, f1), BET = __fsub_rn(f3, f2);
  const float v4 = __fadd_rn(__fmul_rn(C, __fsub_rn(__fdiv_rn(__fmul_rn(2.f, BET), B), __fdiv_rn(ALF, A))), f3);
  if (ke == nk1e - 1) return v4;
  float r = __fadd_rn(v4, __fdiv_rn(__fmul_rn(v4, D), C));
  r = __fadd_rn(r, __fdiv_rn(__fmul_rn(BET, D), B));
  r = __fsub_rn(r, __fdiv_rn(__fmul_rn(ALF, D), A));
  r = __fsub_rn(r, __fdiv_rn(__fmul_rn(f3, D), C));
  return r;
}

__device__ __forceinline__ float ak1_slope(float m1, float m2, float m3, float m4) {  // slopes_3d, :478-489
  if ((m1 == m2) && (m3 == m4)) return __fmul_rn(0.5f, __fadd_rn(m2, m3));
  const float w2 = fabsf(__fsub_rn(m4, m3)), w3 = fabsf(__fsub_rn(m2, m1));
  return __fdiv_rn(__fadd_rn(__fmul_rn(w2, m2), __fmul_rn(w3, m3)), __fadd_rn(w2, w3));
}

// vze: vz1e(1..nk1+4) at [1..]; vz2: target depths (0-based array)
__global__ void __launch_bounds__(256)
k_ak1d_eval(const float* __restrict__ xf1, float* __restrict__ xf2, size_t n, int nk1, int nk2, const float* __restrict__ vze,
            const float* __restrict__ vz2, const LevelPlan* __restrict__ plan, float rfill) {
  const int jk2 = blockIdx.y;  // 0-based target level
  const LevelPlan lp = plan[jk2];
  float* out = xf2 + (size_t)jk2 * n;
  if (lp.mode == LV_FILL) {
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) __stcs(out + p, rfill);
    return;
  }
  if (lp.mode == LV_COPY) {
    const float* src = xf1 + (size_t)(lp.k - 1) * n;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) __stcs(out + p, src[p]);
    return;
  }
  const int jp1 = lp.k, jp2 = lp.k + 1;
  float z[6];
#pragma unroll
  for (int q = 0; q < 6; q++) z[q] = vze[jp1 - 2 + q];
  float rz[5];
#pragma unroll
  for (int q = 0; q < 5; q++) rz[q] = __fsub_rn(z[q + 1], z[q]);  // vz(k+1) - vz(k): the divisors of slopes_3d
  const float dv = __fsub_rn(vze[jp2], vze[jp1]);
  const float dz = __fdiv_rn(1.f, dv), dz2 = __fmul_rn(dz, dz);
  const float dzt = __fsub_rn(vz2[jk2], vze[jp1]), dzt2 = __fmul_rn(dzt, dzt);
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
    float f[6];
#pragma unroll
    for (int q = 0; q < 6; q++) f[q] = ak1_fe(xf1, n, p, nk1, vze, jp1 - 2 + q);
    float d[5];
#pragma unroll
    for (int q = 0; q < 5; q++) d[q] = __fdiv_rn(__fsub_rn(f[q + 1], f[q]), rz[q]);
    const float s1 = ak1_slope(d[0], d[1], d[2], d[3]);  // xslp(jp1)
    const float s2 = ak1_slope(d[1], d[2], d[3], d[4]);  // xslp(jp2)
    const float a0 = f[2], a1 = s1, df = __fsub_rn(f[3], f[2]);
    // a2 = ( 3*(xf1e(jp2)-xf1e(jp1))*dz - 2*xslp(jp1) - xslp(jp2) ) * dz
    const float a2 = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(__fmul_rn(3.f, df), dz), __fmul_rn(2.f, s1)), s2), dz);
    // a3 = ( xslp(jp1) + xslp(jp2) - 2*(xf1e(jp2)-xf1e(jp1))/(vz1e(jp2)-vz1e(jp1)) ) * dz2
    const float a3 = __fmul_rn(__fsub_rn(__fadd_rn(s1, s2), __fdiv_rn(__fmul_rn(2.f, df), dv)), dz2);
    // xf2 = a0 + a1*dz + a2*dz2 + a3*dz2*dz
    float r = __fadd_rn(a0, __fmul_rn(a1, dzt));
    r = __fadd_rn(r, __fmul_rn(a2, dzt2));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(a3, dzt2), dzt));
    __stcs(out + p, r);
  }
}

// the level search of AKIMA_1D_3D (:293-333) and the two persistence loops after it (:372-385), on the host
static void ak1d_plan(int nk1, const float* vz1, int nk2, const float* vz2, std::vector<LevelPlan>& plan) {
  plan.assign(nk2, LevelPlan{LV_FILL, 0});
  auto VZ1 = [&](int k) { return vz1[k - 1]; };
  auto VZ2 = [&](int k) { return vz2[k - 1]; };
  for (int jk2 = 1; jk2 <= nk2; jk2++) {
    LevelPlan lp{LV_FILL, 0};
    bool l_do_interp = true, lfnd = false;
    int jk1 = 1;
    if (VZ2(jk2) < VZ1(1)) { lp = LevelPlan{LV_COPY, 1}; l_do_interp = false; }
    while ((!lfnd) && l_do_interp) {
      if (jk1 == nk1) break;
      if ((VZ2(jk2) > VZ1(jk1)) && (VZ2(jk2) <= VZ1(jk1 + 1))) { lp = LevelPlan{LV_INTERP, jk1 + 2}; lfnd = true; }
      else {
        if (VZ2(jk2) == VZ1(jk1)) { lp = LevelPlan{LV_COPY, jk1}; l_do_interp = false; break; }
        else if (VZ2(jk2) == VZ1(jk1 + 1)) { lp = LevelPlan{LV_COPY, jk1 + 1}; l_do_interp = false; break; }
      }
      jk1 = jk1 + 1;
    }
    if ((jk1 == nk1) && (!lfnd)) { lp = LevelPlan{LV_COPY, nk1}; l_do_interp = false; }
    if (!lfnd) { lp = LevelPlan{LV_FILL, 0}; l_do_interp = false; }   // (:329-332) overrides every copy above
    plan[jk2 - 1] = lp;
  }
  int jk2 = 1;
  while ((VZ2(jk2) < VZ1(1)) && (jk2 < nk2)) { plan[jk2 - 1] = LevelPlan{LV_COPY, 1}; jk2 = jk2 + 1; }
  jk2 = nk2;
  while ((jk2 > 0) && (VZ2(jk2) > VZ1(nk1))) { plan[jk2 - 1] = LevelPlan{LV_COPY, nk1}; jk2 = jk2 - 1; }  // defined stop at 0
}

int akima_1d_3d_impl(sosie_ctx* ctx, int nx, int ny, int nk1, const float* vz1_host, const float* dxf1, int nk2, const float* vz2_host,
                     float* dxf2, float rfill_val) {
  if (nx < 1 || ny < 1 || nk1 < 3 || nk2 < 1 || !vz1_host || !vz2_host || !dxf1 || !dxf2) {
    ctx->err = "sosie_akima_1d_3d: bad arguments (at least 3 source levels are needed)"; return SOSIE_ERR_ARG;
  }
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)nx * ny;
  const int nk1e = nk1 + 4;
  std::vector<float> vze(nk1e + 1, 0.f);
  for (int k = 1; k <= nk1; k++) vze[k + 2] = vz1_host[k - 1];
  vze[2] = vze[4] - (vze[5] - vze[3]);                              // (:268-273), REAL(4)
  vze[1] = vze[3] - (vze[5] - vze[3]);
  vze[nk1e - 1] = vze[nk1e - 3] + vze[nk1e - 2] - vze[nk1e - 4];
  vze[nk1e] = vze[nk1e - 2] + vze[nk1e - 2] - vze[nk1e - 4];
  std::vector<LevelPlan> plan;
  ak1d_plan(nk1, vz1_host, nk2, vz2_host, plan);
  CK(ctx->v_vze.alloc(nk1e + 1)); CK(ctx->v_vz2.alloc(nk2)); CK(ctx->v_plan.alloc((size_t)nk2 * 2));
  CK(cudaMemcpyAsync(ctx->v_vze.p, vze.data(), sizeof(float) * (nk1e + 1), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->v_vz2.p, vz2_host, sizeof(float) * nk2, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->v_plan.p, plan.data(), sizeof(LevelPlan) * nk2, cudaMemcpyHostToDevice, st));
  dim3 grid(grid_for(ctx, n, 256, 4), nk2);
  k_ak1d_eval<<<grid, 256, 0, st>>>(dxf1, dxf2, n, nk1, nk2, ctx->v_vze.p, ctx->v_vz2.p, reinterpret_cast<const LevelPlan*>(ctx->v_plan.p),
                                    rfill_val);
  KLAUNCH_CHECK();
  CK(cudaStreamSynchronize(st));  // the pageable host vectors above must outlive the copies
  return SOSIE_OK;
}

// ---------------------------------------------------------------------------------------------------
// AKIMA_1D_1D (src/mod_akima_1d.f90:26-232): the per-column variant, used by the generic-coordinate (sigma) branch of
// INTERP_3D (src/mod_interp.f90:374-438) and by interp_to_hydro_section (:628-629).  One thread per column.  As for the
// cube variant nothing is materialised: the four frame depths / values of the REAL(4) extended arrays are computed once
// per column, the six values and five divided differences a target level needs are rebuilt at its interval.
// The scan for the interval is the reference's (first match from k1_1 upwards, no monotonicity assumed on vz2).
template <class T>
struct ColView {      // element k (1-based, in the possibly flipped frame) of one column
  const T* p;
  long long ks;
  int n;
  int flip;
  __device__ __forceinline__ double at(int k) const { const int kk = flip ? n + 1 - k : k; return (double)p[(long long)(kk - 1) * ks]; }
};

__device__ __forceinline__ void ak11_extra(float xa, float xb, float xc, float xd, float xe, float ya, float yb, float yc, float* y4, float* y5) {
  // extra_2_east_r4(x1..x5, y1..y3) and, with the arguments reversed by the caller, extra_2_west_r4 (src/mod_manip.f90:571-587, 634-649)
  const float A = xb - xa, B = xc - xb, C = xd - xc, D = xe - xd, ALF = yb - ya, BET = yc - yb;
  if ((A == 0.f) || (B == 0.f) || (C == 0.f)) { *y4 = yc; *y5 = yc; return; }
  const float v4 = C * (2.f * BET / B - ALF / A) + yc;
  *y4 = v4;
  *y5 = v4 + v4 * D / C + BET * D / B - ALF * D / A - yc * D / C;
}

// returns the column status (0 done, 1 all missing, -1 / -2 undefined in the reference: see include/sosie_b200.h)
template <class T>
__device__ int ak11_column(const ColView<T>& z1, const ColView<T>& f1, int nk1, const ColView<T>& z2, int nk2, float* out, long long oks,
                           bool has_rmv, double rmv, bool l_sp, bool l_bp) {
  for (int k = 0; k < nk2; k++) out[(long long)k * oks] = -6666.0f;
  int nz1 = nk1, k1_1 = 1, k1_n = nk1;
  if (has_rmv) {
    int nmissv = 0, first0 = 0;
    for (int k = 1; k <= nk1; k++) { const bool m = (f1.at(k) == rmv); nmissv += m; if (!m && !first0) first0 = k; }
    if (!(nmissv < nk1)) return 1;
    if (nmissv > 0) {
      k1_1 = first0;
      int f = 0;
      for (int k = k1_1; k <= nk1; k++) if (f1.at(k) == rmv) { f = k - k1_1 + 1; break; }
      k1_n = f + k1_1 - 2;
      nz1 = nk1 - nmissv;
      if (k1_n - k1_1 + 1 != nz1) return -1;
    }
  }
  if (nz1 < 3) return -2;
  const int nk1e = nz1 + 4;
  auto zi = [&](int i) { return (float)z1.at(k1_1 + i - 3); };   // interior of vz1e / vf1e, i = 3 .. nz1+2
  auto fi = [&](int i) { return (float)f1.at(k1_1 + i - 3); };
  float zb2, zb1, zt1, zt2, fb2, fb1, ft1, ft2;                    // vz1e(2), vz1e(1), vz1e(nk1e-1), vz1e(nk1e); same for vf1e
  {
    const float z3 = zi(3), z4 = zi(4), z5 = zi(5);
    zb2 = z4 - (z5 - z3);
    zb1 = z3 - (z5 - z3);
    const float za = zi(nk1e - 4), zb = zi(nk1e - 3), zc = zi(nk1e - 2);
    zt1 = zb + zc - za;
    zt2 = zc + zc - za;
    ak11_extra(z5, z4, z3, zb2, zb1, fi(5), fi(4), fi(3), &fb2, &fb1);
    ak11_extra(za, zb, zc, zt1, zt2, fi(nk1e - 4), fi(nk1e - 3), fi(nk1e - 2), &ft1, &ft2);
  }
  auto ze = [&](int i) { return i <= 2 ? (i == 2 ? zb2 : zb1) : (i >= nk1e - 1 ? (i == nk1e ? zt2 : zt1) : zi(i)); };
  auto fe = [&](int i) { return i <= 2 ? (i == 2 ? fb2 : fb1) : (i >= nk1e - 1 ? (i == nk1e ? ft2 : ft1) : fi(i)); };
  const double z1_first = z1.at(k1_1);
  for (int jk2 = 1; jk2 <= nk2; jk2++) {
    const double zt = z2.at(jk2);
    float* o = out + (long long)(jk2 - 1) * oks;
    bool l_do_interp = true, lfnd = false;
    int jk1 = k1_1, jp1 = 0;
    if ((zt < z1_first) && l_sp) { *o = (float)f1.at(k1_1); l_do_interp = false; lfnd = true; }
    if (!lfnd) {
      double za = z1_first;
      while (true) {
        if (jk1 == k1_n) break;
        const double zb = z1.at(jk1 + 1);
        if ((zt > za) && (zt <= zb)) { jp1 = jk1 + 2 - k1_1 + 1; lfnd = true; break; }
        if (zt == za) { *o = (float)f1.at(jk1); l_do_interp = false; break; }
        if (zt == zb) { *o = (float)f1.at(jk1 + 1); l_do_interp = false; break; }
        jk1 = jk1 + 1;
        za = zb;
      }
    }
    if (!lfnd) {
      l_do_interp = false;
      if ((jk1 == k1_n) && l_bp) *o = (float)f1.at(k1_n);
    }
    if (!l_do_interp) continue;
    const int jp2 = jp1 + 1;
    float z[6], f[6], d[5];
#pragma unroll
    for (int q = 0; q < 6; q++) { z[q] = ze(jp1 - 2 + q); f[q] = fe(jp1 - 2 + q); }
#pragma unroll
    for (int q = 0; q < 5; q++) d[q] = (f[q + 1] - f[q]) / (z[q + 1] - z[q]);
    const float s1 = ak1_slope(d[0], d[1], d[2], d[3]);   // vslp(jp1): the generic branch of slopes_1d (3 <= jp1, jp2 <= nk1e-2)
    const float s2 = ak1_slope(d[1], d[2], d[3], d[4]);   // vslp(jp2)
    const float df = f[3] - f[2], dv = z[3] - z[2];
    const double a0 = (double)f[2], a1 = (double)s1;
    double dz = (double)(1.f / dv), dz2 = dz * dz;
    const double a2 = ((double)(3.f * df) * dz - (double)(2.f * s1) - (double)s2) * dz;
    const double a3 = (double)(s1 + s2 - 2.f * df / dv) * dz2;
    dz = zt - (double)z[2];
    dz2 = dz * dz;
    *o = (float)(a0 + a1 * dz + a2 * dz2 + a3 * dz2 * dz);
    (void)jp2;
  }
  return 0;
}

template <class T>
__global__ void __launch_bounds__(128)
k_ak11_cols(long long ncol, int nk1, const T* vz1, long long z1cs, long long z1ks, const T* vf1, long long f1cs, long long f1ks, int nk2,
            const T* vz2, long long z2cs, long long z2ks, float* vf2, long long f2cs, long long f2ks, int has_rmv, double rmv, int l_sp,
            int l_bp, int* worst, int32_t* status) {
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncol; c += (long long)gridDim.x * blockDim.x) {
    ColView<T> z1{vz1 + c * z1cs, z1ks, nk1, 0}, f1{vf1 + c * f1cs, f1ks, nk1, 0}, z2{vz2 + c * z2cs, z2ks, nk2, 0};
    const int st = ak11_column(z1, f1, nk1, z2, nk2, vf2 + c * f2cs, f2ks, has_rmv != 0, rmv, l_sp != 0, l_bp != 0);
    if (status) status[c] = st;
    if (st < 0) atomicMin(worst, st);
  }
}

// the generic-coordinate branch of INTERP_3D (src/mod_interp.f90:374-438), one thread per target column; the
// reference's in-place flips of its work arrays become index reversals, the inputs are not modified
__global__ void __launch_bounds__(128)
k_i3d_sigma(long long n, int nk_src, const float* __restrict__ depth_src, const float* __restrict__ data3d_tmp, int nk_trg,
            const float* __restrict__ depth_trg, const int32_t* __restrict__ mask_trg, int lmout, int src_sigma, int trg_sigma,
            float* data3d_trg, int* worst) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    ColView<float> zs{depth_src + p, n, nk_src, src_sigma}, fs{data3d_tmp + p, n, nk_src, src_sigma}, zt{depth_trg + p, n, nk_trg, trg_sigma};
    float zmax_src = depth_src[p], zmax_trg = depth_trg[p];
    for (int k = 1; k < nk_src; k++) zmax_src = fmaxf(zmax_src, depth_src[p + (long long)k * n]);
    for (int k = 1; k < nk_trg; k++) zmax_trg = fmaxf(zmax_trg, depth_trg[p + (long long)k * n]);
    int jklast;
    if (zmax_trg > zmax_src) {
      jklast = 1;
      while (jklast < nk_trg) {
        if ((float)zt.at(jklast + 1) > zmax_src) break;
        jklast = jklast + 1;
      }
    } else jklast = nk_trg;
    if (lmout && mask_trg[p] != 1) continue;
    int nlev = jklast;
    if (lmout) {
      nlev = 0;
      for (int jk = 0; jk < nk_trg; jk++) nlev += (mask_trg[p + (long long)jk * n] == 1);
    }
    float* col = data3d_trg + p;
    zt.n = nk_trg;   // the flip is of the WHOLE column; only the first nlev levels of the flipped column are targets
    const int st = ak11_column(zs, fs, nk_src, zt, nlev, col, n, false, 0.0, false, false);
    if (st < 0) atomicMin(worst, st);
    if ((jklast > 0) && (jklast < nk_trg)) {
      const float v = col[(long long)(jklast - 1) * n];
      for (int jk = jklast + 1; jk <= nk_trg; jk++) col[(long long)(jk - 1) * n] = v;
    }
    if (trg_sigma) {
      for (int a = 0, b = nk_trg - 1; a < b; a++, b--) {
        const float t = col[(long long)a * n];
        col[(long long)a * n] = col[(long long)b * n];
        col[(long long)b * n] = t;
      }
    }
  }
}

static int ak11_status(sosie_ctx* ctx, int* dworst, const char* who) {
  int hw = 0;
  if (int rc = fetch_small(ctx, dworst, &hw, sizeof(int))) return rc;
  if (hw == -1) {
    ctx->err = std::string(who) + ": AKIMA_1D_1D: missing values are not confined to the ends of a column (array shapes mismatch in the reference)";
    return SOSIE_ERR_REF_STOP;
  }
  if (hw == -2) { ctx->err = std::string(who) + ": AKIMA_1D_1D: a column has fewer than 3 valid source points (the reference reads unset elements)"; return SOSIE_ERR_REF_STOP; }
  return SOSIE_OK;
}

template <class T>
static int akima_1d_1d_launch(sosie_ctx* ctx, long long ncol, int nk1, const T* vz1, long long z1cs, long long z1ks, const T* vf1, long long f1cs,
                              long long f1ks, int nk2, const T* vz2, long long z2cs, long long z2ks, float* vf2, long long f2cs, long long f2ks,
                              int has_rmask, double rmask_val, int l_sp, int l_bp, int32_t* dstatus) {
  cudaStream_t st = ctx->stream;
  CK(ctx->d_iscratch.alloc(SOSIE_ISCRATCH));
  int* dworst = ctx->d_iscratch.p;
  CK(cudaMemsetAsync(dworst, 0, sizeof(int), st));
  k_ak11_cols<T><<<grid_for(ctx, (size_t)ncol, 128), 128, 0, st>>>(ncol, nk1, vz1, z1cs, z1ks, vf1, f1cs, f1ks, nk2, vz2, z2cs, z2ks, vf2, f2cs,
                                                                    f2ks, has_rmask, rmask_val, l_sp, l_bp, dworst, dstatus);
  KLAUNCH_CHECK();
  return ak11_status(ctx, dworst, "sosie_akima_1d_1d");
}

int akima_1d_1d_impl(sosie_ctx* ctx, long long ncol, int nk1, const double* vz1, long long z1cs, long long z1ks, const double* vf1, long long f1cs,
                     long long f1ks, int nk2, const double* vz2, long long z2cs, long long z2ks, float* vf2, long long f2cs, long long f2ks,
                     int has_rmask, double rmask_val, int l_sp, int l_bp, int32_t* dstatus) {
  if (ncol < 1 || nk1 < 1 || nk2 < 1 || !vz1 || !vf1 || !vz2 || !vf2) { ctx->err = "sosie_akima_1d_1d: bad arguments"; return SOSIE_ERR_ARG; }
  return akima_1d_1d_launch<double>(ctx, ncol, nk1, vz1, z1cs, z1ks, vf1, f1cs, f1ks, nk2, vz2, z2cs, z2ks, vf2, f2cs, f2ks, has_rmask, rmask_val,
                                    l_sp, l_bp, dstatus);
}

int interp3d_sigma_impl(sosie_ctx* ctx, int ni, int nj, int nk_src, const float* ddepth_src, const float* ddata3d_tmp, int nk_trg,
                        const float* ddepth_trg, const int32_t* dmask_trg, int lmout, int src_sigma, int trg_sigma, float* ddata3d_trg) {
  if (ni < 1 || nj < 1 || nk_src < 1 || nk_trg < 1 || !ddepth_src || !ddata3d_tmp || !ddepth_trg || !ddata3d_trg || (lmout && !dmask_trg)) {
    ctx->err = "sosie_interp3d_sigma: bad arguments"; return SOSIE_ERR_ARG;
  }
  cudaStream_t st = ctx->stream;
  const long long n = (long long)ni * nj;
  CK(ctx->d_iscratch.alloc(SOSIE_ISCRATCH));
  int* dworst = ctx->d_iscratch.p;
  CK(cudaMemsetAsync(dworst, 0, sizeof(int), st));
  k_i3d_sigma<<<grid_for(ctx, (size_t)n, 128), 128, 0, st>>>(n, nk_src, ddepth_src, ddata3d_tmp, nk_trg, ddepth_trg, dmask_trg, lmout, src_sigma,
                                                             trg_sigma, ddata3d_trg, dworst);
  KLAUNCH_CHECK();
  return ak11_status(ctx, dworst, "sosie_interp3d_sigma");
}

// ---------------------------------------------------------------------------------------------------
// lbc_lnk (REAL(4) variant: through REAL(8), src/mod_nemotools.f90:154-190): one block per slab executes the
// statements of lbc_lnk_2d_r8 / lbc_nfd_2d in order.  Loops that copy between different rows (or columns) are
// parallel; loops that read and write the SAME row are replayed by one thread in the reference's index order (the
// T-pivot U-point fold reads elements it has just overwritten).
__device__ __forceinline__ float lbc_mul(double psgn, float x) { return (float)__dmul_rn(psgn, (double)x); }
__device__ __forceinline__ double lbc_mul(double psgn, double x) { return __dmul_rn(psgn, x); }

// T = float: lbc_lnk_2d_r4 (through REAL(8)); T = double: lbc_lnk_2d_r8.  slab_stride in elements.
template <class T>
__global__ void __launch_bounds__(1024)
k_lbc_lnk(T* Xb, size_t slab_stride, int jpi, int jpj, int nperio, int cd, double psgn, int has_pval, double pval) {
  T* X = Xb + (size_t)blockIdx.x * slab_stride;
#define P(i, j) X[(size_t)((i)-1) + (size_t)((j)-1) * jpi]
  const T zland = has_pval ? (T)pval : (T)0;
  const int t = threadIdx.x, nt = blockDim.x;
  const bool tuvw = (cd == 'T' || cd == 'U' || cd == 'V' || cd == 'W');
  // ---- East-West
  if (nperio == 1 || nperio == 4 || nperio == 6) {
    for (int j = 1 + t; j <= jpj; j += nt) P(1, j) = P(jpi - 1, j);
    __syncthreads();
    for (int j = 1 + t; j <= jpj; j += nt) P(jpi, j) = P(2, j);
  } else {
    if (tuvw) { for (int j = 1 + t; j <= jpj; j += nt) { P(1, j) = zland; P(jpi, j) = zland; } }
    else if (cd == 'F') { for (int j = 1 + t; j <= jpj; j += nt) P(jpi, j) = zland; }
  }
  __syncthreads();
  // ---- North-South
  if (nperio == 2) {
    if (cd == 'T' || cd == 'U' || cd == 'W') { for (int i = 1 + t; i <= jpi; i += nt) { P(i, 1) = P(i, 3); P(i, jpj) = zland; } }
    else if (cd == 'V' || cd == 'F') { for (int i = 1 + t; i <= jpi; i += nt) { P(i, 1) = lbc_mul(psgn, P(i, 2)); P(i, jpj) = zland; } }
  } else if (nperio >= 3 && nperio <= 6) {
    if (tuvw || cd == 'I') { for (int i = 1 + t; i <= jpi; i += nt) P(i, 1) = zland; }
    __syncthreads();
    const int jg = jpi, jj = jpj, jm1 = jpj - 1;
    if (nperio == 3 || nperio == 4) {  // T-point pivot
      if (cd == 'T' || cd == 'W') {
        for (int ji = 2 + t; ji <= jg; ji += nt) P(ji, jj) = lbc_mul(psgn, P(jg - ji + 2, jj - 2));
        __syncthreads();
        if (t == 0) {
          P(1, jj) = lbc_mul(psgn, P(3, jj - 2));
          for (int ji = jg / 2 + 1; ji <= jg; ji++) P(ji, jj - 1) = lbc_mul(psgn, P(jg - ji + 2, jj - 1));
        }
      } else if (cd == 'U') {
        for (int ji = 1 + t; ji <= jg - 1; ji += nt) P(ji, jj) = lbc_mul(psgn, P(jg - ji + 1, jj - 2));
        __syncthreads();
        if (t == 0) {
          P(1, jj) = lbc_mul(psgn, P(2, jj - 2));
          P(jg, jj) = lbc_mul(psgn, P(jg - 1, jj - 2));
          P(1, jj - 1) = lbc_mul(psgn, P(jg, jj - 1));
          for (int ji = jg / 2; ji <= jg - 1; ji++) P(ji, jm1) = lbc_mul(psgn, P(jg - ji + 1, jm1));
        }
      } else if (cd == 'V') {
        for (int ji = 2 + t; ji <= jg; ji += nt) P(ji, jj - 1) = lbc_mul(psgn, P(jg - ji + 2, jj - 2));  // jl = -1
        __syncthreads();
        for (int ji = 2 + t; ji <= jg; ji += nt) P(ji, jj) = lbc_mul(psgn, P(jg - ji + 2, jj - 3));      // jl = 0
        __syncthreads();
        if (t == 0) P(1, jj) = lbc_mul(psgn, P(3, jj - 3));
      } else if (cd == 'F') {
        for (int ji = 1 + t; ji <= jg - 1; ji += nt) P(ji, jj - 1) = lbc_mul(psgn, P(jg - ji + 1, jj - 2));
        __syncthreads();
        for (int ji = 1 + t; ji <= jg - 1; ji += nt) P(ji, jj) = lbc_mul(psgn, P(jg - ji + 1, jj - 3));
        __syncthreads();
        if (t == 0) {
          P(1, jj) = lbc_mul(psgn, P(2, jj - 3));
          P(jg, jj) = lbc_mul(psgn, P(jg - 1, jj - 3));
          P(jg, jj - 1) = lbc_mul(psgn, P(jg - 1, jj - 2));
          P(1, jj - 1) = lbc_mul(psgn, P(2, jj - 2));
        }
      }
    } else {  // F-point pivot
      if (cd == 'T' || cd == 'W') {
        for (int ji = 1 + t; ji <= jg; ji += nt) P(ji, jj) = lbc_mul(psgn, P(jg - ji + 1, jj - 1));
      } else if (cd == 'U') {
        for (int ji = 1 + t; ji <= jg - 1; ji += nt) P(ji, jj) = lbc_mul(psgn, P(jg - ji, jj - 1));
        __syncthreads();
        if (t == 0) P(jg, jj) = lbc_mul(psgn, P(1, jj - 1));
      } else if (cd == 'V') {
        for (int ji = 1 + t; ji <= jg; ji += nt) P(ji, jj) = lbc_mul(psgn, P(jg - ji + 1, jj - 2));
        __syncthreads();
        if (t == 0) for (int ji = jg / 2 + 1; ji <= jg; ji++) P(ji, jm1) = lbc_mul(psgn, P(jg - ji + 1, jm1));
      } else if (cd == 'F') {
        for (int ji = 1 + t; ji <= jg - 1; ji += nt) P(ji, jj) = lbc_mul(psgn, P(jg - ji, jj - 2));
        __syncthreads();
        if (t == 0) {
          P(jg, jj) = lbc_mul(psgn, P(1, jj - 2));
          for (int ji = jg / 2 + 1; ji <= jg - 1; ji++) P(ji, jm1) = lbc_mul(psgn, P(jg - ji, jm1));
        }
      }
    }
  } else {
    if (tuvw) { for (int i = 1 + t; i <= jpi; i += nt) { P(i, 1) = zland; P(i, jpj) = zland; } }
    else if (cd == 'F') { for (int i = 1 + t; i <= jpi; i += nt) P(i, jpj) = zland; }
  }
#undef P
}

int lbc_lnk_impl(sosie_ctx* ctx, int nperio, int jpi, int jpj, int nslab, float* dX, int cd_type, double psgn, int has_pval, double pval) {
  if (jpi < 4 || jpj < 4 || nslab < 1 || !dX) { ctx->err = "sosie_lbc_lnk: bad arguments"; return SOSIE_ERR_ARG; }
  const bool fold = nperio >= 3 && nperio <= 6;
  if (fold && !(cd_type == 'T' || cd_type == 'W' || cd_type == 'U' || cd_type == 'V' || cd_type == 'F')) {
    ctx->err = "sosie_lbc_lnk: ice-point types (I, J, K) are not used by SOSIE and not implemented"; return SOSIE_ERR_ARG;
  }
  k_lbc_lnk<float><<<nslab, 1024, 0, ctx->stream>>>(dX, (size_t)jpi * jpj, jpi, jpj, nperio, cd_type, psgn, has_pval, pval);
  KLAUNCH_CHECK();
  return SOSIE_OK;
}

// ---------------------------------------------------------------------------------------------------
// post-interpolation block of INTERP_3D (src/mod_interp.f90:447-511), the optional SMOOTHER calls excepted
__global__ void k_i3d_persist_clamp(float* X, const int32_t* __restrict__ mask, size_t n, int nk, int ixtrpl_bot, float vmin, float vmax,
                                    float rmiss) {
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
    int jk_bot = 0;
    float zwet = 0.f;
    if (ixtrpl_bot == 1 && mask[p] == 1) {
      for (int k = 1; k <= nk; k++) if (mask[p + (size_t)(k - 1) * n] == 0) { jk_bot = k; break; }  // FINDLOC(mask_trg(ji,jj,:), 0)
      if (jk_bot > 1) zwet = X[p + (size_t)(jk_bot - 2) * n];
    }
    for (int k = 1; k <= nk; k++) {
      float v = X[p + (size_t)(k - 1) * n];
      if (jk_bot > 1 && k >= jk_bot) v = zwet;                 // persistence into the sea-bed (:447-466)
      if ((v > vmax) && (v != rmiss)) v = vmax;                // (:474-477)
      if ((v < vmin) && (v != rmiss)) v = vmin;
      X[p + (size_t)(k - 1) * n] = v;
    }
  }
}
__global__ void k_i3d_mask(float* X, const int32_t* __restrict__ mask, size_t tot, float rmiss) {
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const float m = (float)mask[t];
    // data3d_trg*mask_trg + (1. - mask_trg)*rmiss_val  (:507)
    X[t] = __fadd_rn(__fmul_rn(X[t], m), __fmul_rn(__fsub_rn(1.f, m), rmiss));
  }
}

int interp3d_post_impl(sosie_ctx* ctx, int ni, int nj, int nk, float* dX, const int32_t* dmask_trg, int ixtrpl_bot, float vmin, float vmax,
                       float rmiss_val, int i_orca_trg, int c_orca_trg, int lmout) {
  if (ni < 1 || nj < 1 || nk < 1 || !dX) { ctx->err = "sosie_interp3d_post: bad arguments"; return SOSIE_ERR_ARG; }
  if ((ixtrpl_bot == 1 || lmout) && !dmask_trg) { ctx->err = "sosie_interp3d_post: mask_trg is required (ixtrpl_bot = 1 or lmout)"; return SOSIE_ERR_ARG; }
  const size_t n = (size_t)ni * nj;
  k_i3d_persist_clamp<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(dX, dmask_trg, n, nk, ixtrpl_bot, vmin, vmax, rmiss_val);
  KLAUNCH_CHECK();
  if (i_orca_trg > 0) { if (int rc = lbc_lnk_impl(ctx, i_orca_trg, ni, nj, nk, dX, c_orca_trg, 1.0, 0, 0.0)) return rc; }
  if (lmout) { k_i3d_mask<<<grid_for(ctx, n * nk, 256, 8), 256, 0, ctx->stream>>>(dX, dmask_trg, n * nk, rmiss_val); KLAUNCH_CHECK(); }
  return SOSIE_OK;
}

// ---------------------------------------------------------------------------------------------------
// extrp_hl (src/mod_interp.f90:521-542): rows beyond the source's latitude range of a regular target take the last
// interpolated row (persistence).  The source rows (jj_ex_top - icr, jj_ex_btm + icr) are never written themselves.
__global__ void k_extrp_hl(float* Xb, int ni, int nj, int jj_ex_top, int jj_ex_btm, int icr) {
  float* X = Xb + (size_t)blockIdx.y * ni * nj;
  const size_t n = (size_t)ni * nj;
  const int jend = (icr + 1) / 2 * nj + (1 - icr) / 2, jbeg = (icr + 1) / 2 + (1 - icr) / 2 * nj;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    const int jj = (int)(k / ni) + 1, i = (int)(k % ni);
    // the two loops of the reference run one after the other: a row in both ranges ends with the bottom source row
    int src = 0;
    if (jj_ex_top > 0 && ((icr > 0) ? (jj >= jj_ex_top && jj <= jend) : (jj <= jj_ex_top && jj >= jend))) src = jj_ex_top - icr;
    if (jj_ex_btm > 0 && ((icr > 0) ? (jj >= jbeg && jj <= jj_ex_btm) : (jj <= jbeg && jj >= jj_ex_btm))) src = -(jj_ex_btm + icr);
    if (src > 0) X[k] = X[(size_t)i + (size_t)(src - 1) * ni];
    else if (src < 0) X[k] = X[(size_t)i + (size_t)(-src - 1) * ni];
  }
}

int extrp_hl_impl(sosie_ctx* ctx, int ni, int nj, int nslab, float* dX, int jj_ex_top, int jj_ex_btm, int nlat_icr_trg) {
  if (ni < 1 || nj < 2 || nslab < 1 || !dX || (nlat_icr_trg != 1 && nlat_icr_trg != -1)) { ctx->err = "sosie_extrp_hl: bad arguments"; return SOSIE_ERR_ARG; }
  const int icr = nlat_icr_trg;
  // source rows must exist and must not be rewritten by the other loop (then the reference reads an already replaced row)
  const int st = jj_ex_top > 0 ? jj_ex_top - icr : 0, sb = jj_ex_btm > 0 ? jj_ex_btm + icr : 0;
  if ((jj_ex_top > 0 && (st < 1 || st > nj || jj_ex_top > nj)) || (jj_ex_btm > 0 && (sb < 1 || sb > nj || jj_ex_btm > nj)) ||
      (jj_ex_top > 0 && jj_ex_btm > 0 && ((icr > 0) ? (jj_ex_btm >= jj_ex_top - 1) : (jj_ex_btm <= jj_ex_top + 1)))) {
    ctx->err = "sosie_extrp_hl: jj_ex_top / jj_ex_btm out of range or overlapping"; return SOSIE_ERR_ARG;
  }
  if (jj_ex_top <= 0 && jj_ex_btm <= 0) return SOSIE_OK;
  k_extrp_hl<<<dim3(grid_for(ctx, (size_t)ni * nj, 256, 4), nslab), 256, 0, ctx->stream>>>(dX, ni, nj, jj_ex_top, jj_ex_btm, icr);
  KLAUNCH_CHECK();
  return SOSIE_OK;
}

// ---------------------------------------------------------------------------------------------------
// angle2 (src/mod_nemotools.f90:607-823): cos / sin of the angle between the grid lines and north at T, U, V, F points
// (the XCOS_t, XSIN_t ... that corr_vect rotates with, src/corr_vect.f90:553-560).  One thread per point for the
// interior, then lbc_lnk (psgn = -1) on a NEMO grid or the three Akima edge extrapolations otherwise.
// rpi = 3.141592653 is a default-REAL literal in the reference (Q1).  Unassigned cells of the reference are DEFINED 0.
#define A2_RPI ((double)3.141592653f)
#define A2_RAD (A2_RPI / (double)180.0f)
__device__ __forceinline__ double a2_tan(double zphi) { return pm_tan(A2_RPI / 4. - A2_RAD * zphi / 2.); }

__global__ void __launch_bounds__(128)
k_angle2(int nx, int ny, int l_cut_180, const double* __restrict__ plamt, const double* __restrict__ pphit,
         const double* __restrict__ plamu, const double* __restrict__ pphiu, const double* __restrict__ plamv,
         const double* __restrict__ pphiv, const double* __restrict__ plamf, const double* __restrict__ pphif, double* out) {
  const size_t n = (size_t)nx * ny;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    const int jj = (int)(k / nx) + 1, ji = (int)(k % nx) + 1;
    double r[8] = {0., 0., 0., 0., 0., 0., 0., 0.};   // cost, sint, cosu, sinu, cosv, sinv, cosf, sinf
    if (jj >= 2 && jj <= ny - 1 && ji >= 2) {
      double zlamt = plamt[k], zlamu = plamu[k], zlamv = plamv[k], zlamf = plamf[k];
      double zlamvjm1 = plamv[k - nx], zlamfjm1 = plamf[k - nx], zlamfim1 = plamf[k - 1], zlamujp1 = plamu[k + nx];
      if (l_cut_180 && ((fabs(zlamt) > 170.) || (fabs(zlamu) > 170.) || (fabs(zlamv) > 170.) || (fabs(zlamf) > 170.))) {
        zlamt = fmod(360. + zlamt, 360.); zlamu = fmod(360. + zlamu, 360.); zlamv = fmod(360. + zlamv, 360.); zlamf = fmod(360. + zlamf, 360.);
        zlamvjm1 = fmod(360. + zlamvjm1, 360.); zlamfjm1 = fmod(360. + zlamfjm1, 360.);
        zlamfim1 = fmod(360. + zlamfim1, 360.); zlamujp1 = fmod(360. + zlamujp1, 360.);
      }
      double st, ct, su, cu, sv, cv, sf, cf, sq, cq;
      pm_sincos(A2_RAD * zlamt, &st, &ct); pm_sincos(A2_RAD * zlamu, &su, &cu);
      pm_sincos(A2_RAD * zlamv, &sv, &cv); pm_sincos(A2_RAD * zlamf, &sf, &cf);
      const double tt = a2_tan(pphit[k]), tu = a2_tan(pphiu[k]), tv = a2_tan(pphiv[k]), tf = a2_tan(pphif[k]);
      const double zxnpt = 0. - 2. * ct * tt, zynpt = 0. - 2. * st * tt, znnpt = zxnpt * zxnpt + zynpt * zynpt;
      const double zxnpu = 0. - 2. * cu * tu, zynpu = 0. - 2. * su * tu, znnpu = zxnpu * zxnpu + zynpu * zynpu;
      const double zxnpv = 0. - 2. * cv * tv, zynpv = 0. - 2. * sv * tv, znnpv = zxnpv * zxnpv + zynpv * zynpv;
      const double zxnpf = 0. - 2. * cf * tf, zynpf = 0. - 2. * sf * tf, znnpf = zxnpf * zxnpf + zynpf * zynpf;
      double th;
      th = a2_tan(pphiv[k - nx]); pm_sincos(A2_RAD * zlamvjm1, &sq, &cq);
      const double zxvvt = 2. * cv * tv - 2. * cq * th, zyvvt = 2. * sv * tv - 2. * sq * th;
      const double znvvt = fmax(sqrt(znnpt * (zxvvt * zxvvt + zyvvt * zyvvt)), (double)1.e-14f);
      th = a2_tan(pphif[k - nx]); pm_sincos(A2_RAD * zlamfjm1, &sq, &cq);
      const double zxffu = 2. * cf * tf - 2. * cq * th, zyffu = 2. * sf * tf - 2. * sq * th;
      const double znffu = fmax(sqrt(znnpu * (zxffu * zxffu + zyffu * zyffu)), (double)1.e-14f);
      th = a2_tan(pphif[k - 1]); pm_sincos(A2_RAD * zlamfim1, &sq, &cq);
      const double zxffv = 2. * cf * tf - 2. * cq * th, zyffv = 2. * sf * tf - 2. * sq * th;
      const double znffv = fmax(sqrt(znnpv * (zxffv * zxffv + zyffv * zyffv)), (double)1.e-14f);
      th = a2_tan(pphiu[k + nx]); pm_sincos(A2_RAD * zlamujp1, &sq, &cq);
      const double zxuuf = 2. * cq * th - 2. * cu * tu, zyuuf = 2. * sq * th - 2. * su * tu;
      const double znuuf = fmax(sqrt(znnpf * (zxuuf * zxuuf + zyuuf * zyuuf)), (double)1.e-14f);
      r[1] = (zxnpt * zyvvt - zynpt * zxvvt) / znvvt; r[0] = (zxnpt * zxvvt + zynpt * zyvvt) / znvvt;
      r[3] = (zxnpu * zyffu - zynpu * zxffu) / znffu; r[2] = (zxnpu * zxffu + zynpu * zyffu) / znffu;
      r[7] = (zxnpf * zyuuf - zynpf * zxuuf) / znuuf; r[6] = (zxnpf * zxuuf + zynpf * zyuuf) / znuuf;
      r[5] = (zxnpv * zxffv + zynpv * zyffv) / znffv; r[4] = -(zxnpv * zyffv - zynpv * zxffv) / znffv;
    }
#pragma unroll
    for (int q = 0; q < 8; q++) out[(size_t)q * n + k] = r[q];
  }
}

__device__ __forceinline__ double a2_extra_1(double xa, double xb, double xc, double xd, double ya, double yb, double yc) {
  const double A = xb - xa, B = xc - xb, C = xd - xc, ALF = yb - ya, BET = yc - yb;
  if ((A == 0.) || (B == 0.) || (C == 0.)) return yc;
  return C * (2 * BET / B - ALF / A) + yc;
}
// which = 0 northernmost row, 1 southernmost row, 2 westernmost column (in that order, :789-819)
__global__ void k_angle2_edges(int nx, int ny, int which, const double* __restrict__ plamt, const double* __restrict__ pphit,
                               const double* __restrict__ plamu, const double* __restrict__ pphiu, const double* __restrict__ plamv,
                               const double* __restrict__ pphiv, const double* __restrict__ plamf, const double* __restrict__ pphif,
                               double* out) {
  const size_t n = (size_t)nx * ny;
  const double* PH[4] = {pphit, pphiu, pphiv, pphif};
  const double* LA[4] = {plamt, plamu, plamv, plamf};
  const int len = (which == 2) ? ny : nx;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < len * 8; t += gridDim.x * blockDim.x) {
    const int q = t / len, e = t % len + 1;
    double* g = out + (size_t)q * n;
#define GI(i, j) g[(size_t)((i)-1) + (size_t)((j)-1) * nx]
#define CI(A_, i, j) A_[(size_t)((i)-1) + (size_t)((j)-1) * nx]
    if (which == 0) { const double* P = PH[q / 2]; GI(e, ny) = a2_extra_1(CI(P, e, ny - 3), CI(P, e, ny - 2), CI(P, e, ny - 1), CI(P, e, ny), GI(e, ny - 3), GI(e, ny - 2), GI(e, ny - 1)); }
    else if (which == 1) { const double* P = PH[q / 2]; GI(e, 1) = a2_extra_1(CI(P, e, 4), CI(P, e, 3), CI(P, e, 2), CI(P, e, 1), GI(e, 4), GI(e, 3), GI(e, 2)); }
    else { const double* L = LA[q / 2]; GI(1, e) = a2_extra_1(CI(L, 4, e), CI(L, 3, e), CI(L, 2, e), CI(L, 1, e), GI(4, e), GI(3, e), GI(2, e)); }
#undef GI
#undef CI
  }
}

// din: 8 coordinate arrays (lamt, phit, lamu, phiu, lamv, phiv, lamf, phif), each (nx,ny), contiguous one after the other
// dout: the 8 results in the reference's argument order (cost, sint, cosu, sinu, cosv, sinv, cosf, sinf)
int angle2_impl(sosie_ctx* ctx, int nperio, int nx, int ny, const double* din, double* dout) {
  if (nx < 4 || ny < 5 || !din || !dout) { ctx->err = "sosie_angle2: bad arguments"; return SOSIE_ERR_ARG; }
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)nx * ny;
  const double* a[8];
  for (int q = 0; q < 8; q++) a[q] = din + (size_t)q * n;
  // l_cut_180 (:661-662): MAXVAL / MINVAL of plamt
  std::vector<double> hl(n);   // small grids only reach this routine (one horizontal grid); read back and reduce on the host
  CK(cudaMemcpyAsync(hl.data(), a[0], n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  double mx = hl[0], mn = hl[0];
  for (size_t k = 1; k < n; k++) { mx = std::max(mx, hl[k]); mn = std::min(mn, hl[k]); }
  const int l_cut_180 = ((mx <= 180.) && (mx > 175.) && (mn >= -180.) && (mn < -175.)) ? 1 : 0;
  k_angle2<<<grid_for(ctx, n, 128, 8), 128, 0, st>>>(nx, ny, l_cut_180, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], dout); KLAUNCH_CHECK();
  if (nperio > 0) {
    static const char cd[4] = {'T', 'U', 'V', 'F'};
    for (int q = 0; q < 4; q++) {   // cos and sin of one point type are adjacent: two slabs per launch
      k_lbc_lnk<double><<<2, 1024, 0, st>>>(dout + (size_t)(2 * q) * n, n, nx, ny, nperio, cd[q], -1.0, 0, 0.0); KLAUNCH_CHECK();
    }
  } else {
    for (int which = 0; which < 3; which++) {
      const int len = (which == 2) ? ny : nx;
      k_angle2_edges<<<cdiv((int64_t)len * 8, 128), 128, 0, st>>>(nx, ny, which, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], dout); KLAUNCH_CHECK();
    }
  }
  return SOSIE_OK;
}
