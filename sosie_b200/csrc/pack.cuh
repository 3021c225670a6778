// pack.cuh -- the 32-byte apply record of one target, shared by the pack kernel (bilin_apply.cu) and the fused
// EASY mapping kernel (bilin_map.cu), which writes the record while the mapping is still in registers.
#pragma once
#include "geom.cuh"

#define REC_COPY (-1)   // idx.x: copy Z1w(idx.y)
#define REC_CONST (-2)  // idx.x: constant w.x

// (iP,jP,iqd,alpha,beta) -> int4 corner offsets in the working source slab (column-major, pole rows included) +
// float4 REAL(4) weights exactly as INTERP_BL rounds them (src/mod_bilin_2d.f90:299-337); the identical-point copy
// of the apply loop (:227-230) and INTERP_BL's error codes (:347-356) are encoded in the record.
// far_from_nearest: the caller KNOWS the target is not the identical-point case and did not load (lon_t, lat_t) -- see
// PACK_FAR_KM below.  false: the test of the apply loop is evaluated literally.
//
// PACK_FAR_KM: the mapping stores dist_np = DISTANCE(target, nearest source point) in km (REAL(4)).  The identical-point
// copy needs both coordinate differences below 1e-5 degrees, i.e. the two points less than 1.6 m apart, for which DISTANCE
// (unit vectors, ACOS) returns at most 0.0017 km; beyond 0.01 km the test cannot pass, and its 16 bytes of target
// coordinates per target need not be read.  A mapping without distances (loaded from a file: dist_np zeroed) and targets
// the body never reached (dist_np 0) take the literal test.
#define PACK_FAR_KM 0.01f
__device__ __forceinline__ void pack_record(const SrcGeom& sg, int k_ew, int iP_raw, int jP_raw, int iqd, double xa, double xb,
                                            double lon_t, double lat_t, int4& id, float4& w, bool far_from_nearest = false) {
  const int nxi = sg.nx, nyi = sg.nyw;
  const double eps5 = (double)1.E-5f;
  int iP = max(iP_raw, 1), jP = max(jP_raw, 1);  // IMETRICS(:,:,1:2) = MAX(IMETRICS(:,:,1:2), 1)  (:217)
  w = make_float4(0.f, 0.f, 0.f, 0.f);
  // a stale mapping could hold iP/jP beyond the source: the reference would read out of bounds
  int iPc = min(iP, nxi), jPc = min(jP, nyi);
  bool same = !far_from_nearest && (fabs(d_degE_to_degWE(src_lon(sg, iPc, jPc)) - d_degE_to_degWE(lon_t)) < eps5) &&
              (fabs(src_lat(sg, iPc, jPc) - lat_t) < eps5);
  if (same) {
    id = make_int4(REC_COPY, (iPc - 1) + (jPc - 1) * nxi, 0, 0);
    return;
  }
  int jiPm1 = iP - 1, jiPp1 = iP + 1;
  if ((jiPm1 == 0) && (k_ew >= 0)) jiPm1 = nxi - k_ew;
  if ((jiPp1 == nxi + 1) && (k_ew >= 0)) jiPp1 = 1 + k_ew;
  int i1 = 0, j1 = 0, i2 = 0, j2 = 0, i3 = 0, j3 = 0, i4 = 0, j4 = 0;  // Q19 (iqd outside 1..4 cannot occur after the fix-ups)
  switch (iqd) {
    case 1: i1 = iP; j1 = jP; i2 = jiPp1; j2 = jP; i3 = jiPp1; j3 = jP + 1; i4 = iP; j4 = jP + 1; break;
    case 2: i1 = iP; j1 = jP; i2 = iP; j2 = jP - 1; i3 = jiPp1; j3 = jP - 1; i4 = jiPp1; j4 = jP; break;
    case 3: i1 = iP; j1 = jP; i2 = jiPm1; j2 = jP; i3 = jiPm1; j3 = jP - 1; i4 = iP; j4 = jP - 1; break;
    case 4: i1 = iP; j1 = jP; i2 = iP; j2 = jP + 1; i3 = jiPm1; j3 = jP + 1; i4 = jiPm1; j4 = jP; break;
    default: break;
  }
  float w1 = (float)((1.0 - xa) * (1.0 - xb));
  float w2 = (float)(xa * (1.0 - xb));
  float w3 = (float)(xa * xb);
  float w4 = (float)((1.0 - xa) * xb);
  float wup = w1 + w2 + w3 + w4;
  float cst = 0.f;
  bool isc = true;
  if (wup == 0.f) cst = -9998.f;
  else if ((i1 < 1) || (i2 < 1) || (i3 < 1) || (i4 < 1)) cst = -9997.f;
  else if ((j1 < 1) || (j2 < 1) || (j3 < 1) || (j4 < 1)) cst = -9996.f;
  else if ((j1 > nyi) || (j2 > nyi) || (j3 > nyi) || (j4 > nyi)) cst = -9995.f;
  else if ((i1 > nxi) || (i2 > nxi) || (i3 > nxi) || (i4 > nxi)) cst = -9994.f;
  else isc = false;
  if (isc) { id = make_int4(REC_CONST, 0, 0, 0); w.x = cst; }
  else {
    id = make_int4((i1 - 1) + (j1 - 1) * nxi, (i2 - 1) + (j2 - 1) * nxi, (i3 - 1) + (j3 - 1) * nxi, (i4 - 1) + (j4 - 1) * nxi);
    w = make_float4(w1, w2, w3, w4);
  }
}
