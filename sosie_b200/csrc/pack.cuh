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
  // The four corners of quadrant iqd (:299-327), counter-clockwise from (iP,jP):
  //   1: (iP,jP) (iP+1,jP) (iP+1,jP+1) (iP,jP+1)      2: (iP,jP) (iP,jP-1) (iP+1,jP-1) (iP+1,jP)
  //   3: (iP,jP) (iP-1,jP) (iP-1,jP-1) (iP,jP-1)      4: (iP,jP) (iP,jP+1) (iP-1,jP+1) (iP-1,jP)
  // i.e. the opposite corner is (ix,jy) with ix east for 1, 2 / west for 3, 4 and jy north for 1, 4 / south for 2, 3; odd
  // quadrants walk along i first, even ones along j first.  Selects instead of a four-way branch (the warps of the apply
  // kernels mix quadrants).  iqd outside 1..4 (Q19: cannot occur after the fix-ups of MAPPING_BL) leaves all eight at 0.
  const bool qok = (iqd >= 1) && (iqd <= 4);
  const int ix = qok ? ((iqd <= 2) ? jiPp1 : jiPm1) : 0;
  const int jy = qok ? ((iqd == 1 || iqd == 4) ? jP + 1 : jP - 1) : 0;
  const bool odd = (iqd & 1) != 0;
  const int i1 = qok ? iP : 0, j1 = qok ? jP : 0;
  const int i2 = odd ? ix : i1, j2 = odd ? j1 : jy;
  const int i3 = ix, j3 = jy;
  const int i4 = odd ? i1 : ix, j4 = odd ? jy : j1;
  float w1 = (float)((1.0 - xa) * (1.0 - xb));
  float w2 = (float)(xa * (1.0 - xb));
  float w3 = (float)(xa * xb);
  float w4 = (float)((1.0 - xa) * xb);
  float wup = w1 + w2 + w3 + w4;
  float cst = 0.f;
  bool isc = true;
  // INTERP_BL's index tests (:347-356), in its order; {i1..i4} = {i1, ix} and {j1..j4} = {j1, jy}
  if (wup == 0.f) cst = -9998.f;
  else if (min(i1, ix) < 1) cst = -9997.f;
  else if (min(j1, jy) < 1) cst = -9996.f;
  else if (max(j1, jy) > nyi) cst = -9995.f;
  else if (max(i1, ix) > nxi) cst = -9994.f;
  else isc = false;
  if (isc) { id = make_int4(REC_CONST, 0, 0, 0); w.x = cst; }
  else {
    id = make_int4((i1 - 1) + (j1 - 1) * nxi, (i2 - 1) + (j2 - 1) * nxi, (i3 - 1) + (j3 - 1) * nxi, (i4 - 1) + (j4 - 1) * nxi);
    w = make_float4(w1, w2, w3, w4);
  }
}
