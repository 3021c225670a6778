// bilin_map.cu -- device construction of the bilinear mapping.
//
// GPU counterpart of MAPPING_BL (src/mod_bilin_2d.f90:366-691) and of the nearest-point search
// FIND_NEAREST_POINT / _EASY / _TWISTED (src/mod_manip.f90:834-1440).  Results are bit-identical to
// the reference algorithm evaluated with the pmath libm (see pmath.h): every decision is taken
// by the same IEEE operations in the same order; only the *schedule* is different:
//   * EASY (regular source): one thread per target, search + quadrant + Newton fused in one kernel;
//     the separable 1-D MINLOCs are bracket searches with the reference's first-index tie-break;
//     the per-target 1.E12 fill of the whole source array (:1038) is a dead store and is skipped.
//   * TWISTED (curvilinear source): the chain-independent latitude-band result is computed for all
//     targets in parallel (one warp per target), then the sequential "luck window" chain
//     s(t) = f_t(s(t-1)) (:1342-1349) is solved by parallel fixed-point iteration -- the fixed point
//     of a chain recurrence is unique, so the converged state IS the sequential result.
//   * acos is only evaluated for candidates whose dot product is within 1e-13 of the running best
//     (a margin 100x larger than any rounding wiggle of acos): identical argmin, ~10x fewer acos.
#include <cub/cub.cuh>

#include <algorithm>
#include "geom.cuh"

// ---------------------------------------------------------------------------------------------------
// order-preserving encoding of doubles for atomicMin/atomicMax
__device__ __forceinline__ unsigned long long enc_d(double x) {
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
__device__ __forceinline__ double dec_d(unsigned long long u) {
  unsigned long long v = (u >> 63) ? (u & 0x7fffffffffffffffULL) : ~u;
  return __longlong_as_double((long long)v);
}
static inline double dec_d_host(unsigned long long u) {
  unsigned long long v = (u >> 63) ? (u & 0x7fffffffffffffffULL) : ~u;
  double x;
  memcpy(&x, &v, 8);
  return x;
}

struct Prelude {            // scalars produced on device, read back once
  unsigned long long ymin_s, ymax_s, ymin_t, ymax_t, rmin_dlat;  // encoded doubles
  int irreg_s, irreg_t;     // != 0 : L_IS_GRID_REGULAR false
  int j_strt, j_stop;       // MINVAL / MAXVAL over columns
  int err;
  double rtmp;              // SUM(ztmp_t(:,ny_t/2))
  unsigned long long ymaxabs_dummy;
};

__global__ void k_prelude_init(Prelude* p) {
  p->ymin_s = ~0ULL; p->ymax_s = 0ULL; p->ymin_t = ~0ULL; p->ymax_t = 0ULL; p->rmin_dlat = ~0ULL;
  p->irreg_s = 0; p->irreg_t = 0; p->j_strt = 2147483647; p->j_stop = 0; p->err = 0; p->rtmp = 0.0;
}

__global__ void k_minmax(const double* __restrict__ a, size_t n, unsigned long long* omin, unsigned long long* omax) {
  unsigned long long lmin = ~0ULL, lmax = 0ULL;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    unsigned long long e = enc_d(a[k]);
    lmin = min(lmin, e);
    lmax = max(lmax, e);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
    lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(omin, lmin); atomicMax(omax, lmax); }
}

// L_IS_GRID_REGULAR (src/mod_manip.f90:1632-1663), both tests, on a 2-D (nx,ny) pair
// ya..yb (1-based): the rows the column test b/ sums over -- all of them, or (prelude exchange) one rank's rows, where a
// partial sum above the tolerance already proves the total is
__global__ void k_is_regular(const double* __restrict__ X, const double* __restrict__ Y, int nx, int ny, int ja, int jb,
                             int* irreg, int ya = 1, int yb = 0x7fffffff) {
  const double tol = (double)1.E-12f;
  __shared__ double sh[32];
  __shared__ int stop;
  // a/ one block per row jj>=2 : SUM(ABS(Xlon(:,jj)-Xlon(:,1)))
  for (int jj = ja + blockIdx.x; jj <= jb; jj += gridDim.x) {   // all rows: ja = 2, jb = ny
    // the reference EXITs at the first irregular row (:1651): stop early, block-uniformly
    if (threadIdx.x == 0) stop = *(volatile int*)irreg;
    __syncthreads();
    if (stop) return;
    double s = 0.0;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) s += fabs(X[(size_t)i + (size_t)(jj - 1) * nx] - X[i]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (blockDim.x >> 5); w++) t += sh[w];
      if (t > tol) atomicOr(irreg, 1);
    }
    __syncthreads();
  }
  // b/ one thread per column ji : SUM(ABS(Xlat(ji,:)-Xlat(1,:))), sequential in j as the reference
  //    (only evaluated when a/ passed, :1655)
  if (*(volatile int*)irreg) return;
  for (int ji = blockIdx.x * blockDim.x + threadIdx.x; ji < nx; ji += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int jj = max(ya, 1) - 1; jj < min(yb, ny); jj++) s += fabs(Y[(size_t)ji + (size_t)jj * nx] - Y[(size_t)jj * nx]);
    if (s > tol) atomicOr(irreg, 2);
  }
}

// rtmp = SUM(Ytrg(:,jm) - Ytrg(:,jm-1)) (:921); only SIGN(1., rtmp) is used (:922).  A floating-point sum of terms that
// are all >= 0 (or all <= 0) has that sign in any order, so the usual case (latitudes monotone along j on that row pair)
// is decided by a parallel sign census and rtmp is set to +-(number of non-zero terms); only mixed signs need the
// sequential sum in index order, as SUM() does.
__global__ void k_rtmp(TrgGeom tg, Prelude* p) {
  __shared__ double sh[1024];
  __shared__ int npos, nneg;
  int jm = tg.ny / 2;
  if (threadIdx.x == 0) { npos = 0; nneg = 0; }
  __syncthreads();
  if (jm >= 2) {
    int lp = 0, ln = 0;
    for (int i = 1 + threadIdx.x; i <= tg.nx; i += blockDim.x) {
      const double d = trg_lat(tg, i, jm) - trg_lat(tg, i, jm - 1);
      lp += (d > 0.0); ln += (d < 0.0) || (d != d);   // a NaN term: let the sequential sum produce the reference's NaN
    }
    if (lp) atomicAdd(&npos, lp);
    if (ln) atomicAdd(&nneg, ln);
  }
  __syncthreads();
  if (jm < 2 || nneg == 0 || npos == 0) {
    // all terms >= 0 -> sum >= 0 (zero only if every term is zero: SIGN(1., +0.) = +1 as for any non-negative sum);
    // all terms <= 0 -> sum <= 0; a sum of non-positive terms that are all zero is +0 in Fortran's SUM (starts from +0)
    if (threadIdx.x == 0) p->rtmp = (jm >= 2 && nneg > 0) ? -(double)nneg : (double)npos;
    return;
  }
  double r = 0.0;  // carried by thread 0 only
  for (int i0 = 1; i0 <= tg.nx; i0 += 1024) {
    int i = i0 + threadIdx.x;
    if (i <= tg.nx) sh[threadIdx.x] = trg_lat(tg, i, jm) - trg_lat(tg, i, jm - 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      int n = min(1024, tg.nx - i0 + 1);
      for (int k = 0; k < n; k++) r += sh[k];  // sequential, in index order, as SUM() does
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) p->rtmp = r;
}

// the head of the device-planned prelude in ONE single-block launch: k_prelude_init, MINVAL/MAXVAL of a 1-D source latitude
// axis (k_minmax on a few thousand values) and k_rtmp
__global__ void __launch_bounds__(1024) k_prelude_head(TrgGeom tg, const double* __restrict__ lat1, int nyw, int do_dlat, Prelude* p) {
  __shared__ unsigned long long smin[32], smax[32];
  unsigned long long lmin = ~0ULL, lmax = 0ULL;
  for (int k = threadIdx.x; k < nyw; k += blockDim.x) { const unsigned long long e = enc_d(lat1[k]); lmin = min(lmin, e); lmax = max(lmax, e); }
  for (int o = 16; o > 0; o >>= 1) { lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o)); lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o)); }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lmin; smax[threadIdx.x >> 5] = lmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) { smin[0] = min(smin[0], smin[w]); smax[0] = max(smax[0], smax[w]); }
    p->ymin_s = smin[0]; p->ymax_s = smax[0]; p->ymin_t = ~0ULL; p->ymax_t = 0ULL; p->rmin_dlat = ~0ULL;
    p->irreg_s = 0; p->irreg_t = 0; p->j_strt = 2147483647; p->j_stop = 0; p->err = 0; p->rtmp = 0.0;
  }
  __syncthreads();
  // Early verdict for k_is_regular (which follows on the stream): ONE term |Xlon(i,2) - Xlon(i,1)| above the tolerance makes the
  // row sum of L_IS_GRID_REGULAR (:1647-1652) exceed it -- a sum of non-negative terms is at least each of them, in any
  // order -- so a curvilinear target is settled here from 1024 columns of two rows, and every block of k_is_regular leaves
  // at its first look at the flag instead of summing a first wave of 592 rows (40 MB on the eNATL60-shaped target).
  if (!tg.is_1d && tg.ny >= 2) {
    const double tol = (double)1.E-12f;
    int hit = 0;
    for (int i = threadIdx.x; i < min(tg.nx, 1024); i += blockDim.x) hit |= (fabs(tg.X[(size_t)i + (size_t)tg.nx] - tg.X[i]) > tol);
    if (__syncthreads_or(hit) && threadIdx.x == 0) atomicOr(&p->irreg_t, 1);
  }
  if (!do_dlat) return;
  // ---- k_rtmp (same code: sign census, sequential sum only for mixed signs)
  __shared__ double sh[1024];
  __shared__ int npos, nneg;
  const int jm = tg.ny / 2;
  if (threadIdx.x == 0) { npos = 0; nneg = 0; }
  __syncthreads();
  if (jm >= 2) {
    int lp = 0, ln = 0;
    for (int i = 1 + threadIdx.x; i <= tg.nx; i += blockDim.x) {
      const double d = trg_lat(tg, i, jm) - trg_lat(tg, i, jm - 1);
      lp += (d > 0.0); ln += (d < 0.0) || (d != d);
    }
    if (lp) atomicAdd(&npos, lp);
    if (ln) atomicAdd(&nneg, ln);
  }
  __syncthreads();
  if (jm < 2 || nneg == 0 || npos == 0) {
    if (threadIdx.x == 0) p->rtmp = (jm >= 2 && nneg > 0) ? -(double)nneg : (double)npos;
    return;
  }
  double r = 0.0;
  for (int i0 = 1; i0 <= tg.nx; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    if (i <= tg.nx) sh[threadIdx.x] = trg_lat(tg, i, jm) - trg_lat(tg, i, jm - 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      const int n = min(1024, tg.nx - i0 + 1);
      for (int k = 0; k < n; k++) r += sh[k];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) p->rtmp = r;
}

// i1dum = MINLOC(Ytrg, mask=(Ytrg>=y_min_s), dim=2) ; MAXLOC(Ytrg, mask=(Ytrg<=y_max_s), dim=2)  (:939-941): see k_ypass
__global__ void k_fill_i32(int32_t* a, size_t n, int32_t v);
#define JR_ROWS 64
// ONE pass over the target latitudes for everything the prelude reduces from them: MINVAL/MAXVAL(Ytrg) (:888-889),
// rmin_dlat_dj = MINVAL(SIGN(1,rtmp)*(Ytrg(:,j) - Ytrg(:,j-1))) (:921-923) and, per (column, chunk of JR_ROWS rows), the
// extreme value under the mask of :939-941 together with the FIRST row attaining it (MINLOC / MAXLOC return the first
// occurrence).  y_min_s / y_max_s / rtmp are read from the Prelude on the device (written by earlier kernels of the
// stream), so no host round trip separates the reductions.  All of them are min/max: order independent, exact.
struct ColPart {
  unsigned long long emin, emax;   // order-preserving encodings; ~0 / 0 = mask empty in this chunk
  int jmin, jmax;                  // first rows (1-based) attaining them
};
__global__ void k_ypass(TrgGeom tg, Prelude* p, int do_dlat, ColPart* part, int jfirst, int jlast) {
  const double y_min_s = dec_d(p->ymin_s), y_max_s = dec_d(p->ymax_s);   // MINVAL / MAXVAL(Ysrc), reduced by k_minmax before
  const double sgn = d_sign(1.0, p->rtmp);
  const int nchunk = (jlast - jfirst + 1 + JR_ROWS - 1) / JR_ROWS;   // rows jfirst..jlast (1-based): all of them, or one rank's
  const size_t total = (size_t)tg.nx * nchunk;
  unsigned long long gmin = ~0ULL, gmax = 0ULL, gdl = ~0ULL;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t % tg.nx) + 1, c = (int)(t / tg.nx);
    const int j0 = jfirst + c * JR_ROWS, j1 = min(jlast, j0 + JR_ROWS - 1);
    ColPart cp;
    cp.emin = ~0ULL; cp.emax = 0ULL; cp.jmin = 0; cp.jmax = 0;
    double prev = (j0 > 1 && do_dlat) ? trg_lat(tg, i, j0 - 1) : 0.0;
    // the loads of a batch of rows are issued together (they are independent; the reductions below are not)
    const double* __restrict__ colp = tg.is_1d ? tg.Y : tg.Y + (i - 1);
    const size_t rstride = tg.is_1d ? 1 : (size_t)tg.nx;
    for (int jb = j0; jb <= j1; jb += 8) {
      double vv[8];
#pragma unroll
      for (int q = 0; q < 8; q++) vv[q] = (jb + q <= j1) ? __ldcs(colp + (size_t)(jb + q - 1) * rstride) : 0.0;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const int j = jb + q;
        if (j > j1) break;
        const double v = vv[q];
        const unsigned long long e = enc_d(v);
        gmin = min(gmin, e); gmax = max(gmax, e);
        if (do_dlat && j >= 2) gdl = min(gdl, enc_d(sgn * (v - prev)));
        prev = v;
        if ((v >= y_min_s) && e < cp.emin) { cp.emin = e; cp.jmin = j; }   // strict: the first occurrence wins
        if ((v <= y_max_s) && (e > cp.emax || cp.jmax == 0)) { cp.emax = e; cp.jmax = j; }
      }
    }
    part[t] = cp;   // [chunk][column]
  }
  for (int o = 16; o > 0; o >>= 1) {
    gmin = min(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
    gmax = max(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    gdl = min(gdl, __shfl_xor_sync(0xffffffffu, gdl, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&p->ymin_t, gmin); atomicMax(&p->ymax_t, gmax);
    if (do_dlat) atomicMin(&p->rmin_dlat, gdl);
  }
}
// per column: combine the chunks in row order, then j_strt = MINVAL(i1dum, mask=(i1dum>0)), j_stop = MAXVAL(MAXLOC(...))
__global__ void k_jrange_cols(int nx, int nchunk, const ColPart* __restrict__ part, Prelude* p) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x) {
    unsigned long long emin = ~0ULL, emax = 0ULL;
    int jmin = 0, jmax = 0;
    for (int c = 0; c < nchunk; c++) {
      const ColPart cp = part[(size_t)c * nx + i];
      if (cp.jmin > 0 && cp.emin < emin) { emin = cp.emin; jmin = cp.jmin; }
      if (cp.jmax > 0 && (cp.emax > emax || jmax == 0)) { emax = cp.emax; jmax = cp.jmax; }
    }
    if (jmin > 0) atomicMin(&p->j_strt, jmin);
    atomicMax(&p->j_stop, jmax);   // MAXLOC of an empty mask is 0
  }
}
// One rank's contribution to the prelude exchanges: [XchgHead | ColPart per column]
struct XchgHead {
  int32_t j_a, j_b;         // rows reduced by this rank (1-based, inclusive)
  int32_t irreg_t, nx;
  unsigned long long ymin_t, ymax_t, rmin_dlat;
  double rtmp;
};
// the same combination, kept per column: what one rank contributes to the prelude exchange (head != nullptr: the device-side
// exchange, whose record header is written here too -- k_ypass, which reduces the scalars, has finished)
__global__ void k_colpart_combine(int nx, int nchunk, const ColPart* __restrict__ part, ColPart* out, const Prelude* p = nullptr, int ja = 0, int jb = 0,
                                  XchgHead* head = nullptr) {
  if (head && blockIdx.x == 0 && threadIdx.x == 0) {
    head->j_a = ja; head->j_b = jb; head->irreg_t = p->irreg_t; head->nx = nx;
    head->ymin_t = p->ymin_t; head->ymax_t = p->ymax_t; head->rmin_dlat = p->rmin_dlat; head->rtmp = p->rtmp;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x) {
    ColPart r;
    r.emin = ~0ULL; r.emax = 0ULL; r.jmin = 0; r.jmax = 0;
    for (int c = 0; c < nchunk; c++) {
      const ColPart cp = part[(size_t)c * nx + i];
      if (cp.jmin > 0 && cp.emin < r.emin) { r.emin = cp.emin; r.jmin = cp.jmin; }
      if (cp.jmax > 0 && (cp.emax > r.emax || r.jmax == 0)) { r.emax = cp.emax; r.jmax = cp.jmax; }
    }
    out[i] = r;
  }
}
// ---- the prelude exchange as ONE producer kernel and ONE consumer kernel over peer memory (no collective call) -----------
// Every rank owns a mailbox (sosie_p2p_init) that its peers mapped through CUDA IPC (sosie_p2p_connect).  k_xchg_push is
// k_colpart_combine whose output goes, with plain stores over NVLink, into slot [parity][my rank] of EVERY rank's mailbox;
// the last block to finish raises flag [parity][my rank] = epoch on every rank (data first, system-scope fence, then the
// flag).  k_xchg_merge is k_prelude_merge on the local mailbox once all nranks flags show the epoch.  Two parities: a
// rank cannot be two exchanges ahead of a peer (its merge of exchange e+1 needs the peer's push of e+1, which the peer
// issues after its own merge of e), so the slot it overwrites for e+2 has been consumed.  The consumer's wait is bounded
// (ctx->p2p_spin_cycles: 10 s unless SOSIE_P2P_TIMEOUT_MS says otherwise): a peer that never arrives is an error at the next
// synchronising entry, not a hung GPU.
__host__ __device__ inline size_t p2p_flags_offset(int nranks, size_t slot) { return 2 * (size_t)nranks * slot; }
__global__ void __launch_bounds__(128)
k_xchg_push(int nx, int nchunk, const ColPart* __restrict__ part, const Prelude* p, int ja, int jb, unsigned char* const* __restrict__ peers,
            int nranks, int rank, size_t slot, int par, unsigned epoch, unsigned* done) {
  const size_t my = ((size_t)par * nranks + rank) * slot;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    XchgHead h;
    h.j_a = ja; h.j_b = jb; h.irreg_t = p->irreg_t; h.nx = nx;
    h.ymin_t = p->ymin_t; h.ymax_t = p->ymax_t; h.rmin_dlat = p->rmin_dlat; h.rtmp = p->rtmp;
    for (int q = 0; q < nranks; q++) *reinterpret_cast<XchgHead*>(peers[q] + my) = h;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x) {
    ColPart r;
    r.emin = ~0ULL; r.emax = 0ULL; r.jmin = 0; r.jmax = 0;
    for (int c = 0; c < nchunk; c++) {
      const ColPart cp = part[(size_t)c * nx + i];
      if (cp.jmin > 0 && cp.emin < r.emin) { r.emin = cp.emin; r.jmin = cp.jmin; }
      if (cp.jmax > 0 && (cp.emax > r.emax || r.jmax == 0)) { r.emax = cp.emax; r.jmax = cp.jmax; }
    }
    for (int q = 0; q < nranks; q++) reinterpret_cast<ColPart*>(peers[q] + my + sizeof(XchgHead))[i] = r;
  }
  __threadfence_system();   // this thread's stores are visible to every GPU before the block reports
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(done, 1u);
    if (t == gridDim.x - 1) {   // every block has reported: the record is complete on every rank
      *done = 0;
      __threadfence_system();
      const size_t foff = p2p_flags_offset(nranks, slot) + ((size_t)par * nranks + rank) * sizeof(unsigned);
      for (int q = 0; q < nranks; q++) *reinterpret_cast<volatile unsigned*>(peers[q] + foff) = epoch;
      __threadfence_system();
    }
  }
}
__device__ __forceinline__ ColPart ldcg_colpart(const ColPart* q) {   // L2 only: the lines were written by other GPUs
  ColPart c;
  const unsigned long long* u = reinterpret_cast<const unsigned long long*>(q);
  c.emin = __ldcg(u); c.emax = __ldcg(u + 1);
  const unsigned long long jj = __ldcg(u + 2);
  c.jmin = (int)(unsigned)(jj & 0xffffffffull); c.jmax = (int)(unsigned)(jj >> 32);
  return c;
}
__global__ void __launch_bounds__(256)
k_xchg_merge(int nranks, const unsigned char* __restrict__ box, size_t slot, int par, unsigned epoch, int nx, int ny, Prelude* p, long long spin_cycles) {
  __shared__ int order[256];
  __shared__ int bad;
  const unsigned char* all = box + (size_t)par * nranks * slot;
  if (threadIdx.x == 0) {
    bad = 0;
    const volatile unsigned* flags = reinterpret_cast<const volatile unsigned*>(box + p2p_flags_offset(nranks, slot)) + (size_t)par * nranks;
    const long long t0 = clock64();
    for (int r = 0; r < nranks && !bad; r++)
      while (flags[r] != epoch) {
        if (clock64() - t0 > spin_cycles) { bad = 2; break; }
        __nanosleep(200);
      }
    __threadfence_system();
    if (bad) { if (blockIdx.x == 0) atomicOr(&p->err, 4); }
    else {
      for (int r = 0; r < nranks; r++) order[r] = r;
      for (int a = 1; a < nranks; a++) {   // insertion sort by first row
        const int v = order[a];
        const int ja = __ldcg(&reinterpret_cast<const XchgHead*>(all + slot * (size_t)v)->j_a);
        int b = a - 1;
        while (b >= 0 && __ldcg(&reinterpret_cast<const XchgHead*>(all + slot * (size_t)order[b])->j_a) > ja) { order[b + 1] = order[b]; b--; }
        order[b + 1] = v;
      }
      int expect = 1;
      for (int q = 0; q < nranks; q++) {
        const XchgHead* h = reinterpret_cast<const XchgHead*>(all + slot * (size_t)order[q]);
        const int hja = __ldcg(&h->j_a), hjb = __ldcg(&h->j_b);
        if (__ldcg(&h->nx) != nx || hja != expect || hjb < hja) bad = 1;
        expect = hjb + 1;
        if (blockIdx.x == 0) {
          atomicMin(&p->ymin_t, __ldcg(&h->ymin_t)); atomicMax(&p->ymax_t, __ldcg(&h->ymax_t)); atomicMin(&p->rmin_dlat, __ldcg(&h->rmin_dlat));
        }
      }
      if (expect != ny + 1) bad = 1;
      if (bad && blockIdx.x == 0) atomicOr(&p->err, 2);
    }
  }
  __syncthreads();
  if (bad) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x) {
    unsigned long long emin = ~0ULL, emax = 0ULL;
    int jmin = 0, jmax = 0;
    for (int q = 0; q < nranks; q++) {
      const ColPart cp = ldcg_colpart(reinterpret_cast<const ColPart*>(all + slot * (size_t)order[q] + sizeof(XchgHead)) + i);
      if (cp.jmin > 0 && cp.emin < emin) { emin = cp.emin; jmin = cp.jmin; }
      if (cp.jmax > 0 && (cp.emax > emax || jmax == 0)) { emax = cp.emax; jmax = cp.jmax; }
    }
    if (jmin > 0) atomicMin(&p->j_strt, jmin);
    atomicMax(&p->j_stop, jmax);
  }
}

// ---- the prelude's decisions ON THE DEVICE (regular source: the EASY search is known to run, only its row range depends
// on the target) -- the host part prelude_finish() restated as a kernel, so that sosie_bilin_map_async needs no round trip
struct DevPlan { int jlo, jhi, j_strt, j_stop, icr, err; };   // err 1: rows not covered (the reference loops out of bounds), 2: shards do not tile the grid
__global__ void k_prelude_finish(Prelude* p, int nx, int ny, DevPlan* plan) {
  double rmin_dlat_dj = (double)-100.f;
  if (nx > 2 && ny >= 2) rmin_dlat_dj = dec_d(p->rmin_dlat);
  int j_strt = 1, j_stop = ny, icr = 1, err = p->err;
  if ((rmin_dlat_dj > (double)-1.E-12f) || (p->irreg_t == 0)) {
    j_strt = p->j_strt; j_stop = p->j_stop;
    if (j_strt > j_stop) icr = -1;
    if (j_strt < 1 || j_strt > ny || j_stop < 1 || j_stop > ny) err |= 1;
  }
  plan->j_strt = j_strt; plan->j_stop = j_stop; plan->icr = icr; plan->err = err;
  plan->jlo = err ? 1 : min(j_strt, j_stop);
  plan->jhi = err ? 0 : max(j_strt, j_stop);   // on error nothing is searched; the host reports it at its next synchronisation
}

// combine the ranks' blobs in row order into the Prelude a pass over the whole array gives (min / max / first occurrence)
__global__ void k_prelude_merge(int nranks, const unsigned char* __restrict__ all, size_t blob, int nx, int ny, Prelude* p) {
  __shared__ int order[256];
  __shared__ int bad;
  if (threadIdx.x == 0) {
    bad = 0;
    for (int r = 0; r < nranks; r++) order[r] = r;
    for (int a = 1; a < nranks; a++) {   // insertion sort by first row
      const int v = order[a];
      const int ja = reinterpret_cast<const XchgHead*>(all + blob * (size_t)v)->j_a;
      int b = a - 1;
      while (b >= 0 && reinterpret_cast<const XchgHead*>(all + blob * (size_t)order[b])->j_a > ja) { order[b + 1] = order[b]; b--; }
      order[b + 1] = v;
    }
    int expect = 1;
    for (int q = 0; q < nranks; q++) {
      const XchgHead* h = reinterpret_cast<const XchgHead*>(all + blob * (size_t)order[q]);
      if (h->nx != nx || h->j_a != expect || h->j_b < h->j_a) bad = 1;
      expect = h->j_b + 1;
      if (blockIdx.x == 0) {
        atomicMin(&p->ymin_t, h->ymin_t); atomicMax(&p->ymax_t, h->ymax_t); atomicMin(&p->rmin_dlat, h->rmin_dlat);
      }
    }
    if (expect != ny + 1) bad = 1;
    if (bad && blockIdx.x == 0) atomicOr(&p->err, 2);
  }
  __syncthreads();
  if (bad) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x) {
    unsigned long long emin = ~0ULL, emax = 0ULL;
    int jmin = 0, jmax = 0;
    for (int q = 0; q < nranks; q++) {
      const ColPart cp = reinterpret_cast<const ColPart*>(all + blob * (size_t)order[q] + sizeof(XchgHead))[i];
      if (cp.jmin > 0 && cp.emin < emin) { emin = cp.emin; jmin = cp.jmin; }
      if (cp.jmax > 0 && (cp.emax > emax || jmax == 0)) { emax = cp.emax; jmax = cp.jmax; }
    }
    if (jmin > 0) atomicMin(&p->j_strt, jmin);
    atomicMax(&p->j_stop, jmax);
  }
}

__global__ void k_fill_u64(unsigned long long* a, size_t n, unsigned long long v) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) a[k] = v;
}

// ---------------------------------------------------------------------------------------------------
// source preparation
__global__ void k_trig_table(const double* __restrict__ deg, int n, double* c, double* s) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    double sv, cv;
    pm_sincos(deg[k] * G_ZCONV, &sv, &cv);
    c[k] = cv;
    s[k] = sv;
  }
}
// 2-D working arrays (nx,nyw): copy + the two rows of EXT_NORTH_TO_90_REGG (:1813-1819), unit vectors
__global__ void k_build_src2d(const double* __restrict__ Xin, const double* __restrict__ Yin, int nx, int ny, int nyw,
                              int in_1d, double dyp, double* X, double* Y, double* vx, double* vy, double* vz) {
  size_t n = (size_t)nx * nyw;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    int j = (int)(k / nx), i = (int)(k % nx);
    double lo, la;
    if (j < ny) {
      lo = in_1d ? Xin[i] : Xin[(size_t)i + (size_t)j * nx];
      la = in_1d ? Yin[j] : Yin[(size_t)i + (size_t)j * nx];
    } else {
      lo = in_1d ? Xin[i] : Xin[(size_t)i + (size_t)(ny - 1) * nx];
      la = (j == ny) ? 90.0 : 90.0 + dyp;
    }
    X[k] = lo;
    Y[k] = la;
    double ux, uy, uz;
    unit_vec(lo, la, ux, uy, uz);
    vx[k] = ux; vy[k] = uy; vz[k] = uz;
  }
}
__global__ void k_merc_table(const double* __restrict__ lat, size_t n, double* merc) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) merc[k] = d_merc(lat[k]);
}
// hE/hW per column and hN/hS per row of a separable source (see map_body): the arguments are exactly those the
// body would form -- lonN = lonS = lonP (same column), yb = ya for the E/W neighbours (same row)
__global__ void k_heading_tables(SrcGeom sg, int k_ew, double* hE, double* hW, double* hN, double* hS, double* lonm) {
  const int nx = sg.nx, nyw = sg.nyw;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nx + nyw; k += gridDim.x * blockDim.x) {
    if (k < nx) {
      const int iP = k + 1;
      int iPm1 = iP - 1, iPp1 = iP + 1;
      if (iPm1 == 0) { if (k_ew >= 0) iPm1 = nx - k_ew; }
      if (iPp1 == nx + 1) { if (k_ew >= 0) iPp1 = 1 + k_ew; }
      const double lonP = d_mod(sg.lon1[iP - 1], 360.0);
      double e = 0.0, w = 0.0;
      if (iPp1 >= 1 && iPp1 <= nx) e = d_heading_y(lonP, d_mod(sg.lon1[iPp1 - 1], 360.0), 0.0, 0.0);
      if (iPm1 >= 1 && iPm1 <= nx) w = d_heading_y(lonP, d_mod(sg.lon1[iPm1 - 1], 360.0), 0.0, 0.0);
      hE[k] = e; hW[k] = w; lonm[k] = lonP;
    } else {
      const int j = k - nx;  // 0-based row jP-1
      double n = 0.0, s_ = 0.0;
      if (j + 1 < nyw) n = d_heading_y(0.0, 0.0, sg.merc[j], sg.merc[j + 1]);
      if (j >= 1) s_ = d_heading_y(0.0, 0.0, sg.merc[j], sg.merc[j - 1]);
      hN[j] = n; hS[j] = s_;
    }
  }
}
__global__ void k_lat_table(const double* __restrict__ Yin, int ny, int nyw, double dyp, double* lat) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nyw; j += gridDim.x * blockDim.x)
    lat[j] = (j < ny) ? Yin[j] : ((j == ny) ? 90.0 : 90.0 + dyp);
}
// ji_m = MINLOC(ABS(XX(:,ny)-MOD(zlon+180.,360.)))  (src/mod_manip.f90:1821-1826)
__device__ __forceinline__ int minloc_abs_sorted(const double* __restrict__ v, int n, double r, double v0, double inv_d, bool well_spaced);
__global__ void k_jim_table(const double* __restrict__ lonrow, int nx, int sorted, int32_t* im) {
  for (int ji = blockIdx.x * blockDim.x + threadIdx.x; ji < nx; ji += gridDim.x * blockDim.x) {
    double zlon_m = fmod(lonrow[ji] + 180.0, 360.0);
    if (sorted) { im[ji] = minloc_abs_sorted(lonrow, nx, zlon_m, 0.0, 0.0, false); continue; }
    int best = 0;
    double bv = fabs(lonrow[0] - zlon_m);
    for (int i = 1; i < nx; i++) {
      double v = fabs(lonrow[i] - zlon_m);
      if (v < bv) { bv = v; best = i; }
    }
    im[ji] = best + 1;
  }
}
__global__ void k_check_sorted(const double* __restrict__ v, int n, int* unsorted) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += gridDim.x * blockDim.x)
    if (v[i] < v[i - 1]) atomicOr(unsorted, 1);
}
__global__ void k_gather_axes(SrcGeom sg, double* vlon, double* vlat) {  // vlon = Xsrc(:,3), vlat = Ysrc(3,:)
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < sg.nx + sg.nyw; k += gridDim.x * blockDim.x) {
    if (k < sg.nx) vlon[k] = src_lon(sg, k + 1, 3);
    else vlat[k - sg.nx] = src_lat(sg, 3, k - sg.nx + 1);
  }
}

// ---------------------------------------------------------------------------------------------------
// MINLOC(ABS(v(:)-r)) : first minimum, 1-based
// lower_bound (first index with v[idx] >= r) of a sorted axis.  On a uniform axis the index is guessed from
// (r - v0) * inv_d and VERIFIED against the two neighbouring entries (the definition of lower_bound), so the result
// is the bracket search's in 2-3 loads instead of log2(n) dependent ones; any miss falls back to the bisection.
__device__ __forceinline__ int lower_bound_axis(const double* __restrict__ v, int n, double r, double v0, double inv_d) {
  if (inv_d > 0.0) {
    const double t = (r - v0) * inv_d;
    if (t > -2.0 && t < (double)n + 2.0) {  // (false for NaN)
      const int g = (int)ceil(t);
#pragma unroll
      for (int d = 0; d < 3; d++) {
        const int lo = g + (d == 0 ? 0 : (d == 1 ? -1 : 1));
        if (lo < 0 || lo > n) continue;
        if ((lo == 0 || v[lo - 1] < r) && (lo == n || v[lo] >= r)) return lo;
      }
    }
  }
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (v[mid] < r) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// well_spaced: the axis is strictly increasing with steps >= 1e-9 and |v| <= 1e5 (is_sorted_host).  For |r| <= 1e6 every
// |v - r| is then formed with an error below 2.4e-10, so entries further from r than the chosen neighbour of the bracket
// are strictly further after rounding too: the walk back over equal distances (first minimum of MINLOC) cannot move, and
// its dependent load is skipped.
__device__ __forceinline__ int minloc_abs_sorted(const double* __restrict__ v, int n, double r, double v0 = 0.0,
                                                 double inv_d = 0.0, bool well_spaced = false) {
  const int lo = lower_bound_axis(v, n, r, v0, inv_d);
  int best;
  if (lo == 0) best = 0;
  else if (lo == n) best = n - 1;
  else best = (fabs(v[lo - 1] - r) <= fabs(v[lo] - r)) ? lo - 1 : lo;
  if (well_spaced && fabs(r) <= 1.0e6) return best + 1;
  double bv = fabs(v[best] - r);
  while (best > 0 && fabs(v[best - 1] - r) == bv) best--;
  return best + 1;
}
static __device__ __noinline__ int minloc_abs_linear(const double* __restrict__ v, int n, double r) {
  int best = 0;
  double bv = fabs(v[0] - r);
  for (int i = 1; i < n; i++) {
    double x = fabs(v[i] - r);
    if (x < bv) { bv = x; best = i; }
  }
  return best + 1;
}

// first-minimum scan of zr*ACOS(MIN(pds,1)) over the window (i1:i2, j1:j2), column-major order
struct Best {
  double d, pds;
  int i, j;
};
#define PDS_MARGIN 1.0e-13
#ifndef MAP_MIN_BLOCKS
#define MAP_MIN_BLOCKS 8
#endif
// (out of line, and the winner comes back as ONE 64-bit word -- j in the high half, i in the low half; the caller forms the
// dot product and the distance of that point with the very expressions of the loops below: a `Best&` argument, or a record
// returned by value, lives on the kernel's stack)
template <bool SEP>
static __device__ __noinline__ unsigned long long scan_window_exact_r(const SrcGeom& sg, double ux, double uy, double uz, int i1, int i2, int j1,
                                                         int j2, int ic, int jc) {
  Best b;
  // MINLOC = lexicographic minimum of (distance, column-major position).  The window centre (ic,jc) -- the
  // index-nearest point, almost always the winner or next to it -- is evaluated first, so that nearly every other
  // candidate is rejected by the dot-product margin without an acos; ties are still resolved by position, which
  // makes the result independent of the evaluation order.
  ic = min(max(ic, i1), i2);
  jc = min(max(jc, j1), j2);
  if (SEP) {  // callers guarantee i2 - i1 <= 8 (the EASY 9x9 window)
    // Regular source: u.v = clat_j*(ux*clon_i + uy*slon_i) + uz*slat_j up to ~4e-16 of rounding.  That cheap form
    // (2 flops per candidate) is only used to REJECT (with the margin widened by 1e-14, 20x its error); survivors are
    // re-evaluated in the reference's own operation order before any comparison that decides the result.
    // A whole row is rejected at once through the largest cheap form it can hold (rounding is monotone, so
    // cl*amax + c0 -- or cl*amin + c0 on the virtual row beyond the pole where cl < 0 -- IS that maximum).
    // Pass 0 is the centre alone (it seeds the running best); passes 1.. are the rows j1..j2.  One code copy of
    // the exact evaluation (acos) serves both: the kernel is instruction-cache bound.
    // (the column terms are recomputed where they are used: this path is rare, and an array indexed by the survivor loop
    // would live on the stack)
    double amax = -4.0, amin = 4.0;
    for (int q = 0; q < 9; q++) {
      int ji = min(i1 + q, i2);
      const double aq = ux * sg.clon[ji - 1] + uy * sg.slon[ji - 1];
      amax = fmax(amax, aq); amin = fmin(amin, aq);
    }
    b.d = 1.0e300; b.pds = -4.0; b.i = 0; b.j = 0;
    const int nq = i2 - i1 + 1;
    for (int pass = 0; pass <= j2 - j1 + 1; pass++) {
      const int jj = (pass == 0) ? jc : j1 + pass - 1;
      const double cl = sg.clat[jj - 1], sl = sg.slat[jj - 1];
      const double c0 = uz * sl;
      const double thr = b.pds - (PDS_MARGIN + 1.0e-14);
      unsigned surv = 0;
      if (pass == 0) surv = 1u << (ic - i1);
      else if (!(((cl >= 0.0) ? cl * amax : cl * amin) + c0 < thr)) {
        for (int q = 0; q < nq; q++) {
          const double aq = ux * sg.clon[i1 + q - 1] + uy * sg.slon[i1 + q - 1];
          if (!(cl * aq + c0 < thr)) surv |= 1u << q;
        }
        if (jj == jc) surv &= ~(1u << (ic - i1));  // the centre is already in the running best
      }
      while (surv) {
        const int q = __ffs(surv) - 1;
        surv &= surv - 1;
        const int ji = i1 + q;
        double zpds = ux * (sg.clon[ji - 1] * cl) + uy * (sg.slon[ji - 1] * cl) + uz * sl;
        if (zpds < b.pds - PDS_MARGIN) continue;
        double d = G_ZR * pm_acos(fmin(zpds, 1.0));
        bool earlier = (jj < b.j) || (jj == b.j && ji < b.i);
        if (d < b.d || (d == b.d && earlier)) { b.d = d; b.pds = zpds; b.i = ji; b.j = jj; }
      }
    }
    return ((unsigned long long)(unsigned)b.j << 32) | (unsigned)b.i;
  }
  b.d = 0.0; b.pds = -4.0; b.i = 0; b.j = 0;
  {
    double vx, vy, vz;
    src_vec(sg, ic, jc, vx, vy, vz);
    double zpds = ux * vx + uy * vy + uz * vz;
    b.d = G_ZR * pm_acos(fmin(zpds, 1.0));
    b.pds = zpds; b.i = ic; b.j = jc;
  }
  for (int jj = j1; jj <= j2; jj++) {
    for (int ji = i1; ji <= i2; ji++) {
      double vx, vy, vz;
      src_vec(sg, ji, jj, vx, vy, vz);
      double zpds = ux * vx + uy * vy + uz * vz;
      if (zpds < b.pds - PDS_MARGIN) continue;   // => strictly farther than the current best
      double d = G_ZR * pm_acos(fmin(zpds, 1.0));
      bool earlier = (jj < b.j) || (jj == b.j && ji < b.i);
      if (d < b.d || (d == b.d && earlier)) { b.d = d; b.pds = zpds; b.i = ji; b.j = jj; }
    }
  }
  return ((unsigned long long)(unsigned)b.j << 32) | (unsigned)b.i;
}
template <bool SEP>
__device__ __forceinline__ void scan_window_exact(const SrcGeom& sg, double ux, double uy, double uz, int i1, int i2, int j1, int j2,
                                                  Best& b, int ic, int jc) {
  const unsigned long long r = scan_window_exact_r<SEP>(sg, ux, uy, uz, i1, i2, j1, j2, ic, jc);
  b.i = (int)(unsigned)(r & 0xffffffffull); b.j = (int)(unsigned)(r >> 32);
  if (b.i < 1 || b.j < 1) { b.pds = -4.0; b.d = 1.0e300; return; }   // (NaN coordinates: no candidate ever won)
  if (SEP) {
    const double cl = sg.clat[b.j - 1], sl = sg.slat[b.j - 1];
    b.pds = ux * (sg.clon[b.i - 1] * cl) + uy * (sg.slon[b.i - 1] * cl) + uz * sl;
  } else {
    double vx, vy, vz;
    src_vec(sg, b.i, b.j, vx, vy, vz);
    b.pds = ux * vx + uy * vy + uz * vz;
  }
  b.d = G_ZR * pm_acos(fmin(b.pds, 1.0));
}

// The scan the kernel runs.  Regular source: the decision is taken on the dot products alone -- the candidate with the
// largest u.v wins outright when every other candidate is more than PDS_MARGIN below it (acos is monotone and a margin
// of 1e-13 in u.v is far above what its rounding can hide), and only ONE acos is evaluated per target, after the loop,
// by all lanes together.  (With the acos inside the loop, the lanes of a warp that met a better candidate at different
// iterations each made the whole warp execute the 18-term polynomial again: ~10 executions per warp.)  A second
// candidate within the margin of the best -- equidistant source points, the rows at and beyond the pole -- sends the
// target to the exact scan above, which evaluates and compares distances like the reference.
// acl: an upper bound of |ux*clon_i + uy*slon_i| over every column (|cos(lat_t)| plus a few ulp, unit_vec_cp): a row whose
//   best conceivable dot product |clat_j|*acl + uz*slat_j stays below the running best is skipped without touching its
//   columns (r01 kept the nine column terms a[0..8] in registers across the row loop for this: 18 registers, spills).
// unimodal: the longitudes of the window are sorted and within 90 degrees of the target's, so along a row with clat_j >= 0
//   the dot product decreases on both sides of the index-nearest column ic: the columns are visited outwards from ic and a
//   side is abandoned at the first column the cheap form rejects -- every column beyond it is smaller still in exact
//   arithmetic, and the 1e-14 slack of the rejection threshold (25x the rounding of either form) covers the evaluation.
//   Typically 3-4 column evaluations per surviving row instead of 9.
template <bool SEP>
__device__ __forceinline__ void scan_window(const SrcGeom& sg, double ux, double uy, double uz, int i1, int i2, int j1,
                                            int j2, Best& b, int ic, int jc, double acl = 1.0, bool unimodal = false, bool rows_unimodal = false) {
  if (!SEP) { scan_window_exact<SEP>(sg, ux, uy, uz, i1, i2, j1, j2, b, ic, jc); return; }
  ic = min(max(ic, i1), i2);
  jc = min(max(jc, j1), j2);
  double p1, p2 = -4.0;   // largest and second largest exact u.v met so far
  int bi = ic, bj = jc;
  {
    const double cl = sg.clat[jc - 1], sl = sg.slat[jc - 1];
    p1 = ux * (sg.clon[ic - 1] * cl) + uy * (sg.slon[ic - 1] * cl) + uz * sl;
  }
  // rows: with a sorted latitude axis the row bound |clat_j|*acl + uz*slat_j (the cosine of the target's distance to the
  // row's parallel) falls off on both sides of the index-nearest row jc, so the rows are visited outwards from jc as well
  // and a side is abandoned at the first row the bound rejects (rows_unimodal); otherwise all rows j1..j2 are tested.
  // Visit order: jc, jc+1, ..., j2 (or the first reject), then jc-1, ..., j1 (or the first reject) -- two plain loops (a single
  // loop with a direction variable spent a quarter of the scan's instructions on its own control flow).  Rows that are not
  // unimodal: one side, j1..j2, every row tested.
  for (int side = 0; side < 2; side++) {
    if (side && !rows_unimodal) break;
    const int jbeg = side ? jc - 1 : (rows_unimodal ? jc : j1), jend = side ? j1 : j2, jstep = side ? -1 : 1;
    for (int jrow = jbeg; side ? (jrow >= jend) : (jrow <= jend); jrow += jstep) {
      const double cl = sg.clat[jrow - 1], sl = sg.slat[jrow - 1];
      const double c0 = uz * sl;
      const double thr = p1 - (PDS_MARGIN + 1.0e-14);
      if (fabs(cl) * acl + c0 < thr) {
        if (!rows_unimodal) continue;
        break;   // this side is exhausted
      }
      auto cand = [&](int ji) -> bool {   // false: the cheap form rejects the column
        const double c = sg.clon[ji - 1], s = sg.slon[ji - 1];
        if (cl * (ux * c + uy * s) + c0 < thr) return false;
        const double zpds = ux * (c * cl) + uy * (s * cl) + uz * sl;
        if (zpds > p1) { p2 = p1; p1 = zpds; bi = ji; bj = jrow; }
        else p2 = fmax(p2, zpds);
        return true;
      };
      if (unimodal && cl >= 0.0) {
        if (jrow != jc) { if (!cand(ic)) continue; }   // (on row jc the centre seeded p1)
        for (int ji = ic + 1; ji <= i2; ji++) if (!cand(ji)) break;
        for (int ji = ic - 1; ji >= i1; ji--) if (!cand(ji)) break;
      } else {
        for (int ji = i1; ji <= i2; ji++)
          if (!(jrow == jc && ji == ic)) cand(ji);
      }
    }
  }
  if (p2 >= p1 - PDS_MARGIN) { scan_window_exact<SEP>(sg, ux, uy, uz, i1, i2, j1, j2, b, ic, jc); return; }
  b.pds = p1; b.i = bi; b.j = bj;
  b.d = G_ZR * pm_acos(fmin(p1, 1.0));
}

// ---------------------------------------------------------------------------------------------------
struct MapOut {
  int iq;
  double a, b;
  int ipb;
  float dnp;
};

// body of the target loop of MAPPING_BL, src/mod_bilin_2d.f90:459-634
// DISTANCE(xP, lonP, yP, latP) (:549), general form, kept out of line (code size: the kernel is instruction-cache bound).
// The unit vectors are pure functions of (lon, lat): the target's is reused when the search computed it for the same xP,
// the source table entry when MOD() left the longitude unchanged.
static __device__ __noinline__ float dnp_noinline(const SrcGeom& sg, double xP, double yP, double lonP, double latP, int iP, int jP,
                                           bool have_u, double ux0, double uy0, double uz0) {
  double ux, uy, uz, vx, vy, vz;
  if (have_u) { ux = ux0; uy = uy0; uz = uz0; } else unit_vec(xP, yP, ux, uy, uz);
  if (lonP == src_lon(sg, iP, jP)) src_vec(sg, iP, jP, vx, vy, vz); else unit_vec(lonP, latP, vx, vy, vz);
  double zpds = ux * vx + uy * vy + uz * vz;
  return (zpds >= 1.0) ? 0.f : (float)(G_ZR * pm_acos(zpds));
}

// first-guess quadrant of MAPPING_BL (:511-540), literal: five headings and the four interval tests
template <bool SEP>
__device__ __forceinline__ int first_guess_quadrant(const SrcGeom& sg, double xP, double yP, int iP, int jP, int iPm1, int iPp1,
                                                        int jPm1, int jPp1, double lonP, double lonN, double lonE, double lonS,
                                                        double lonW) {
  double ya = SEP ? sg.merc[jP - 1] : src_merc(sg, iP, jP);
  double hP = d_heading_y(lonP, xP, ya, d_merc(yP));
  double hN, hE, hS, hW;
  if (SEP) {
    // regular source: lonN == lonS == lonP and the E/W neighbours share the row's ordinate, so these four are
    // functions of jP (N, S) or iP (E, W) alone -- read from the tables k_heading_tables built with this very function
    hN = sg.hN[jP - 1]; hE = sg.hE[iP - 1]; hS = sg.hS[jP - 1]; hW = sg.hW[iP - 1];
  } else {
    hN = d_heading_y(lonP, lonN, ya, src_merc(sg, iP, jPp1));
    hE = d_heading_y(lonP, lonE, ya, src_merc(sg, iPp1, jP));
    hS = d_heading_y(lonP, lonS, ya, src_merc(sg, iP, jPm1));
    hW = d_heading_y(lonP, lonW, ya, src_merc(sg, iPm1, jP));
  }
  int iqdrn = 4;
  double hPp = d_degE_to_degWE(hP);
  if (hN > hE) hN = hN - 360.0;
  if (hPp > hN && hPp <= hE) iqdrn = 1;
  if (hP > hE && hP <= hS) iqdrn = 2;
  if (hP > hS && hP <= hW) iqdrn = 3;
  if (hP > hW && hPp <= hN) iqdrn = 4;
  return iqdrn;
}

// the same for a regular source, out of line: only targets within 1e-7 degrees of a grid line (or beyond 89 degrees) get here
static __device__ __noinline__ int first_guess_quadrant_rare(const SrcGeom& sg, double xP, double yP, int iP, int jP, int iPm1,
                                                             int iPp1, int jPm1, int jPp1, double lonP, double lonN, double lonE,
                                                             double lonS, double lonW) {
  return first_guess_quadrant<true>(sg, xP, yP, iP, jP, iPm1, iPp1, jPm1, jPp1, lonP, lonN, lonE, lonS, lonW);
}

template <bool SEP>
__device__ __forceinline__ void map_body(const SrcGeom& sg, int k_ew, double xP, double yP, int iP, int jP, MapOut& o,
                                         int* err, bool have_u = false, double ux0 = 0.0, double uy0 = 0.0, double uz0 = 0.0,
                                         double best_pds = 0.0, double best_d = 0.0) {
  o.iq = 0; o.a = 0.0; o.b = 0.0; o.ipb = 0; o.dnp = 0.f;
  const int nxi = sg.nx, nyi = sg.nyw;
  if (!((iP != (int)G_RMV) && (jP != (int)G_RMV) && (jP < nyi))) return;
  int iPm1 = iP - 1, iPp1 = iP + 1, jPm1 = jP - 1, jPp1 = jP + 1;
  if (iPm1 == 0) { if (k_ew >= 0) iPm1 = nxi - k_ew; }
  if (iPp1 == nxi + 1) { if (k_ew >= 0) iPp1 = 1 + k_ew; }
  if ((iPm1 < 1) || (jPm1 < 1) || (iPp1 > nxi)) return;
  double lonP, latP, lonN, latN, lonE, latE, lonS, latS, lonW, latW;
  if (SEP) {  // regular source: three longitudes (already MOD 360: sg.lonm) and three latitudes describe the whole cross
    lonP = sg.lonm[iP - 1]; lonE = sg.lonm[iPp1 - 1]; lonW = sg.lonm[iPm1 - 1];
    latP = sg.lat1[jP - 1]; latN = sg.lat1[jPp1 - 1]; latS = sg.lat1[jPm1 - 1];
    lonN = lonP; lonS = lonP; latE = latP; latW = latP;
  } else {
    lonP = d_mod(src_lon(sg, iP, jP), 360.0); latP = src_lat(sg, iP, jP);
    lonN = d_mod(src_lon(sg, iP, jPp1), 360.0); latN = src_lat(sg, iP, jPp1);
    lonE = d_mod(src_lon(sg, iPp1, jP), 360.0); latE = src_lat(sg, iPp1, jP);
    lonS = d_mod(src_lon(sg, iP, jPm1), 360.0); latS = src_lat(sg, iP, jPm1);
    lonW = d_mod(src_lon(sg, iPm1, jP), 360.0); latW = src_lat(sg, iPm1, jP);
  }
  xP = d_mod(xP, 360.0);
  int iqdrn = 4;
  bool quad_known = false;
  if (SEP) {
    // Regular source: the neighbour headings are table entries (functions of jP or iP alone, see k_heading_tables) and on
    // a grid whose rows / columns are parallels / meridians they are exactly 0, 90, 180, 270.  The first-guess quadrant
    // (:531-540) is then decided by the signs of the two arguments of heading()'s ATAN2 -- the wrapped longitude
    // difference, formed here exactly as heading() forms it, and the difference of the Mercator ordinates, whose sign is
    // that of yP - latP (the ordinate is increasing with slope >= 1 and both values carry < 1e-13 of rounding error below
    // 89 degrees; a margin of 1e-7 degrees is 1.7e-9).  With both arguments at least 2e-10 of each other the angle stays
    // > 1e-10 rad inside its quarter, far above the rounding of ATAN2 and of the conversion to degrees: the comparisons
    // of the reference cannot come out differently.  Anything closer to an axis takes the literal path below.
    const double hN = sg.hN[jP - 1], hE = sg.hE[iP - 1], hS = sg.hS[jP - 1], hW = sg.hW[iP - 1];
    if (hN == 0.0 && hE == 90.0 && hS == 180.0 && hW == 270.0) {
      double dx = d_mod((xP * G_ZCONV - lonP * G_ZCONV), 2 * G_PI);
      if (dx >= G_PI) dx = dx - 2 * G_PI;
      if (dx <= -G_PI) dx = dx + 2 * G_PI;
      const double dphi = yP - latP;
      const double adx = fabs(dx), adp = fabs(dphi);
      if (adx >= 2.0e-9 && adx <= 0.8 && adp >= 1.0e-7 && adp <= 45.0 && fabs(yP) <= 89.0 && fabs(latP) <= 89.0) {
        quad_known = true;
        iqdrn = (dx > 0.0) ? ((dphi > 0.0) ? 1 : 2) : ((dphi > 0.0) ? 4 : 3);
      }
    }
  }
  if (!quad_known) {
    if (SEP) iqdrn = first_guess_quadrant_rare(sg, xP, yP, iP, jP, iPm1, iPp1, jPm1, jPp1, lonP, lonN, lonE, lonS, lonW);
    else iqdrn = first_guess_quadrant<false>(sg, xP, yP, iP, jP, iPm1, iPp1, jPm1, jPp1, lonP, lonN, lonE, lonS, lonW);
  }
  double loni[5], lati[5];
  loni[0] = xP; lati[0] = yP;
  loni[1] = lonP; lati[1] = latP;
  {
    // DISTANCE(xP, lonP, yP, latP) (:549).  The unit vectors are pure functions of (lon, lat): reuse the target's
    // (computed by the search for the same xP) and the source table entry when MOD() left the longitudes unchanged.
    if (SEP && have_u && (sg.lonm_same || lonP == sg.lon1[iP - 1])) {
      // the search's winning dot product was formed from these very unit vectors in this very order, and its
      // distance is zr*ACOS(MIN(pds,1)): DISTANCE differs only by returning 0 when pds >= 1
      o.dnp = (best_pds >= 1.0) ? 0.f : (float)best_d;
    } else {
      o.dnp = dnp_noinline(sg, xP, yP, lonP, latP, iP, jP, have_u, ux0, uy0, uz0);
    }
  }
  if (SEP) {
    // regular source: the quad of quadrant q is an axis-parallel cell with corners (lonP | lonE/W) x (latP | latN/S); the
    // "WHERE(loni <= 0) loni + 360" of :577 depends on the value alone and is applied to the three longitudes once
    const double xa = (lonP <= 0.0) ? lonP + 360.0 : lonP;
    const double xe = (lonE <= 0.0) ? lonE + 360.0 : lonE;
    const double xw = (lonW <= 0.0) ? lonW + 360.0 : lonW;
    int icpt = 0, qbuilt = iqdrn;
    bool lagain = true;
    const int iqdrn0 = 0;  // Q2
    while (lagain) {
      qbuilt = iqdrn;      // 1..4 here: the first guess, then icpt = 1..4
      const double xb = (iqdrn <= 2) ? xe : xw;
      const double yb = (iqdrn == 1 || iqdrn == 4) ? latN : latS;
      int ok = d_l_inpoly_rect(xa, xb, latP, yb, (iqdrn & 1) != 0, xP, yP);  // Q8: the un-moved xP
      if (ok < 0) { atomicOr(err, 1); ok = 0; }
      if ((!ok) && (yP < 88.0)) { icpt = icpt + 1; iqdrn = icpt; }
      else lagain = false;
      if (icpt == 5) { lagain = false; iqdrn = iqdrn0; }
    }
    // the arrays LOCAL_COORD sees are those of the last quad built
    const double xb = (qbuilt <= 2) ? xe : xw;
    const double yb = (qbuilt == 1 || qbuilt == 4) ? latN : latS;
    const bool pa = (qbuilt & 1) != 0;
    if (loni[0] <= 0.0) loni[0] = loni[0] + 360.0;
    loni[1] = xa; lati[1] = latP;
    loni[2] = pa ? xb : xa; lati[2] = pa ? latP : yb;
    loni[3] = xb; lati[3] = yb;
    loni[4] = pa ? xa : xb; lati[4] = pa ? yb : latP;
  } else {
    int icpt = 0;
    bool lagain = true;
    const int iqdrn0 = 0;  // Q2
    while (lagain) {
      switch (iqdrn) {
        case 1:
          loni[2] = lonE; lati[2] = latE;
          loni[3] = d_mod(src_lon(sg, iPp1, jPp1), 360.0); lati[3] = src_lat(sg, iPp1, jPp1);
          loni[4] = lonN; lati[4] = latN;
          break;
        case 2:
          loni[2] = lonS; lati[2] = latS;
          loni[3] = d_mod(src_lon(sg, iPp1, jPm1), 360.0); lati[3] = src_lat(sg, iPp1, jPm1);
          loni[4] = lonE; lati[4] = latE;
          break;
        case 3:
          loni[2] = lonW; lati[2] = latW;
          loni[3] = d_mod(src_lon(sg, iPm1, jPm1), 360.0); lati[3] = src_lat(sg, iPm1, jPm1);
          loni[4] = lonS; lati[4] = latS;
          break;
        case 4:
          loni[2] = lonN; lati[2] = latN;
          loni[3] = d_mod(src_lon(sg, iPm1, jPp1), 360.0); lati[3] = src_lat(sg, iPm1, jPp1);
          loni[4] = lonW; lati[4] = latW;
          break;
        default: break;
      }
#pragma unroll
      for (int k = 0; k < 5; k++) if (loni[k] <= 0.0) loni[k] = loni[k] + 360.0;
      int ok = d_l_inpoly(&loni[1], &lati[1], xP, yP);  // Q8: the un-moved xP
      if (ok < 0) { atomicOr(err, 1); ok = 0; }
      if ((!ok) && (yP < 88.0)) { icpt = icpt + 1; iqdrn = icpt; }
      else lagain = false;
      if (icpt == 5) { lagain = false; iqdrn = iqdrn0; }
    }
  }
  d_local_coord(loni, lati, o.a, o.b, o.ipb);
  o.iq = iqdrn;
}

// whole-array fix-ups of MAPPING_BL, :642-683, per element, in the reference's order (Q25)
__device__ __forceinline__ void map_store(const MapOut& o, int iP, int jP, int m, size_t k, size_t nt, int32_t* MT,
                                          double* AB, int16_t* IDP, float* DNP) {
  double a = o.a, b = o.b;
  int iq = o.iq;
  int id = o.ipb;
  if ((iP == (int)G_RMV) || (jP == (int)G_RMV)) { a = G_RMV; b = G_RMV; }
  if ((a < 0.0) && (a > -G_REPS)) a = 0.0;
  if ((b < 0.0) && (b > -G_REPS)) b = 0.0;
  if ((a > G_RMV) && (a < 0.0)) { a = 0.5; id = 4; }
  if (a > 1.0) { a = 0.5; id = 5; }
  if ((b > G_RMV) && (b < 0.0)) { b = 0.5; id = 6; }
  if (b > 1.0) { b = 0.5; id = 7; }
  if (iq < 1) { iq = 1; id = 1; }
  if (m <= -1) id = -1;
  if (m == 0) id = -2;
  if (m < -2) id = -3;
  MT[k] = iP; MT[k + nt] = jP; MT[k + 2 * nt] = iq;
  AB[k] = a; AB[k + nt] = b;
  IDP[k] = (int16_t)id;
  DNP[k] = o.dnp;
}

// ---------------------------------------------------------------------------------------------------
// EASY search (src/mod_manip.f90:967-1069) fused with the MAPPING_BL body: one thread per target.
// (__grid_constant__: the out-of-line rare paths take the geometry by reference; without it every thread copies the 280-byte
// parameter struct to its stack first)
template <bool SEP>
__global__ void __launch_bounds__(128, MAP_MIN_BLOCKS)
k_map_easy(const __grid_constant__ SrcGeom sg, const __grid_constant__ TrgGeom tg, int k_ew, const int32_t* __restrict__ mask, const double* __restrict__ vlon,
           const double* __restrict__ vlat, int sorted_lon, int sorted_lat, int jlo, int jhi, int32_t* MT, double* AB,
           int16_t* IDP, float* DNP, int* err, size_t k0, size_t kcnt, int search_only, const DevPlan* __restrict__ plan, FastDiv dnx) {
  const size_t nt = (size_t)tg.nx * tg.ny;
  const bool k32 = (k0 + kcnt) <= 0xffffffffull;   // the row / column split of k by a multiply-high (block-uniform)
  if (plan) { jlo = plan->jlo; jhi = plan->jhi; }   // the searched rows come from the device-side prelude (sosie_bilin_map_async)
  for (size_t kk = blockIdx.x * (size_t)blockDim.x + threadIdx.x; kk < kcnt; kk += (size_t)gridDim.x * blockDim.x) {
    const size_t k = k0 + kk;
    int jj_t, ji_t;
    if (k32) { const unsigned jq = fdiv((unsigned)k, dnx); jj_t = (int)jq + 1; ji_t = (int)((unsigned)k - jq * (unsigned)tg.nx) + 1; }
    else { jj_t = (int)(k / tg.nx) + 1; ji_t = (int)(k % tg.nx) + 1; }
    int m = mask ? mask[k] : 1;
    int iP = -1, jP = -1;
    double rlon = trg_lon(tg, ji_t, jj_t), rlat = trg_lat(tg, ji_t, jj_t);
    // the coordinates of this thread's NEXT target head a ~1800-instruction dependent chain: ask for their lines now (-1 %)
    if (!tg.is_1d) {
      const size_t kn = k + (size_t)gridDim.x * blockDim.x;
      if (kn < k0 + kcnt) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(tg.X + kn));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(tg.Y + kn));
      }
    }
    double ux = 0.0, uy = 0.0, uz = 0.0, bpds = 0.0, bd = 0.0;
    bool have_u = false;
    if (m == 1 && jj_t >= jlo && jj_t <= jhi) {
      int ip = sorted_lon ? minloc_abs_sorted(vlon, sg.nx, rlon, sg.lon0, sg.inv_dlon, sorted_lon == 2) : minloc_abs_linear(vlon, sg.nx, rlon);
      int jp = sorted_lat ? minloc_abs_sorted(vlat, sg.nyw, rlat, sg.lat0, sg.inv_dlat, sorted_lat == 2) : minloc_abs_linear(vlat, sg.nyw, rlat);
      int i1s = max(ip - 4, 1), i2s = min(ip + 4, sg.nx);
      int j1s = max(jp - 4, 1), j2s = min(jp + 4, sg.nyw);
      double cpt;
      unit_vec_cp(rlon, rlat, ux, uy, uz, cpt);
      have_u = (fabs(rlon) < 360.0);   // MOD(xP,360) == xP: the body's DISTANCE sees the same longitude
      Best b;
      const bool unimodal = SEP && sorted_lon && fabs(vlon[i1s - 1] - rlon) <= 90.0 && fabs(vlon[i2s - 1] - rlon) <= 90.0;
      scan_window<SEP>(sg, ux, uy, uz, i1s, i2s, j1s, j2s, b, ip, jp, fabs(cpt) * (1.0 + 4.0e-15), unimodal, SEP && sorted_lat != 0);
      iP = b.i; jP = b.j; bpds = b.pds; bd = b.d;
      if (iP == 0 || jP == 0) atomicOr(err, 2);
    }
    MapOut o;
    o.iq = 0; o.a = 0.0; o.b = 0.0; o.ipb = 0; o.dnp = 0.f;
    if (search_only) { MT[k] = iP; MT[k + nt] = jP; continue; }   // FIND_NEAREST_POINT alone: JIp, JJp (-1 where not searched)
    if (m == 1) map_body<SEP>(sg, k_ew, rlon, rlat, iP, jP, o, err, have_u, ux, uy, uz, bpds, bd);
    map_store(o, iP, jP, m, k, nt, MT, AB, IDP, DNP);
  }
}

// body only (after TWISTED): JI/JJ given
__global__ void __launch_bounds__(128)
k_map_body(const __grid_constant__ SrcGeom sg, const __grid_constant__ TrgGeom tg, int k_ew, const int32_t* __restrict__ mask, const int32_t* __restrict__ JI,
           const int32_t* __restrict__ JJ, int32_t* MT, double* AB, int16_t* IDP, float* DNP, int* err) {
  const size_t nt = (size_t)tg.nx * tg.ny;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < nt; k += (size_t)gridDim.x * blockDim.x) {
    int jj_t = (int)(k / tg.nx) + 1, ji_t = (int)(k % tg.nx) + 1;
    int m = mask[k];
    int iP = JI[k], jP = JJ[k];
    MapOut o;
    o.iq = 0; o.a = 0.0; o.b = 0.0; o.ipb = 0; o.dnp = 0.f;
    if (m == 1) map_body<false>(sg, k_ew, trg_lon(tg, ji_t, jj_t), trg_lat(tg, ji_t, jj_t), iP, jP, o, err);
    map_store(o, iP, jP, m, k, nt, MT, AB, IDP, DNP);
  }
}

// ---------------------------------------------------------------------------------------------------
// TWISTED pieces
// e1_s/e2_s (src/mod_manip.f90:1152-1164)
__global__ void k_src_metrics(SrcGeom sg, double* e1, double* e2) {
  size_t n = (size_t)sg.nx * sg.nyw;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    int j = (int)(k / sg.nx) + 1, i = (int)(k % sg.nx) + 1;
    double a = (double)40000.f, b = (double)40000.f;
    int ie = (i <= sg.nx - 1) ? i : sg.nx - 1;  // e1_s(nx_s,:) = e1_s(nx_s-1,:)
    int je = (j <= sg.nyw - 1) ? j : sg.nyw - 1;
    if (sg.nx > 1) a = d_distance(src_lon(sg, ie, j), src_lon(sg, ie + 1, j), src_lat(sg, ie, j), src_lat(sg, ie + 1, j)) * 1000.0;
    if (sg.nyw > 1) b = d_distance(src_lon(sg, i, je), src_lon(sg, i, je + 1), src_lat(sg, i, je), src_lat(sg, i, je + 1)) * 1000.0;
    e1[k] = a;
    e2[k] = b;
  }
}

// argmin |Y - 90| with first-index tie-break (IS_POLAR_PROJ, src/mod_manip.f90:1862)
__global__ void k_argmin_np(const double* __restrict__ Y, size_t n, int is1d, int nx, MinIdx* partial) {
  MinIdx m;
  m.v = 0; m.k = -1;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    double y = is1d ? Y[k / nx] : Y[k];
    MinIdx c;
    c.v = fabs(y - 90.0); c.k = (long long)k;
    m = minidx_better(m, c);
  }
  m = warp_minidx(m);
  __shared__ MinIdx sh[32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    MinIdx r = sh[0];
    for (int w = 1; w < (blockDim.x >> 5); w++) r = minidx_better(r, sh[w]);
    partial[blockIdx.x] = r;
  }
}

// J_VLAT_S (src/mod_manip.f90:1263-1286): thread per (column, bin)
__global__ void k_jvlat(SrcGeom sg, int nsplit, const double* __restrict__ VB, int* jmaxb, int* jminb) {
  int total = sg.nx * nsplit;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    int b = t / sg.nx, i = t % sg.nx + 1;
    double r1 = VB[b], r2 = VB[b + 1];
    int jx = 0, jn = 0;
    double vmax = 0, vmin = 0;
    for (int j = 1; j <= sg.nyw; j++) {
      double v = src_lat(sg, i, j);
      if (v <= r2) { if (jx == 0 || v > vmax) { vmax = v; jx = j; } }
      if (v > r1)  { if (jn == 0 || v < vmin) { vmin = v; jn = j; } }
    }
    atomicMax(&jmaxb[b], jx);
    if (jn > 0) atomicMin(&jminb[b], jn);
  }
}

// flags of targets to search, in processing order p = (row_k, i); order[] = compacted target indices
__global__ void k_twisted_flags(TrgGeom tg, const int32_t* __restrict__ mask, int j_strt, int icr, int nrows, uint8_t* flags,
                                int64_t* lin) {
  size_t n = (size_t)nrows * tg.nx;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
    int rk = (int)(p / tg.nx), i = (int)(p % tg.nx);
    int jj = j_strt + rk * icr;
    size_t k = (size_t)i + (size_t)(jj - 1) * tg.nx;
    flags[p] = (mask[k] == 1) ? 1 : 0;
    lin[p] = (int64_t)k;
  }
}

struct TwParams {
  int lStupido, nsplit;
  const double* VB;   // [nsplit+1]
  const int* JV;      // [2*nsplit] column-major (nsplit,2)
  const double* e1;
  const double* e2;
  double frac_emax, sqrt2f;
  const double* seg;  // [nty * ntx * 5]: centre (3), cos and sin of the angular radius of each 8x4 tile of the source grid
  int nseg;           // ntx: tiles per tile row
};

// chain-independent band search (niter = 0,1,...), one warp per target (:1350-1426)
// Spherical cap around each tile of 8 columns x 4 rows of the source grid: centre = normalised mean of the unit vectors,
// radius = largest angle to a member (its cosine lowered by 1e-12, its sine raised: the cap can only be too large).  One
// warp per tile, one lane per cell (cells beyond the grid repeat the tile's first cell).  A degenerate mean (points spread
// around a pole) gets the whole sphere.  Tile t = tj * ntx + ti covers columns 8*ti .. 8*ti+7, rows 4*tj .. 4*tj+3 (0-based).
#define TW_TX 8
#define TW_TY 4
__global__ void k_seg_table(SrcGeom sg, int ntx, int nty, double* seg) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5, total = ntx * nty;
  for (int t = warp; t < total; t += nwarps) {
    const int tj = t / ntx, ti = t % ntx;
    const int ji = ti * TW_TX + (lane & (TW_TX - 1)), jj = tj * TW_TY + (lane >> 3);
    const bool in = (ji < sg.nx) && (jj < sg.nyw);
    const size_t k = in ? (size_t)jj * sg.nx + ji : (size_t)(tj * TW_TY) * sg.nx + ti * TW_TX;
    const double vx = sg.vx[k], vy = sg.vy[k], vz = sg.vz[k];
    double sx = in ? vx : 0.0, sy = in ? vy : 0.0, sz = in ? vz : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    const double nrm = sqrt(sx * sx + sy * sy + sz * sz);
    double cx = 0.0, cy = 0.0, cz = 1.0, cr = -1.0;
    if (nrm > 1.0e-3) {
      cx = sx / nrm; cy = sy / nrm; cz = sz / nrm;
      double d = cx * vx + cy * vy + cz * vz;
      for (int o = 16; o > 0; o >>= 1) d = fmin(d, __shfl_xor_sync(0xffffffffu, d, o));
      cr = fmax(d - 1.0e-12, -1.0);
    }
    if (!(cr == cr)) cr = -1.0;   // NaN coordinates: never prune
    if (lane == 0) {
      double* o = seg + (size_t)t * 5;
      o[0] = cx; o[1] = cy; o[2] = cz; o[3] = cr; o[4] = fmin(sqrt(fmax(0.0, 1.0 - cr * cr)) + 1.0e-12, 1.0);
    }
  }
}
// upper bound of u.v over the cap (slack 1e-12 for the rounding of this very expression)
__device__ __forceinline__ double seg_bound(const double* __restrict__ e, double ux, double uy, double uz) {
  const double a = ux * e[0] + uy * e[1] + uz * e[2];
  if (!(a < e[3]) || a < -0.99) return 2.0;   // inside the cap (or NaN): anything is possible; near the antipode sqrt(1-a*a) loses its digits
  return a * e[3] + sqrt(fmax(0.0, 1.0 - a * a)) * e[4] + 1.0e-12;
}

__global__ void __launch_bounds__(256)
k_twisted_band(SrcGeom sg, TrgGeom tg, TwParams tp, const int64_t* __restrict__ order, int T, int2* band_pos,
               int8_t* band_found, int* err) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int q = warp; q < T; q += nwarps) {
    size_t k = (size_t)order[q];
    int jj_t = (int)(k / tg.nx) + 1, ji_t = (int)(k % tg.nx) + 1;
    double rlon = trg_lon(tg, ji_t, jj_t), rlat = trg_lat(tg, ji_t, jj_t);
    double ux, uy, uz;
    unit_vec(rlon, rlat, ux, uy, uz);
    int jlat = 0;
    if (!tp.lStupido) {
      for (jlat = 1; jlat <= tp.nsplit; jlat++) {
        if (rlat == tp.VB[jlat - 1]) break;
        if ((rlat > tp.VB[jlat - 1]) && (rlat <= tp.VB[jlat])) break;
      }
    }
    int niter = 0, found = 0, bi = 0, bj = 0;
    while (true) {
      int j1s, j2s;
      if (tp.lStupido) { j1s = 1; j2s = sg.nyw; }
      else {
        int a1 = max(jlat - niter, 1), a2 = min(jlat + niter, tp.nsplit);
        j1s = tp.JV[(a1 - 1)];                    // (a1,1); a1 == nsplit+1 aliases (1,2)  (Q4)
        j2s = tp.JV[(a2 - 1) + tp.nsplit];        // (a2,2)
      }
      long long ncand = (j2s >= j1s) ? (long long)sg.nx * (j2s - j1s + 1) : 0;
      double bd = 0.0, bp = -4.0;
      long long bk = -1;
      // Pass 1 decides on the dot products alone (rows outer, lane-strided columns inner: no 64-bit division, no acos in the
      // loop): every lane keeps its largest and second largest u.v; if, over the whole warp, the runner-up lies more than
      // PDS_MARGIN below the winner, the winner is the unique nearest point and ONE acos gives its distance.  Otherwise
      // (equidistant candidates) pass 2 compares distances exactly like the reference: first minimum in flattened order.
      bool exact = sg.separable || ncand <= 0;
      if (!exact) {
        double p1 = -4.0, p2 = -4.0;
        long long k1 = -1;
        // one 8x4 tile: a candidate per lane (rows outside the band and cells beyond the grid are skipped)
        const int tj0 = (j1s - 1) / TW_TY, tj1 = (j2s - 1) / TW_TY;             // tile rows the band touches
        auto scan_seg = [&](int sid) {
          const int tj = sid / tp.nseg + tj0, ti = sid % tp.nseg;
          const int ji = ti * TW_TX + (lane & (TW_TX - 1)), jj = tj * TW_TY + (lane >> 3) + 1;   // 0-based column, 1-based row
          if (ji < sg.nx && jj >= j1s && jj <= j2s) {
            const size_t ro = (size_t)(jj - 1) * sg.nx;
            const double zpds = ux * __ldg(sg.vx + ro + ji) + uy * __ldg(sg.vy + ro + ji) + uz * __ldg(sg.vz + ro + ji);
            const long long kc = (long long)(jj - j1s) * sg.nx + ji;
            // several tiles feed one lane in no particular order: the first index among equal dot products is kept
            // explicitly (an exact tie makes the target ambiguous anyway: p2 == p1)
            if (zpds > p1) { p2 = p1; p1 = zpds; k1 = kc; }
            else { p2 = fmax(p2, zpds); if (zpds == p1 && kc < k1) k1 = kc; }
          }
        };
        const int nsb = (tj1 - tj0 + 1) * tp.nseg;                       // tiles of the band, row-major
        const double* segb = tp.seg ? tp.seg + (size_t)tj0 * tp.nseg * 5 : nullptr;
        if (tp.seg) {
          // (a) the segment whose cap comes closest to the target seeds the running best
          double bb = -8.0;
          int bs = 0;
          for (int sid = lane; sid < nsb; sid += 32) {
            const double b = seg_bound(segb + (size_t)sid * 5, ux, uy, uz);
            if (b > bb) { bb = b; bs = sid; }
          }
          for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, bb, o);
            const int os = __shfl_xor_sync(0xffffffffu, bs, o);
            if (ob > bb || (ob == bb && os < bs)) { bb = ob; bs = os; }
          }
          scan_seg(bs);
          double best = p1;
          for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
          // (b) every other segment whose cap can hold a point within the margin of the best is scanned; the rest cannot
          //     contain the nearest point nor tie with it
          for (int base = 0; base < nsb; base += 32) {
            const int sid = base + lane;
            const bool want = (sid < nsb) && (sid != bs) && !(seg_bound(segb + (size_t)sid * 5, ux, uy, uz) < best - (PDS_MARGIN + 1.0e-12));
            unsigned todo = __ballot_sync(0xffffffffu, want);
            const bool any = todo != 0;
            while (todo) {
              const int b = __ffs(todo) - 1;
              todo &= todo - 1;
              scan_seg(base + b);
            }
            if (any) {   // tighten the threshold once per batch of 32 segments (the seed is almost always the winner already)
              double nb = p1;
              for (int o = 16; o > 0; o >>= 1) nb = fmax(nb, __shfl_xor_sync(0xffffffffu, nb, o));
              best = nb;
            }
          }
        } else {
          for (int sid = 0; sid < nsb; sid++) scan_seg(sid);   // no table: every tile
        }
        double g1 = p1;
        long long gk = k1;
        for (int o = 16; o > 0; o >>= 1) {
          const double op = __shfl_xor_sync(0xffffffffu, g1, o);
          const long long ok = __shfl_xor_sync(0xffffffffu, gk, o);
          if (ok >= 0 && (gk < 0 || op > g1 || (op == g1 && ok < gk))) { g1 = op; gk = ok; }
        }
        double m2 = (k1 == gk) ? p2 : p1;
        for (int o = 16; o > 0; o >>= 1) m2 = fmax(m2, __shfl_xor_sync(0xffffffffu, m2, o));
        if (gk >= 0 && !(m2 >= g1 - PDS_MARGIN) && !(g1 != g1)) { bd = G_ZR * pm_acos(fmin(g1, 1.0)); bp = g1; bk = (lane == 0) ? gk : -1; }
        else exact = true;
      }
      if (exact && ncand > 0) {
        bd = 0.0; bp = -4.0; bk = -1;
        for (long long c = lane; c < ncand; c += 32) {
          int ji = (int)(c % sg.nx) + 1, jj = (int)(c / sg.nx) + j1s;
          double vx, vy, vz;
          src_vec(sg, ji, jj, vx, vy, vz);
          double zpds = ux * vx + uy * vy + uz * vz;
          if (bk >= 0 && zpds < bp - PDS_MARGIN) continue;
          double d = G_ZR * pm_acos(fmin(zpds, 1.0));
          if (bk < 0 || d < bd) { bd = d; bp = zpds; bk = c; }
        }
      }
      MinIdx m;
      m.v = bd; m.k = bk;
      m = warp_minidx(m);
      if (m.k < 0) { if (lane == 0) atomicOr(err, 4); break; }
      bi = (int)(m.k % sg.nx) + 1;
      bj = (int)(m.k / sg.nx) + j1s;
      size_t ks = (size_t)(bi - 1) + (size_t)(bj - 1) * sg.nx;
      double emax = fmax(tp.e1[ks], tp.e2[ks]) / 1000.0 * tp.sqrt2f;
      if (m.v <= tp.frac_emax * emax) { found = 1; break; }
      if (niter == 0) {
        double lo1 = fmax(rlon - 0.5, 0.0), hi1 = fmin(rlon + 0.5, 360.0);
        double lo2 = fmax(rlat - 0.5, -90.0), hi2 = fmin(rlat + 0.5, 90.0);
        int any = 0;
        size_t ns = (size_t)sg.nx * sg.nyw;
        for (size_t c = lane; c < ns && !any; c += 32) {
          int ji = (int)(c % sg.nx) + 1, jj = (int)(c / sg.nx) + 1;
          double xs = src_lon(sg, ji, jj), ys = src_lat(sg, ji, jj);
          if ((xs > lo1) && (xs < hi1) && (ys > lo2) && (ys < hi2)) any = 1;
        }
        any = __any_sync(0xffffffffu, any);
        if (!any) break;
      }
      if (niter > tp.nsplit / 3) break;
      niter++;
    }
    if (lane == 0) { band_pos[q] = make_int2(bi, bj); band_found[q] = (int8_t)found; }
  }
}

// one pass of the chain: S_new[q] = f_q(S_old[q-1])   (:1336-1349, 1385-1391)
__global__ void __launch_bounds__(128)
k_twisted_chain(const __grid_constant__ SrcGeom sg, const __grid_constant__ TrgGeom tg, const __grid_constant__ TwParams tp, const int64_t* __restrict__ order, int T,
                const int2* __restrict__ band_pos, const int8_t* __restrict__ band_found,
                const int2* __restrict__ S_old, int2* S_new, int8_t* found, int* changed, const int2* __restrict__ S_prev) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < T; q += gridDim.x * blockDim.x) {
    // S_prev (a third buffer, not written by this pass) is the state of two passes ago.  A target whose predecessor did
    // not move between that pass and the last one would recompute exactly what it produced last time (S_old[q],
    // found[q]): copy it.  After the first couple of passes only the few targets downstream of a change do any work.
    if (S_prev && q > 0) {
      const int2 pp = S_prev[q - 1], pl = S_old[q - 1];
      if (pp.x == pl.x && pp.y == pl.y) { S_new[q] = S_old[q]; continue; }
    }
    size_t k = (size_t)order[q];
    int jj_t = (int)(k / tg.nx) + 1, ji_t = (int)(k % tg.nx) + 1;
    double rlat = trg_lat(tg, ji_t, jj_t);
    int2 res = band_pos[q];
    int fnd = band_found[q];
    if (!(rlat > 60.0)) {
      int2 pr = (q == 0) ? make_int2(0, 0) : S_old[q - 1];  // Q3: (0,0) before the first target
      double rlon = trg_lon(tg, ji_t, jj_t);
      double ux, uy, uz;
      unit_vec(rlon, rlat, ux, uy, uz);
      int i1s = max(pr.x - 5, 1), i2s = min(pr.x + 5, sg.nx);
      int j1s = max(pr.y - 5, 1), j2s = min(pr.y + 5, sg.nyw);
      Best b;
      scan_window<false>(sg, ux, uy, uz, i1s, i2s, j1s, j2s, b, pr.x, pr.y);
      if (b.i != 0) {
        size_t ks = (size_t)(b.i - 1) + (size_t)(b.j - 1) * sg.nx;
        double emax = fmax(tp.e1[ks], tp.e2[ks]) / 1000.0 * tp.sqrt2f;
        if (b.d <= tp.frac_emax * emax) { res = make_int2(b.i, b.j); fnd = 1; }
      }
    }
    int2 old = S_old[q];
    if (old.x != res.x || old.y != res.y) *changed = 1;
    S_new[q] = res;
    found[q] = (int8_t)fnd;
  }
}

__global__ void k_twisted_scatter(const int64_t* __restrict__ order, int T, const int2* __restrict__ S,
                                  const int8_t* __restrict__ found, int32_t* JI, int32_t* JJ) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < T; q += gridDim.x * blockDim.x) {
    if (found[q]) { size_t k = (size_t)order[q]; JI[k] = S[q].x; JJ[k] = S[q].y; }
  }
}
__global__ void k_fill_i32(int32_t* a, size_t n, int32_t v) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) a[k] = v;
}
int bilin_materialize_mask(sosie_ctx* ctx) {
  if (!ctx->mask_is_ones) return SOSIE_OK;
  const size_t nt = (size_t)ctx->tg.nx * ctx->tg.ny;
  CK(ctx->t_mask.alloc(nt));
  k_fill_i32<<<grid_for(ctx, nt, 256), 256, 0, ctx->stream>>>(ctx->t_mask.p, nt, 1);
  KLAUNCH_CHECK();
  ctx->mask_is_ones = false;
  return SOSIE_OK;
}
int fill_i32_dev(sosie_ctx* ctx, int32_t* p, size_t n, int32_t v) {
  k_fill_i32<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(p, n, v);
  KLAUNCH_CHECK();
  return SOSIE_OK;
}
// WHERE(JIpos==-1) mask_t=-1 ; WHERE(JJpos==-1) mask_t=-2   (:1434-1435)
__global__ void k_twisted_mask(const int32_t* __restrict__ JI, const int32_t* __restrict__ JJ, int32_t* mask, size_t n) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    if (JI[k] == -1) mask[k] = -1;
    if (JJ[k] == -1) mask[k] = -2;
  }
}

// ===================================================================================================
// host orchestration
__global__ void k_to_mailbox(const unsigned char* __restrict__ src, unsigned char* dst, size_t n) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) dst[k] = src[k];
}
int fetch_small(sosie_ctx* ctx, const void* dsrc, void* hdst, size_t bytes) {
  if (bytes == 0) return SOSIE_OK;
  if (bytes > SOSIE_MAILBOX_BYTES) {
    CK(cudaMemcpyAsync(hdst, dsrc, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SOSIE_OK;
  }
  if (!ctx->mailbox) CK(cudaHostAlloc((void**)&ctx->mailbox, SOSIE_MAILBOX_BYTES, cudaHostAllocMapped | cudaHostAllocPortable));
  unsigned char* dmb = nullptr;
  CK(cudaHostGetDevicePointer((void**)&dmb, ctx->mailbox, 0));
  k_to_mailbox<<<bytes > 4096 ? 32 : 1, 256, 0, ctx->stream>>>((const unsigned char*)dsrc, dmb, bytes);
  KLAUNCH_CHECK();
  CK(cudaStreamSynchronize(ctx->stream));
  memcpy(hdst, ctx->mailbox, bytes);
  return SOSIE_OK;
}
// two objects, one stream synchronisation
int fetch_pair(sosie_ctx* ctx, const void* d1, void* h1, size_t b1, const void* d2, void* h2, size_t b2) {
  const size_t o2 = (b1 + 15) & ~(size_t)15;
  if (o2 + b2 > SOSIE_MAILBOX_BYTES) { if (int rc = fetch_small(ctx, d1, h1, b1)) return rc; return fetch_small(ctx, d2, h2, b2); }
  if (!ctx->mailbox) CK(cudaHostAlloc((void**)&ctx->mailbox, SOSIE_MAILBOX_BYTES, cudaHostAllocMapped | cudaHostAllocPortable));
  unsigned char* dmb = nullptr;
  CK(cudaHostGetDevicePointer((void**)&dmb, ctx->mailbox, 0));
  k_to_mailbox<<<32, 256, 0, ctx->stream>>>((const unsigned char*)d1, dmb, b1); KLAUNCH_CHECK();
  k_to_mailbox<<<32, 256, 0, ctx->stream>>>((const unsigned char*)d2, dmb + o2, b2); KLAUNCH_CHECK();
  CK(cudaStreamSynchronize(ctx->stream));
  memcpy(h1, ctx->mailbox, b1);
  memcpy(h2, ctx->mailbox + o2, b2);
  return SOSIE_OK;
}
static int fetch_d(sosie_ctx* ctx, const double* dptr, double* out) { return fetch_small(ctx, dptr, out, sizeof(double)); }

// coordinate part of BILIN_2D's prologue (:97-148): decide l_add_extra_j, build the working source
int bilin_prepare_src(sosie_ctx* ctx) {
  SrcGeom& sg = ctx->sg;
  const int nx = sg.nx, ny = sg.ny;
  const double* Xin = ctx->s_in_X.p;
  const double* Yin = ctx->s_in_Y.p;
  const int in_1d = sg.separable;  // at this point: "input given as 1-D axes"
  cudaStream_t st = ctx->stream;
  ctx->l_add_extra_j = 0;
  double dyp = 0.0;
  ctx->h_axes.clear();
  if (in_1d) {  // one read-back of both axes serves everything the host decides from them (this routine and the EASY search)
    ctx->h_axes.resize((size_t)nx + ny + 2);
    if (ctx->axes_hint_dev && ctx->axes_hint.size() == (size_t)nx + ny) {
      // device-resident axes whose host copy the caller has declared (sosie_bilin_src_axes_hint): no read-back, the host
      // never waits for the stream here
      memcpy(ctx->h_axes.data(), ctx->axes_hint.data(), sizeof(double) * ((size_t)nx + ny));
      if (ctx->axes_hint_verify) {
        std::vector<double> chk((size_t)nx + ny);
        if (int rc = fetch_pair(ctx, Xin, chk.data(), sizeof(double) * nx, Yin, chk.data() + nx, sizeof(double) * ny)) return rc;
        if (memcmp(chk.data(), ctx->axes_hint.data(), sizeof(double) * chk.size()) != 0) {
          ctx->err = "sosie_bilin_src_axes_hint: the declared host copy differs from the device axes"; return SOSIE_ERR_ARG;
        }
      }
    } else if (int rc = fetch_pair(ctx, Xin, ctx->h_axes.data(), sizeof(double) * nx, Yin, ctx->h_axes.data() + nx, sizeof(double) * ny)) return rc;
  }
  if (ctx->l_reg_src && (ny > 20) && (nx > 20)) {
    double ymx, ym1;
    int ih = nx / 2;  // Y1(nx1/2, ny1)
    if (in_1d) { ym1 = ctx->h_axes[nx + ny - 2]; ymx = ctx->h_axes[nx + ny - 1]; }
    else {
      if (fetch_d(ctx, Yin + (size_t)(ih - 1) + (size_t)(ny - 1) * nx, &ymx)) return SOSIE_ERR_CUDA;
      if (fetch_d(ctx, Yin + (size_t)(ih - 1) + (size_t)(ny - 2) * nx, &ym1)) return SOSIE_ERR_CUDA;
    }
    if ((ymx < 90.0) && (2.0 * ymx - ym1 >= 90.0)) {
      ctx->l_add_extra_j = 1;
      dyp = ymx - ym1;  // YY(nx/2,ny) - YY(nx/2,ny-1)
      if (!in_1d) {     // EXT_NORTH_TO_90_REGG sanity STOPs (:1803-1811) on the last row
        std::vector<double> row(nx);
        CK(cudaMemcpyAsync(row.data(), Yin + (size_t)(ny - 1) * nx, sizeof(double) * nx, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        double s = 0.0;
        for (int i = 0; i < nx; i++) { double d = row[i] - ymx; s += d * d; }
        if (s > (double)1.E-12f) { ctx->err = "EXT_NORTH_TO_90_REGG: grid doesnt seem to be regular"; return SOSIE_ERR_REF_STOP; }
      }
    }
  }
  if (in_1d) {  // the working latitude axis (pole rows included), as k_lat_table builds it on the device
    if (ctx->l_add_extra_j) { ctx->h_axes[nx + ny] = 90.0; ctx->h_axes[nx + ny + 1] = 90.0 + dyp; }
    else ctx->h_axes.resize((size_t)nx + ny);
  }
  sg.nyw = ny + (ctx->l_add_extra_j ? 2 : 0);
  const int nyw = sg.nyw;
  // The tables below depend on the axes, k_ew and the pole extension only.  When the caller has declared the axes
  // (sosie_bilin_src_axes_hint) and they equal those of the previous call, the tables of that call are still valid: a record
  // loop that re-enters BILIN_2D's first call per target grid, or the bench's repeated step, skips seven launches.
  const bool hinted = in_1d && ctx->axes_hint_dev && ctx->axes_hint.size() == (size_t)nx + ny;
  if (hinted && ctx->tables_valid && ctx->tables_kew == ctx->k_ew && ctx->tables_extra == ctx->l_add_extra_j && ctx->tables_axes == ctx->h_axes &&
      ctx->tables_Xin == Xin && ctx->s_lat1.n >= (size_t)nyw && ctx->s_clon.n >= (size_t)nx) {
    sg.merc = ctx->s_merc.p; sg.separable = 1; sg.lon1 = Xin; sg.lat1 = ctx->s_lat1.p;
    sg.clon = ctx->s_clon.p; sg.slon = ctx->s_slon.p; sg.clat = ctx->s_clat.p; sg.slat = ctx->s_slat.p;
    double* hd = ctx->s_head.p;
    sg.hE = hd; sg.hW = hd + nx; sg.hN = hd + 2 * nx; sg.hS = hd + 2 * nx + nyw; sg.lonm = hd + 2 * nx + 2 * nyw;
    sg.lonm_same = ctx->tables_lonm_same;
    ctx->tables_reused = true;
    return SOSIE_OK;
  }
  ctx->tables_valid = false; ctx->tables_reused = false;
  if (in_1d) {
    CK(ctx->s_lat1.alloc(nyw));
    CK(ctx->s_clon.alloc(nx)); CK(ctx->s_slon.alloc(nx)); CK(ctx->s_clat.alloc(nyw)); CK(ctx->s_slat.alloc(nyw));
    k_lat_table<<<cdiv(nyw, 256), 256, 0, st>>>(Yin, ny, nyw, dyp, ctx->s_lat1.p); KLAUNCH_CHECK();
    k_trig_table<<<cdiv(nx, 256), 256, 0, st>>>(Xin, nx, ctx->s_clon.p, ctx->s_slon.p); KLAUNCH_CHECK();
    k_trig_table<<<cdiv(nyw, 256), 256, 0, st>>>(ctx->s_lat1.p, nyw, ctx->s_clat.p, ctx->s_slat.p); KLAUNCH_CHECK();
    CK(ctx->s_merc.alloc(nyw));
    k_merc_table<<<cdiv(nyw, 256), 256, 0, st>>>(ctx->s_lat1.p, nyw, ctx->s_merc.p); KLAUNCH_CHECK();
    sg.merc = ctx->s_merc.p;
    sg.separable = 1;
    sg.lon1 = Xin; sg.lat1 = ctx->s_lat1.p;
    sg.clon = ctx->s_clon.p; sg.slon = ctx->s_slon.p; sg.clat = ctx->s_clat.p; sg.slat = ctx->s_slat.p;
    CK(ctx->s_head.alloc(3 * (size_t)nx + 2 * (size_t)nyw));
    double* hd = ctx->s_head.p;
    k_heading_tables<<<cdiv(nx + nyw, 128), 128, 0, st>>>(sg, ctx->k_ew, hd, hd + nx, hd + 2 * nx, hd + 2 * nx + nyw, hd + 2 * nx + 2 * nyw); KLAUNCH_CHECK();
    sg.hE = hd; sg.hW = hd + nx; sg.hN = hd + 2 * nx; sg.hS = hd + 2 * nx + nyw; sg.lonm = hd + 2 * nx + 2 * nyw;
    sg.lonm_same = 1;
    for (int i = 0; i < nx; i++) if (!(fabs(ctx->h_axes[i]) < 360.0)) { sg.lonm_same = 0; break; }
  } else {
    size_t n = (size_t)nx * nyw;
    CK(ctx->s_X.alloc(n)); CK(ctx->s_Y.alloc(n)); CK(ctx->s_vx.alloc(n)); CK(ctx->s_vy.alloc(n)); CK(ctx->s_vz.alloc(n));
    k_build_src2d<<<grid_for(ctx, n, 256), 256, 0, st>>>(Xin, Yin, nx, ny, nyw, 0, dyp, ctx->s_X.p, ctx->s_Y.p, ctx->s_vx.p,
                                                          ctx->s_vy.p, ctx->s_vz.p); KLAUNCH_CHECK();
    CK(ctx->s_merc.alloc(n));
    k_merc_table<<<grid_for(ctx, n, 256), 256, 0, st>>>(ctx->s_Y.p, n, ctx->s_merc.p); KLAUNCH_CHECK();
    sg.merc = ctx->s_merc.p;
    sg.separable = 0;
    sg.X = ctx->s_X.p; sg.Y = ctx->s_Y.p; sg.vx = ctx->s_vx.p; sg.vy = ctx->s_vy.p; sg.vz = ctx->s_vz.p;
  }
  if (ctx->l_add_extra_j) {
    CK(ctx->d_im.alloc(nx));
    const double* lonrow = in_1d ? Xin : Xin + (size_t)(ny - 1) * nx;
    int unsorted = 0;
    if (in_1d) {
      for (int i = 1; i < nx; i++) if (ctx->h_axes[i] < ctx->h_axes[i - 1]) { unsorted = 1; break; }
    } else {
      CK(ctx->d_iscratch.alloc(SOSIE_ISCRATCH));
      CK(cudaMemsetAsync(ctx->d_iscratch.p + 1, 0, sizeof(int), st));
      k_check_sorted<<<cdiv(nx, 256), 256, 0, st>>>(lonrow, nx, ctx->d_iscratch.p + 1); KLAUNCH_CHECK();
      if (int rc = fetch_small(ctx, ctx->d_iscratch.p + 1, &unsorted, sizeof(int))) return rc;
    }
    k_jim_table<<<cdiv(nx, 128), 128, 0, st>>>(lonrow, nx, unsorted ? 0 : 1, ctx->d_im.p); KLAUNCH_CHECK();
  }
  if (hinted) {
    ctx->tables_valid = true; ctx->tables_kew = ctx->k_ew; ctx->tables_extra = ctx->l_add_extra_j; ctx->tables_axes = ctx->h_axes;
    ctx->tables_Xin = Xin; ctx->tables_lonm_same = sg.lonm_same;
  }
  return SOSIE_OK;
}

// 0: not sorted; 1: non-decreasing; 2: increasing with every step >= 1e-9 and |v| <= 1e5 (minloc_abs_sorted's well_spaced)
static int is_sorted_host(const std::vector<double>& v, int off, int n) {
  bool spaced = true;
  for (int i = 0; i < n; i++) {
    if (i > 0 && v[off + i] < v[off + i - 1]) return 0;
    if (i > 0 && !(v[off + i] - v[off + i - 1] >= 1.0e-9)) spaced = false;
    if (!(std::fabs(v[off + i]) <= 1.0e5)) spaced = false;
  }
  return spaced ? 2 : 1;
}

// ---- FIND_NEAREST_POINT prelude (:888-947): regularity tests, latitude ranges, target row range.
// xrow_a..xrow_b (1-based, inclusive; 0,0 = all): rows of the target longitude array that are resident -- the
// pipelined first call (bilin_pipeline.cu) uploads only row 1 and its shard's rows.  L_IS_GRID_REGULAR is then
// evaluated on those rows; an irregular row settles it (any one does), otherwise *need_full_x is set and the
// caller must complete the array and call again.
int prelude_finish(sosie_ctx* ctx, MapPlan* pl, const Prelude& hp);
int bilin_map_prelude(sosie_ctx* ctx, MapPlan* pl, int xrow_a, int xrow_b, bool* need_full_x) {
  SrcGeom& sg = ctx->sg;
  TrgGeom& tg = ctx->tg;
  cudaStream_t st = ctx->stream;
  const size_t nt = (size_t)tg.nx * tg.ny;
  const size_t nsrc = (size_t)sg.nx * sg.nyw;
  if (need_full_x) *need_full_x = false;
  ctx->ws_cursor = 0;
  CK(ctx->d_scratch.alloc(64)); CK(ctx->d_iscratch.alloc(SOSIE_ISCRATCH));
  struct { Prelude* p; } dpre = {reinterpret_cast<Prelude*>(ctx->d_scratch.p)};
  static_assert(sizeof(Prelude) <= 64 * sizeof(double), "Prelude does not fit the scratch buffer");
  k_prelude_init<<<1, 1, 0, st>>>(dpre.p); KLAUNCH_CHECK();

  // y_min_s / y_max_s : MINVAL/MAXVAL(Ysrc) over the working array (pole rows included)
  if (sg.separable) { k_minmax<<<grid_for(ctx, sg.nyw, 256), 256, 0, st>>>(sg.lat1, sg.nyw, &dpre.p->ymin_s, &dpre.p->ymax_s); KLAUNCH_CHECK(); }
  else { k_minmax<<<grid_for(ctx, nsrc, 256), 256, 0, st>>>(sg.Y, nsrc, &dpre.p->ymin_s, &dpre.p->ymax_s); KLAUNCH_CHECK(); }
  if (!sg.separable) { k_is_regular<<<grid_for(ctx, (size_t)sg.nyw * 32, 256), 256, 0, st>>>(sg.X, sg.Y, sg.nx, sg.nyw, 2, sg.nyw, &dpre.p->irreg_s); KLAUNCH_CHECK(); }
  const bool partial_x = !tg.is_1d && xrow_a >= 1 && !(xrow_a <= 2 && xrow_b >= tg.ny);
  if (!tg.is_1d) {
    int ja = partial_x ? std::max(2, xrow_a) : 2, jb = partial_x ? xrow_b : tg.ny;
    k_is_regular<<<grid_for(ctx, (size_t)tg.ny * 32, 256), 256, 0, st>>>(tg.X, tg.Y, tg.nx, tg.ny, ja, jb, &dpre.p->irreg_t); KLAUNCH_CHECK();
  }
  // the target latitudes are read ONCE (k_ypass); the row range is computed unconditionally and used only when the
  // reference computes it
  const int do_dlat = (tg.nx > 2) ? 1 : 0;
  if (do_dlat) { k_rtmp<<<1, 1024, 0, st>>>(tg, dpre.p); KLAUNCH_CHECK(); }
  const int nchunk = (tg.ny + JR_ROWS - 1) / JR_ROWS;
  const size_t tot = (size_t)tg.nx * nchunk;
  WS(ColPart, part, tot);
  k_ypass<<<grid_for(ctx, tot, 256), 256, 0, st>>>(tg, dpre.p, do_dlat, part.p, 1, tg.ny); KLAUNCH_CHECK();
  k_jrange_cols<<<cdiv(tg.nx, 128), 128, 0, st>>>(tg.nx, nchunk, part.p, dpre.p); KLAUNCH_CHECK();
  Prelude hp;
  if (int rc = fetch_small(ctx, dpre.p, &hp, sizeof(Prelude))) return rc;
  if (partial_x && hp.irreg_t == 0) { if (need_full_x) *need_full_x = true; return SOSIE_OK; }
  return prelude_finish(ctx, pl, hp);
}

// ---------------------------------------------------------------------------------------------------
// The prelude over ONE rank's target rows, exchanged between the ranks of a row-sharded job (sosie_set_exchange): every
// quantity the prelude reduces from the target latitudes is a min / max / first-occurrence over rows, so the ranks reduce
// their own rows (plus the row below, for the row-to-row differences; plus the two rows of :921, which every rank holds),
// all-gather one small record each -- scalars and one ColPart per target column -- and combine them in row order into
// exactly the Prelude a pass over the whole array gives.  The regularity test is only conclusive one way: a rank whose
// rows differ from row 1 proves the grid irregular (sums of non-negative terms only grow); when no rank can tell, the
// caller uploads the whole arrays and runs the ordinary prelude.
int bilin_map_prelude_exchange(sosie_ctx* ctx, MapPlan* pl, int ja, int jb, bool* undecided) {
  SrcGeom& sg = ctx->sg;
  TrgGeom& tg = ctx->tg;
  cudaStream_t st = ctx->stream;
  *undecided = false;
  const size_t blob = sizeof(XchgHead) + sizeof(ColPart) * (size_t)tg.nx;
  if (!ctx->xchg_fn || ctx->xchg_nranks < 2 || blob > SOSIE_MAILBOX_BYTES || tg.is_1d || !sg.separable) { *undecided = true; return SOSIE_OK; }
  ctx->ws_cursor = 0;
  CK(ctx->d_scratch.alloc(64)); CK(ctx->d_iscratch.alloc(SOSIE_ISCRATCH));
  Prelude* dp = reinterpret_cast<Prelude*>(ctx->d_scratch.p);
  k_prelude_init<<<1, 1, 0, st>>>(dp); KLAUNCH_CHECK();
  k_minmax<<<grid_for(ctx, sg.nyw, 256), 256, 0, st>>>(sg.lat1, sg.nyw, &dp->ymin_s, &dp->ymax_s); KLAUNCH_CHECK();
  k_is_regular<<<grid_for(ctx, (size_t)tg.ny * 32, 256), 256, 0, st>>>(tg.X, tg.Y, tg.nx, tg.ny, std::max(2, ja), jb, &dp->irreg_t, ja, jb); KLAUNCH_CHECK();
  const int do_dlat = (tg.nx > 2) ? 1 : 0;
  if (do_dlat) { k_rtmp<<<1, 1024, 0, st>>>(tg, dp); KLAUNCH_CHECK(); }
  const int nchunk = (jb - ja + 1 + JR_ROWS - 1) / JR_ROWS;
  const size_t tot = (size_t)tg.nx * nchunk;
  WS(ColPart, part, tot);
  WS(unsigned char, dblob, blob);
  k_ypass<<<grid_for(ctx, tot, 256), 256, 0, st>>>(tg, dp, do_dlat, part.p, ja, jb); KLAUNCH_CHECK();
  k_colpart_combine<<<cdiv(tg.nx, 128), 128, 0, st>>>(tg.nx, nchunk, part.p, reinterpret_cast<ColPart*>(dblob.p + sizeof(XchgHead))); KLAUNCH_CHECK();
  std::vector<unsigned char> mine(blob), all(blob * (size_t)ctx->xchg_nranks);
  Prelude hp;
  if (int rc = fetch_pair(ctx, dp, &hp, sizeof(Prelude), dblob.p + sizeof(XchgHead), mine.data() + sizeof(XchgHead), sizeof(ColPart) * (size_t)tg.nx)) return rc;
  XchgHead hd;
  hd.j_a = ja; hd.j_b = jb; hd.irreg_t = hp.irreg_t; hd.nx = tg.nx;
  hd.ymin_t = hp.ymin_t; hd.ymax_t = hp.ymax_t; hd.rmin_dlat = hp.rmin_dlat; hd.rtmp = hp.rtmp;
  memcpy(mine.data(), &hd, sizeof(hd));
  if (ctx->xchg_fn(ctx->xchg_user, mine.data(), all.data(), blob) != 0) { ctx->err = "prelude exchange: the all-gather callback failed"; return SOSIE_ERR_ARG; }
  // combine in row order
  std::vector<int> order(ctx->xchg_nranks);
  for (int r = 0; r < ctx->xchg_nranks; r++) order[r] = r;
  auto head = [&](int r) { XchgHead h; memcpy(&h, all.data() + blob * (size_t)r, sizeof(h)); return h; };
  std::sort(order.begin(), order.end(), [&](int a, int b) { return head(a).j_a < head(b).j_a; });
  int expect = 1, irreg = 0;
  for (int q = 0; q < ctx->xchg_nranks; q++) {
    const XchgHead h = head(order[q]);
    if (h.nx != tg.nx || h.j_a != expect || h.j_b < h.j_a) { ctx->err = "prelude exchange: the ranks' row shards do not tile the target grid"; return SOSIE_ERR_ARG; }
    expect = h.j_b + 1;
    hp.ymin_t = std::min(hp.ymin_t, h.ymin_t); hp.ymax_t = std::max(hp.ymax_t, h.ymax_t); hp.rmin_dlat = std::min(hp.rmin_dlat, h.rmin_dlat);
    irreg |= h.irreg_t;
  }
  if (expect != tg.ny + 1) { ctx->err = "prelude exchange: the ranks' row shards do not tile the target grid"; return SOSIE_ERR_ARG; }
  if (!irreg) { *undecided = true; return SOSIE_OK; }
  hp.irreg_t = irreg;
  int j_strt = 2147483647, j_stop = 0;
  for (int i = 0; i < tg.nx; i++) {
    unsigned long long emin = ~0ULL, emax = 0ULL;
    int jmin = 0, jmax = 0;
    for (int q = 0; q < ctx->xchg_nranks; q++) {
      ColPart cp;
      memcpy(&cp, all.data() + blob * (size_t)order[q] + sizeof(XchgHead) + sizeof(ColPart) * (size_t)i, sizeof(cp));
      if (cp.jmin > 0 && cp.emin < emin) { emin = cp.emin; jmin = cp.jmin; }
      if (cp.jmax > 0 && (cp.emax > emax || jmax == 0)) { emax = cp.emax; jmax = cp.jmax; }
    }
    if (jmin > 0) j_strt = std::min(j_strt, jmin);
    j_stop = std::max(j_stop, jmax);
  }
  hp.j_strt = j_strt; hp.j_stop = j_stop;
  return prelude_finish(ctx, pl, hp);
}

// host part of the prelude: the decisions of FIND_NEAREST_POINT (:888-947) from the reduced scalars
int prelude_finish(sosie_ctx* ctx, MapPlan* pl, const Prelude& hp) {
  TrgGeom& tg = ctx->tg;
  const double y_min_s = dec_d_host(hp.ymin_s), y_max_s = dec_d_host(hp.ymax_s);
  const double y_min_t = dec_d_host(hp.ymin_t), y_max_t = dec_d_host(hp.ymax_t);
  const bool l_is_reg_s = (hp.irreg_s == 0), l_is_reg_t = (hp.irreg_t == 0);
  double rmin_dlat_dj = (double)-100.f;
  if (tg.nx > 2 && tg.ny >= 2) rmin_dlat_dj = dec_d_host(hp.rmin_dlat);
  int j_strt_t = 1, j_stop_t = tg.ny, jlat_icr = 1;
  if ((rmin_dlat_dj > (double)-1.E-12f) || l_is_reg_t) {
    j_strt_t = hp.j_strt; j_stop_t = hp.j_stop;
    if (j_strt_t > j_stop_t) jlat_icr = -1;
    if (j_strt_t < 1 || j_strt_t > tg.ny || j_stop_t < 1 || j_stop_t > tg.ny) {
      ctx->err = "FIND_NEAREST_POINT: target latitudes not covered by the source (the reference loops out of bounds)";
      return SOSIE_ERR_REF_STOP;
    }
  }
  ctx->map_info[0] = j_strt_t; ctx->map_info[1] = j_stop_t; ctx->map_info[2] = jlat_icr;
  ctx->map_info[3] = l_is_reg_s; ctx->map_info[4] = l_is_reg_t; ctx->map_info[6] = 0; ctx->map_info[7] = ctx->l_add_extra_j;
  pl->j_strt = j_strt_t; pl->j_stop = j_stop_t; pl->icr = jlat_icr;
  pl->jlo = j_strt_t < j_stop_t ? j_strt_t : j_stop_t; pl->jhi = j_strt_t < j_stop_t ? j_stop_t : j_strt_t;
  pl->reg_s = l_is_reg_s; pl->reg_t = l_is_reg_t;
  pl->y_min_s = y_min_s; pl->y_max_s = y_max_s; pl->y_min_t = y_min_t; pl->y_max_t = y_max_t;
  return SOSIE_OK;
}

// EASY search, source-only preparation: Xsrc(:,3) / Ysrc(3,:) axes, their sortedness, uniform-axis hints
int bilin_easy_prepare(sosie_ctx* ctx) {
  SrcGeom& sg = ctx->sg;
  cudaStream_t st = ctx->stream;
  if (ctx->tables_reused && ctx->easy_valid && ctx->d_vax.n >= (size_t)(sg.nx + sg.nyw)) {   // same axes as the previous mapping
    sg.lon0 = ctx->easy_hint[0]; sg.inv_dlon = ctx->easy_hint[1]; sg.lat0 = ctx->easy_hint[2]; sg.inv_dlat = ctx->easy_hint[3];
    return SOSIE_OK;
  }
  ctx->easy_valid = false;
  CK(ctx->d_vax.alloc(sg.nx + sg.nyw));
  double* vax = ctx->d_vax.p;
  k_gather_axes<<<cdiv(sg.nx + sg.nyw, 256), 256, 0, st>>>(sg, vax, vax + sg.nx); KLAUNCH_CHECK();
  std::vector<double> hax(sg.nx + sg.nyw);
  if (sg.separable && ctx->h_axes.size() == hax.size()) hax = ctx->h_axes;   // Xsrc(:,3) = lon axis, Ysrc(3,:) = working lat axis
  else if (int rc = fetch_small(ctx, vax, hax.data(), sizeof(double) * hax.size())) return rc;
  ctx->easy_sorted_lon = is_sorted_host(hax, 0, sg.nx);
  ctx->easy_sorted_lat = is_sorted_host(hax, sg.nx, sg.nyw);
  sg.lon0 = sg.inv_dlon = sg.lat0 = sg.inv_dlat = 0.0;
  if (ctx->easy_sorted_lon && hax[sg.nx - 1] > hax[0]) { sg.lon0 = hax[0]; sg.inv_dlon = (double)(sg.nx - 1) / (hax[sg.nx - 1] - hax[0]); }
  if (ctx->easy_sorted_lat && hax[sg.nx + sg.nyw - 1] > hax[sg.nx]) {
    sg.lat0 = hax[sg.nx]; sg.inv_dlat = (double)(sg.nyw - 1) / (hax[sg.nx + sg.nyw - 1] - hax[sg.nx]);
  }
  ctx->easy_hint[0] = sg.lon0; ctx->easy_hint[1] = sg.inv_dlon; ctx->easy_hint[2] = sg.lat0; ctx->easy_hint[3] = sg.inv_dlat;
  ctx->easy_valid = ctx->tables_valid;
  return SOSIE_OK;
}

// FIND_NEAREST_EASY + MAPPING_BL body for the targets [k0, k0+kcnt), searched rows jlo..jhi; errors OR'ed into *derr
static int bilin_easy_rows_plan(sosie_ctx* ctx, int jlo, int jhi, size_t k0, size_t kcnt, int* derr, const DevPlan* plan);
int bilin_easy_rows(sosie_ctx* ctx, int jlo, int jhi, size_t k0, size_t kcnt, int* derr) { return bilin_easy_rows_plan(ctx, jlo, jhi, k0, kcnt, derr, nullptr); }
static int bilin_easy_rows_plan(sosie_ctx* ctx, int jlo, int jhi, size_t k0, size_t kcnt, int* derr, const DevPlan* plan) {
  SrcGeom& sg = ctx->sg;
  TrgGeom& tg = ctx->tg;
  cudaStream_t st = ctx->stream;
  if (kcnt == 0) return SOSIE_OK;
  const double* vax = ctx->d_vax.p;
  if (sg.separable)
    k_map_easy<true><<<grid_for(ctx, kcnt, 128, 16), 128, 0, st>>>(sg, tg, ctx->k_ew, ctx->mask_is_ones ? nullptr : ctx->t_mask.p, vax, vax + sg.nx, ctx->easy_sorted_lon,
                                                                 ctx->easy_sorted_lat, jlo, jhi, ctx->d_metrics.p, ctx->d_ab.p,
                                                                 ctx->d_ipb.p, ctx->d_dnp.p, derr, k0, kcnt, ctx->search_only ? 1 : 0, plan, make_fastdiv((unsigned)tg.nx));
  else
    k_map_easy<false><<<grid_for(ctx, kcnt, 128, 16), 128, 0, st>>>(sg, tg, ctx->k_ew, ctx->mask_is_ones ? nullptr : ctx->t_mask.p, vax, vax + sg.nx, ctx->easy_sorted_lon,
                                                                  ctx->easy_sorted_lat, jlo, jhi, ctx->d_metrics.p, ctx->d_ab.p,
                                                                  ctx->d_ipb.p, ctx->d_dnp.p, derr, k0, kcnt, ctx->search_only ? 1 : 0, plan, make_fastdiv((unsigned)tg.nx));
  KLAUNCH_CHECK();
  return SOSIE_OK;
}

int bilin_map_alloc(sosie_ctx* ctx) {
  const size_t nt = (size_t)ctx->tg.nx * ctx->tg.ny;
  CK(ctx->d_metrics.alloc(3 * nt)); CK(ctx->d_ab.alloc(2 * nt)); CK(ctx->d_ipb.alloc(nt)); CK(ctx->d_dnp.alloc(nt));
  return SOSIE_OK;
}

static int bilin_map_device_plan(sosie_ctx* ctx, bool async);
int bilin_map_impl(sosie_ctx* ctx, bool async) {
  if (!ctx->grids_set) { ctx->err = "sosie_bilin_map: grids not set"; return SOSIE_ERR_STATE; }
  if (ctx->trg_partial) {
    ctx->err = "sosie_bilin_map: the sharded first call kept only this rank's target coordinates on the device; call sosie_bilin_set_grids again";
    return SOSIE_ERR_STATE;
  }
  SrcGeom& sg = ctx->sg;
  TrgGeom& tg = ctx->tg;
  cudaStream_t st = ctx->stream;
  const size_t nt = (size_t)tg.nx * tg.ny;
  const size_t nsrc = (size_t)sg.nx * sg.nyw;
  if (int rc = bilin_map_alloc(ctx)) return rc;
  ctx->map_check_pending = false;
  if (sg.separable && !ctx->search_only) return bilin_map_device_plan(ctx, async);
  MapPlan pl;
  if (int rc = bilin_map_prelude(ctx, &pl, 0, 0, nullptr)) return rc;
  struct { int* p; } derr = {reinterpret_cast<int*>(ctx->d_iscratch.p)};
  CK(cudaMemsetAsync(derr.p, 0, sizeof(int), st));
  const double y_min_s = pl.y_min_s, y_max_s = pl.y_max_s, y_min_t = pl.y_min_t, y_max_t = pl.y_max_t;
  const bool l_is_reg_s = pl.reg_s;
  const int j_strt_t = pl.j_strt, j_stop_t = pl.j_stop, jlat_icr = pl.icr;

  int32_t* MT = ctx->d_metrics.p;
  double* AB = ctx->d_ab.p;
  int16_t* IDP = ctx->d_ipb.p;
  float* DNP = ctx->d_dnp.p;

  if (l_is_reg_s) {
    // ---- FIND_NEAREST_EASY + body, fused
    ctx->map_info[5] = 0;
    if (int rc = bilin_easy_prepare(ctx)) return rc;
    size_t k0, kcnt;
    shard_range(ctx, &k0, &kcnt);
    prof_begin(ctx, SOSIE_T_MAP);
    if (int rc = bilin_easy_rows(ctx, pl.jlo, pl.jhi, k0, kcnt, derr.p)) return rc;
    prof_end(ctx, SOSIE_T_MAP);
  } else {
    // ---- FIND_NEAREST_TWISTED
    ctx->map_info[5] = 1;
    if (int rc = bilin_materialize_mask(ctx)) return rc;   // this search writes -1 / -2 into the mask
    if ((y_min_t >= y_max_s) || (y_max_t <= y_min_s)) { ctx->err = "FIND_NEAREST_TWISTED: Target and source latitudes do not overlap!"; return SOSIE_ERR_REF_STOP; }
    if (sg.separable) { ctx->err = "internal: TWISTED with separable source"; return SOSIE_ERR_ARG; }
    WS(double, e1, nsrc); WS(double, e2, nsrc);
    k_src_metrics<<<grid_for(ctx, nsrc, 128), 128, 0, st>>>(sg, e1.p, e2.p); KLAUNCH_CHECK();
    // IS_POLAR_PROJ on source and target (:1171,1181)
    auto polar = [&](const double* X, const double* Y, int nx, int ny, int is1d, double ymax, bool* lpp) -> int {
      *lpp = false;
      if (!(ymax > 88.0)) return 0;
      const int nb = 64;
      WS(MinIdx, part, nb);
      size_t n = (size_t)nx * ny;
      k_argmin_np<<<nb, 256, 0, st>>>(Y, n, is1d, nx, part.p); KLAUNCH_CHECK();
      MinIdx hpart[nb];
      CK(cudaMemcpyAsync(hpart, part.p, sizeof(hpart), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      MinIdx r = hpart[0];
      for (int b = 1; b < nb; b++) {
        MinIdx c = hpart[b];
        if (c.k < 0) continue;
        if (r.k < 0 || c.v < r.v || (c.v == r.v && c.k < r.k)) r = c;
      }
      int iNP = (int)(r.k % nx) + 1, jNP = (int)(r.k / nx) + 1;
      if (jNP - 3 < 1 || jNP + 3 > ny) return 0;  // Q5: out-of-range probe DEFINED as "not polar"
      auto getY = [&](int i, int j, double* v) -> int { return fetch_d(ctx, is1d ? Y + (j - 1) : Y + (size_t)(i - 1) + (size_t)(j - 1) * nx, v); };
      auto getX = [&](int i, int j, double* v) -> int { return fetch_d(ctx, is1d ? X + (i - 1) : X + (size_t)(i - 1) + (size_t)(j - 1) * nx, v); };
      double ym, y0, yp, xm, xp;
      if (getY(iNP, jNP - 3, &ym) || getY(iNP, jNP, &y0) || getY(iNP, jNP + 3, &yp)) return SOSIE_ERR_CUDA;
      bool lipp = (ym < y0) && (yp < y0);
      if (lipp) {
        if (getX(iNP, jNP - 3, &xm) || getX(iNP, jNP + 3, &xp)) return SOSIE_ERR_CUDA;
        double d = fabs(xm - xp);
        lipp = (d > 100.0) && (d < 220.0);
      }
      *lpp = lipp;
      return 0;
    };
    bool lNPs, lNPt;
    if (int rc = polar(sg.X, sg.Y, sg.nx, sg.nyw, 0, y_max_s, &lNPs)) return rc;
    if (int rc = polar(tg.X, tg.Y, tg.nx, tg.ny, tg.is_1d, y_max_t, &lNPt)) return rc;
    const bool lStupido = lNPs || lNPt;
    int nsplit = 0;  // undefined in the reference when lStupido; DEFINED 0 (see oracle)
    DBuf<double> dVB; DBuf<int> dJV;
    if (!lStupido) {
      double y_max_bnd = std::min(y_max_t, y_max_s), y_max_bnd0 = y_max_bnd;
      double y_min_bnd = std::max(y_min_t, y_min_s), y_min_bnd0 = y_min_bnd;
      y_max_bnd = std::min((double)(int)(y_max_bnd + 2.0), 90.0);
      y_min_bnd = std::max((double)(int)(y_min_bnd - 2.0), -90.0);
      y_max_bnd = std::min((double)lround(y_max_bnd / 0.5) * 0.5, 90.0);
      y_min_bnd = std::max((double)lround(y_min_bnd / 0.5) * 0.5, -90.0);
      y_max_bnd = std::max(y_max_bnd, y_max_bnd0);
      y_min_bnd = std::min(y_min_bnd, y_min_bnd0);
      nsplit = std::max((int)(((double)40.0f * (y_max_bnd - y_min_bnd)) / 180.0), 1);
      std::vector<double> VB(nsplit + 1);
      double dy = (y_max_bnd - y_min_bnd) / (double)(float)nsplit;
      for (int jl = 1; jl <= nsplit + 1; jl++) VB[jl - 1] = y_min_bnd + (double)(float)(jl - 1) * dy;
      if ((y_min_bnd > y_min_bnd0) || (y_max_bnd < y_max_bnd0)) { ctx->err = "FIND_NEAREST_POINT: Bounds for latitude for VLAT_SPLIT_BOUNDS are bad!"; return SOSIE_ERR_REF_STOP; }
      { WS(double, tvb, nsplit + 1); WS(int, tjv, 2 * nsplit); dVB.borrow(tvb.p, nsplit + 1); dJV.borrow(tjv.p, 2 * nsplit); }
      CK(cudaMemcpyAsync(dVB.p, VB.data(), sizeof(double) * (nsplit + 1), cudaMemcpyHostToDevice, st));
      WS(int, jmx, nsplit); WS(int, jmn, nsplit);
      CK(cudaMemsetAsync(jmx.p, 0, sizeof(int) * nsplit, st));
      k_fill_i32<<<1, 64, 0, st>>>(jmn.p, nsplit, 2147483647); KLAUNCH_CHECK();
      k_jvlat<<<grid_for(ctx, (size_t)sg.nx * nsplit, 128), 128, 0, st>>>(sg, nsplit, dVB.p, jmx.p, jmn.p); KLAUNCH_CHECK();
      std::vector<int> hmx(nsplit), hmn(nsplit), JV(2 * nsplit);
      CK(cudaMemcpyAsync(hmx.data(), jmx.p, sizeof(int) * nsplit, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(hmn.data(), jmn.p, sizeof(int) * nsplit, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      for (int b = 0; b < nsplit; b++) {
        JV[b] = std::max(hmn[b] - 1, 1);
        JV[b + nsplit] = std::min(hmx[b] + 1, sg.nyw);
        if (JV[b + nsplit] <= JV[b]) { ctx->err = "FIND_NEAREST_TWISTED: ERROR: jj_stop > jj_start !"; return SOSIE_ERR_REF_STOP; }
      }
      CK(cudaMemcpyAsync(dJV.p, JV.data(), sizeof(int) * 2 * nsplit, cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
    }
    TwParams tp;
    tp.lStupido = lStupido ? 1 : 0; tp.nsplit = nsplit; tp.VB = dVB.p; tp.JV = dJV.p; tp.e1 = e1.p; tp.e2 = e2.p;
    tp.frac_emax = (double)0.51f; tp.sqrt2f = (double)sqrtf(2.0f);
    tp.seg = nullptr; tp.nseg = (sg.nx + TW_TX - 1) / TW_TX;
    const int nty = (sg.nyw + TW_TY - 1) / TW_TY;
    WS(double, segt, (size_t)nty * tp.nseg * 5 + 8);
    if (!sg.separable && !getenv("SOSIE_TWISTED_NO_PRUNE")) {
      k_seg_table<<<grid_for(ctx, (size_t)nty * tp.nseg * 32, 256), 256, 0, st>>>(sg, tp.nseg, nty, segt.p); KLAUNCH_CHECK();
      tp.seg = segt.p;
    }
    // processing order of the searched targets
    long ntrip = ((long)j_stop_t - (long)j_strt_t + jlat_icr) / jlat_icr;
    if (ntrip < 0) ntrip = 0;
    size_t np = (size_t)ntrip * tg.nx;
    WS(int32_t, JI, nt); WS(int32_t, JJ, nt);
    k_fill_i32<<<grid_for(ctx, nt, 256), 256, 0, st>>>(JI.p, nt, -1); KLAUNCH_CHECK();
    k_fill_i32<<<grid_for(ctx, nt, 256), 256, 0, st>>>(JJ.p, nt, -1); KLAUNCH_CHECK();
    int T = 0;
    WS(int64_t, order, np + 1);
    if (np > 0) {
      WS(uint8_t, flags, np); WS(int64_t, lin, np); WS(int, dnum, 1);
      k_twisted_flags<<<grid_for(ctx, np, 256), 256, 0, st>>>(tg, ctx->t_mask.p, j_strt_t, jlat_icr, (int)ntrip, flags.p, lin.p); KLAUNCH_CHECK();
      size_t tmpb = 0;
      cub::DeviceSelect::Flagged(nullptr, tmpb, lin.p, flags.p, order.p, dnum.p, (int)np, st);
      WS(uint8_t, tmp, tmpb + 16);
      CK(cub::DeviceSelect::Flagged(tmp.p, tmpb, lin.p, flags.p, order.p, dnum.p, (int)np, st));
      ctx->launches += 2;
      CK(cudaMemcpyAsync(&T, dnum.p, sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
    }
    if (T > 0) {
      WS(int2, bpos, T); WS(int2, S0, T); WS(int2, S1, T); WS(int2, S2, T); WS(int8_t, bfound, T); WS(int8_t, found, T);
      prof_begin(ctx, SOSIE_T_MAP);
      k_twisted_band<<<grid_for(ctx, (size_t)T * 32, 256), 256, 0, st>>>(sg, tg, tp, order.p, T, bpos.p, bfound.p, derr.p); KLAUNCH_CHECK();
      prof_end(ctx, SOSIE_T_MAP);
      CK(cudaMemcpyAsync(S0.p, bpos.p, sizeof(int2) * T, cudaMemcpyDeviceToDevice, st));
      int2* Sa = S0.p;
      int2* Sb = S1.p;
      int2* Sp = nullptr;   // no state of two passes ago before the second pass
      // The chain is iterated to its fixed point; once a pass changes nothing, none does.  Passes are launched four at a
      // time with one "changed" flag each, so the host synchronises once per batch instead of once per pass.
      int iters = 0;
      const int NB = 4;
      WS(int, dchg4, NB);
      bool done = false;
      while (!done) {
        CK(cudaMemsetAsync(dchg4.p, 0, sizeof(int) * NB, st));
        for (int b = 0; b < NB; b++) {
          k_twisted_chain<<<grid_for(ctx, T, 128), 128, 0, st>>>(sg, tg, tp, order.p, T, bpos.p, bfound.p, Sa, Sb, found.p, dchg4.p + b, Sp); KLAUNCH_CHECK();
          { int2* t = Sp ? Sp : S2.p; Sp = Sa; Sa = Sb; Sb = t; }   // (prev, old, new) <- (old, new, prev)
        }
        int chg[NB];
        if (int rc = fetch_small(ctx, dchg4.p, chg, sizeof(int) * NB)) return rc;
        for (int b = 0; b < NB; b++) {
          iters++;
          if (!chg[b]) { done = true; break; }   // iters = passes up to and including the first one that changed nothing
        }
        if (!done && iters > T + 2) { ctx->err = "internal: TWISTED chain did not converge"; return SOSIE_ERR_CUDA; }
      }
      ctx->map_info[6] = iters;
      k_twisted_scatter<<<grid_for(ctx, T, 256), 256, 0, st>>>(order.p, T, Sa, found.p, JI.p, JJ.p); KLAUNCH_CHECK();
    }
    k_twisted_mask<<<grid_for(ctx, nt, 256), 256, 0, st>>>(JI.p, JJ.p, ctx->t_mask.p, nt); KLAUNCH_CHECK();
    if (ctx->search_only) {
      CK(cudaMemcpyAsync(MT, JI.p, nt * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(MT + nt, JJ.p, nt * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    } else {
      k_map_body<<<grid_for(ctx, nt, 128, 16), 128, 0, st>>>(sg, tg, ctx->k_ew, ctx->t_mask.p, JI.p, JJ.p, MT, AB, IDP, DNP, derr.p); KLAUNCH_CHECK();
    }
    CK(cudaStreamSynchronize(st));
  }
  int herr = 0;
  if (int rc = fetch_small(ctx, derr.p, &herr, sizeof(int))) return rc;
  if (herr & 1) { ctx->err = "ERROR (L_InPoly of mod_poly.f90): we only expect positive longitudes!"; return SOSIE_ERR_REF_STOP; }
  if (herr & 6) { ctx->err = "FIND_NEAREST: The nearest point was not found! (ji_s or jj_s = 0)"; return SOSIE_ERR_REF_STOP; }
  ctx->map_ready = !ctx->search_only;   // a search-only run leaves no mapping behind
  ctx->map_nt = nt;
  if (l_is_reg_s) map_rows_from_shard(ctx); else ctx->map_j0 = ctx->map_j1 = 0;   // the TWISTED search is never split
  ctx->pack_ready = false; ctx->unpacked_applies = 0;
  return SOSIE_OK;
}


// ---------------------------------------------------------------------------------------------------
// Regular (separable) source: the whole mapping without a host round trip.  The prelude's reductions stay on the device,
// k_prelude_finish turns them into the searched row range, k_map_easy reads it from there.  With a device-side exchange
// (sosie_set_exchange_dev) a row-sharded context reduces only ITS rows of the target latitudes and the ranks all-gather
// one record each (24 bytes per target column) GPU -> GPU on the context stream (NCCL in bench.py), merged by a kernel in
// row order: the per-rank cost of the prelude no longer grows with the whole grid.  L_IS_GRID_REGULAR on the target still
// runs on the whole (resident) arrays -- it exits at the first irregular row.
// async: nothing is read back; errors (rows not covered by the source, L_InPoly's STOP, nearest point not found) and
// map_info are delivered by the next synchronising entry (bilin_map_check).
static int bilin_map_device_plan(sosie_ctx* ctx, bool async) {
  SrcGeom& sg = ctx->sg;
  TrgGeom& tg = ctx->tg;
  cudaStream_t st = ctx->stream;
  const size_t nt = (size_t)tg.nx * tg.ny;
  ctx->ws_cursor = 0;
  CK(ctx->d_scratch.alloc(64)); CK(ctx->d_iscratch.alloc(SOSIE_ISCRATCH));
  static_assert(sizeof(Prelude) <= 48 * sizeof(double) && sizeof(DevPlan) <= 8 * sizeof(double), "scratch layout");
  Prelude* dp = reinterpret_cast<Prelude*>(ctx->d_scratch.p);
  DevPlan* dplan = reinterpret_cast<DevPlan*>(ctx->d_scratch.p + 48);
  int* derr = ctx->d_iscratch.p;
  CK(cudaMemsetAsync(derr, 0, sizeof(int), st));
  const int do_dlat = (tg.nx > 2) ? 1 : 0;
  k_prelude_head<<<1, 1024, 0, st>>>(tg, sg.lat1, sg.nyw, do_dlat, dp); KLAUNCH_CHECK();
  if (!tg.is_1d) { k_is_regular<<<grid_for(ctx, (size_t)tg.ny * 32, 256), 256, 0, st>>>(tg.X, tg.Y, tg.nx, tg.ny, 2, tg.ny, &dp->irreg_t); KLAUNCH_CHECK(); }
  size_t k0, kcnt;
  shard_range(ctx, &k0, &kcnt);
  const bool sharded = ctx->shard_j0 >= 1;
  const size_t blob = sizeof(XchgHead) + sizeof(ColPart) * (size_t)tg.nx;
  const bool xchg = sharded && ctx->xchg_dev_fn && ctx->xchg_dev_nranks >= 2 && ctx->xchg_dev_nranks <= 256 && !tg.is_1d;
  const bool xp2p = sharded && ctx->p2p_ready && ctx->p2p_nranks >= 2 && !tg.is_1d && blob <= ctx->p2p_slot;
  ctx->xchg_used = 0;
  if (xp2p) {
    // one producer + one consumer kernel over peer memory instead of combine + collective + merge
    const int ja = ctx->shard_j0, jb = ctx->shard_j1;
    const int nchunk = (jb - ja + 1 + JR_ROWS - 1) / JR_ROWS;
    WS(ColPart, part, (size_t)tg.nx * nchunk);
    const unsigned epoch = ++ctx->p2p_epoch;
    const int par = (int)(epoch & 1u);
    k_ypass<<<grid_for(ctx, (size_t)tg.nx * nchunk, 256), 256, 0, st>>>(tg, dp, do_dlat, part.p, ja, jb); KLAUNCH_CHECK();
    k_xchg_push<<<cdiv(tg.nx, 128), 128, 0, st>>>(tg.nx, nchunk, part.p, dp, ja, jb, ctx->p2p_peers_dev, ctx->p2p_nranks, ctx->p2p_rank, ctx->p2p_slot,
                                                  par, epoch, ctx->p2p_done); KLAUNCH_CHECK();
    k_xchg_merge<<<std::max(1, std::min(64, cdiv(tg.nx, 256))), 256, 0, st>>>(ctx->p2p_nranks, ctx->p2p_box, ctx->p2p_slot, par, epoch, tg.nx, tg.ny, dp,
                                                                              ctx->p2p_spin_cycles);
    KLAUNCH_CHECK();
    ctx->xchg_used = 2;
  } else if (xchg) {
    const int ja = ctx->shard_j0, jb = ctx->shard_j1;
    const int nchunk = (jb - ja + 1 + JR_ROWS - 1) / JR_ROWS;
    WS(ColPart, part, (size_t)tg.nx * nchunk);
    WS(unsigned char, dblob, blob);
    WS(unsigned char, dall, blob * (size_t)ctx->xchg_dev_nranks);
    k_ypass<<<grid_for(ctx, (size_t)tg.nx * nchunk, 256), 256, 0, st>>>(tg, dp, do_dlat, part.p, ja, jb); KLAUNCH_CHECK();
    k_colpart_combine<<<cdiv(tg.nx, 128), 128, 0, st>>>(tg.nx, nchunk, part.p, reinterpret_cast<ColPart*>(dblob.p + sizeof(XchgHead)), dp, ja, jb,
                                                        reinterpret_cast<XchgHead*>(dblob.p)); KLAUNCH_CHECK();
    if (ctx->xchg_dev_fn(ctx->xchg_dev_user, dblob.p, dall.p, blob, (void*)st) != 0) { ctx->err = "prelude exchange: the device all-gather callback failed"; return SOSIE_ERR_ARG; }
    k_prelude_merge<<<std::max(1, std::min(64, cdiv(tg.nx, 256))), 256, 0, st>>>(ctx->xchg_dev_nranks, dall.p, blob, tg.nx, tg.ny, dp); KLAUNCH_CHECK();
    ctx->xchg_used = 1;
  } else {
    const int nchunk = (tg.ny + JR_ROWS - 1) / JR_ROWS;
    WS(ColPart, part, (size_t)tg.nx * nchunk);
    k_ypass<<<grid_for(ctx, (size_t)tg.nx * nchunk, 256), 256, 0, st>>>(tg, dp, do_dlat, part.p, 1, tg.ny); KLAUNCH_CHECK();
    k_jrange_cols<<<cdiv(tg.nx, 128), 128, 0, st>>>(tg.nx, nchunk, part.p, dp); KLAUNCH_CHECK();
  }
  k_prelude_finish<<<1, 1, 0, st>>>(dp, tg.nx, tg.ny, dplan); KLAUNCH_CHECK();
  ctx->map_info[5] = 0;
  if (int rc = bilin_easy_prepare(ctx)) return rc;
  prof_begin(ctx, SOSIE_T_MAP);
  if (int rc = bilin_easy_rows_plan(ctx, 1, 0, k0, kcnt, derr, dplan)) return rc;
  prof_end(ctx, SOSIE_T_MAP);
  ctx->map_ready = true;
  ctx->map_nt = nt;
  map_rows_from_shard(ctx);
  ctx->pack_ready = false; ctx->unpacked_applies = 0;
  ctx->map_check_pending = true;
  if (async) return SOSIE_OK;
  return bilin_map_check(ctx);
}

// the deferred part of a device-planned mapping: one read-back of the prelude, the plan and the error flags
int bilin_map_check(sosie_ctx* ctx) {
  if (!ctx->map_check_pending) return SOSIE_OK;
  ctx->map_check_pending = false;
  struct { Prelude p; unsigned char pad[48 * sizeof(double) - sizeof(Prelude)]; DevPlan plan; } h;
  static_assert(sizeof(h) <= 64 * sizeof(double), "scratch layout");
  int herr = 0;
  if (int rc = fetch_pair(ctx, ctx->d_scratch.p, &h, sizeof(h), ctx->d_iscratch.p, &herr, sizeof(int))) return rc;
  ctx->map_info[0] = h.plan.j_strt; ctx->map_info[1] = h.plan.j_stop; ctx->map_info[2] = h.plan.icr;
  ctx->map_info[3] = 1; ctx->map_info[4] = (h.p.irreg_t == 0); ctx->map_info[5] = 0; ctx->map_info[6] = 0; ctx->map_info[7] = ctx->l_add_extra_j;
  const char* msg = nullptr;
  int code = SOSIE_ERR_REF_STOP;
  if (h.plan.err & 4) { msg = "prelude exchange over peer memory: a rank's record did not arrive in time (did every rank call the mapping?)"; code = SOSIE_ERR_STATE; }
  else if (h.plan.err & 2) { msg = "prelude exchange: the ranks' row shards do not tile the target grid"; code = SOSIE_ERR_ARG; }
  else if (h.plan.err & 1) msg = "FIND_NEAREST_POINT: target latitudes not covered by the source (the reference loops out of bounds)";
  else if (herr & 1) msg = "ERROR (L_InPoly of mod_poly.f90): we only expect positive longitudes!";
  else if (herr & 6) msg = "FIND_NEAREST: The nearest point was not found! (ji_s or jj_s = 0)";
  if (msg) { ctx->err = msg; ctx->map_ready = false; return code; }
  return SOSIE_OK;
}
