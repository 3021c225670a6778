// bilin_apply.cu -- per-slab application of the cached bilinear mapping.
//
// GPU counterpart of the apply loop of BILIN_2D (src/mod_bilin_2d.f90:209-263) and of INTERP_BL
// (:275-361).  HBM-bound gather-FMA:
//   * a one-off "pack" kernel turns (iP,jP,iqdrn,alpha,beta) into one 32-byte record per target:
//     int4 = the four corner offsets in the working source slab (column-major, pole rows
//     included), float4 = the four REAL(4) weights exactly as INTERP_BL rounds them (:334-337).
//     Identical-point copies (:227-230) and the error codes -9998..-9994 (:347-356) are encoded
//     in the record, so the hot kernel has no coordinate reads and no index arithmetic;
//   * the apply kernel keeps a target's record in registers and loops over the slabs of its
//     chunk (records x levels batched in ONE launch), issuing the 4 gathers of several slabs
//     before consuming them.  Neighbouring targets hit neighbouring source cells, so a slab's
//     gathers are served by L1/L2 and each source element leaves HBM once;
//   * the two virtual pole rows of EXT_NORTH_TO_90_REGG (src/mod_manip.f90:1821-1832) are
//     evaluated on the fly from the slab instead of materialising Z1w per record;
//   * optional fused post-ops of INTERP_2D (clamp :97-98, target mask :133-137) and the
//     corr_vect rotation (src/corr_vect.f90:622-632).
// Arithmetic is REAL(4) in the reference's order with FMA contraction off -> bit-identical fields.
#include "pack.cuh"

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pack(SrcGeom sg, TrgGeom tg, int k_ew, const int32_t* __restrict__ MT, const double* __restrict__ AB, int4* pidx,
       float4* pw, size_t k0, size_t kcnt) {
  const size_t nt = (size_t)tg.nx * tg.ny;
  for (size_t kk = blockIdx.x * (size_t)blockDim.x + threadIdx.x; kk < kcnt; kk += (size_t)gridDim.x * blockDim.x) {
    const size_t k = k0 + kk;
    int jj = (int)(k / tg.nx) + 1, ji = (int)(k % tg.nx) + 1;
    int4 id;
    float4 w;
    pack_record(sg, k_ew, MT[k], MT[k + nt], MT[k + 2 * nt], AB[k], AB[k + nt], trg_lon(tg, ji, jj), trg_lat(tg, ji, jj), id, w);
    pidx[k] = id;
    pw[k] = w;
  }
}

// value of the working slab Z1w at linear offset `lin` (pole rows virtual)
struct SlabView {
  int nx, ny;            // physical slab extents
  int nreal;             // nx*ny
  const int32_t* im;     // ji_m table (1-based), only when pole rows exist
};
__device__ __forceinline__ float zget(const float* __restrict__ Z, const SlabView& v, int lin) {
  if (lin < v.nreal) return __ldg(Z + lin);
  int r = lin - v.nreal;
  if (r < v.nx) {  // row ny+1 : 0.5*(XF(ji,ny) + XF(ji_m,ny))
    int i = r, m = v.im[i] - 1;
    const float* row = Z + (size_t)(v.ny - 1) * v.nx;
    return 0.5f * (__ldg(row + i) + __ldg(row + m));
  }
  int i = r - v.nx, m = v.im[i] - 1;  // row ny+2 : XF(ji_m,ny-1)
  return __ldg(Z + (size_t)(v.ny - 2) * v.nx + m);
}

__device__ __forceinline__ float interp_one(const float* __restrict__ Z, const SlabView& v, const int4& id, const float4& w,
                                            float wup) {
  if (id.x >= 0) {
    float z1 = zget(Z, v, id.x), z2 = zget(Z, v, id.y), z3 = zget(Z, v, id.z), z4 = zget(Z, v, id.w);
    // ( Z_in(i1,j1)*w1 + Z_in(i2,j2)*w2 + Z_in(i3,j3)*w3 + Z_in(i4,j4)*w4 )/wup   (:358)
    return __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(z1, w.x), __fmul_rn(z2, w.y)), __fmul_rn(z3, w.z)), __fmul_rn(z4, w.w)), wup);
  }
  if (id.x == REC_COPY) return zget(Z, v, id.y);
  return w.x;
}

struct TileGeom {  // 32x8 target tiles over rows j0..j1 (0-based, inclusive) of an (nx, ny) target grid
  int nx, j0, j1, tiles_x, ntiles;
};

struct PostOps {
  int use_clamp;
  float vmin, vmax;
  const int32_t* mask_trg;
  float rmiss;
};

#define APPLY_SLABS_PER_CHUNK 8
#define APPLY_SLABS_PER_CHUNK_STAGED 32
// One slab (the first call of a 2-D job, the per-record call of the unchanged driver loop): nothing to amortise the
// record over, so the kernel is a latency chain record -> 4 gathers -> store.  A lean body (no slab unrolling: half the
// registers of a slab-unrolled body) keeps every SM at full occupancy to cover it.
// FUSE: a single-slab apply that finds no packed records yet builds the record of each target from the mapping arrays
// right here and uses it from registers: one pass instead of k_pack (mapping in, records out) followed by a second one
// that reads the records back.  pidx_out / pw_out non-null: the records are also stored for the later calls (the second
// such apply of a mapping; the very first one does not store them -- see bilin_apply_impl).
#ifndef APPLY1_FUSE_BLOCKS
#define APPLY1_FUSE_BLOCKS 8
#endif
template <bool FUSE>
__global__ void __launch_bounds__(256, FUSE ? APPLY1_FUSE_BLOCKS : 8)
k_apply1(const int4* __restrict__ pidx, const float4* __restrict__ pw, SlabView v, const float* __restrict__ Z1, float* __restrict__ Z2,
         PostOps po, TileGeom tgm, SrcGeom sg, TrgGeom tg, int k_ew, const int32_t* __restrict__ MT, const double* __restrict__ AB,
         int4* pidx_out, float4* pw_out) {
  const size_t nt = FUSE ? (size_t)tg.nx * tg.ny : 0;
  for (int tile = blockIdx.x; tile < tgm.ntiles; tile += gridDim.x) {
    const int ti = tile % tgm.tiles_x, tj = tile / tgm.tiles_x;
    const int i = ti * 32 + (threadIdx.x & 31), j = tgm.j0 + tj * 8 + (threadIdx.x >> 5);  // 0-based
    if (i >= tgm.nx || j > tgm.j1) continue;
    const size_t k = (size_t)i + (size_t)j * tgm.nx;
    int4 id;
    float4 w;
    if (FUSE) {
      pack_record(sg, k_ew, __ldcs(MT + k), __ldcs(MT + k + nt), __ldcs(MT + k + 2 * nt), __ldcs(AB + k), __ldcs(AB + k + nt),
                  trg_lon(tg, i + 1, j + 1), trg_lat(tg, i + 1, j + 1), id, w);
      if (pidx_out) {   // kept for the later calls (not on the very first apply of a mapping: see bilin_apply_impl)
        __stcs(pidx_out + k, id);
        __stcs(pw_out + k, w);
      }
    } else {
      id = __ldcs(pidx + k);     // streamed once: do not displace the source slab in L2
      w = __ldcs(pw + k);
    }
    const float wup = __fadd_rn(__fadd_rn(__fadd_rn(w.x, w.y), w.z), w.w);
    float r = interp_one(Z1, v, id, w, wup);
    if (po.use_clamp) { r = fmaxf(r, po.vmin); r = fminf(r, po.vmax); }
    if (po.mask_trg && po.mask_trg[k] == 0) r = po.rmiss;
    __stcs(Z2 + k, r);
  }
}

// ---------------------------------------------------------------------------------------------------
// Staged variant for batched slabs.  The gathers of a 32x8 target tile fall in a small box of the source
// slab (a few hundred cells), but on a target whose rows cut across source rows (polar stereographic: the
// lanes of a warp step through ~10 source rows) every warp-level gather touches ~12 sectors and L1TEX
// saturates at ~30 % of HBM speed (profiles/ r01).  Here the box is computed once per tile at pack time,
// each slab's box is streamed into shared memory with cp.async (row segments, NSTAGE slabs in flight), and
// the four gathers become shared-memory reads: one pass over HBM, L1TEX traffic cut ~5x.
#define BOXMAX 512
#define NSTAGE 8
#define BOX_PER_THREAD (BOXMAX / 256)

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// one block per tile: bounding box (0-based i0, j0, w, h) of the tile's source cells; w == 0 -> not stageable
__global__ void __launch_bounds__(256)
k_tilebox(const int4* __restrict__ pidx, SlabView v, TileGeom tgm, int4* boxes) {
  __shared__ int sh[4][8];
  __shared__ int shbad[8];
  for (int tile = blockIdx.x; tile < tgm.ntiles; tile += gridDim.x) {
    const int ti = tile % tgm.tiles_x, tj = tile / tgm.tiles_x;
    const int i = ti * 32 + (threadIdx.x & 31), j = tgm.j0 + tj * 8 + (threadIdx.x >> 5);
    int imin = 2147483647, imax = -1, jmin = 2147483647, jmax = -1, bad = 0;
    if (i < tgm.nx && j <= tgm.j1) {
      const int4 id = pidx[(size_t)i + (size_t)j * tgm.nx];
      if (id.x < 0) bad = 1;
      else {
        int c[4] = {id.x, id.y, id.z, id.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if (c[q] >= v.nreal) { bad = 1; continue; }
          int ii = c[q] % v.nx, jj = c[q] / v.nx;
          imin = min(imin, ii); imax = max(imax, ii); jmin = min(jmin, jj); jmax = max(jmax, jj);
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      imin = min(imin, __shfl_xor_sync(0xffffffffu, imin, o)); imax = max(imax, __shfl_xor_sync(0xffffffffu, imax, o));
      jmin = min(jmin, __shfl_xor_sync(0xffffffffu, jmin, o)); jmax = max(jmax, __shfl_xor_sync(0xffffffffu, jmax, o));
      bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    const int wq = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sh[0][wq] = imin; sh[1][wq] = imax; sh[2][wq] = jmin; sh[3][wq] = jmax; shbad[wq] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int q = 1; q < 8; q++) {
        sh[0][0] = min(sh[0][0], sh[0][q]); sh[1][0] = max(sh[1][0], sh[1][q]);
        sh[2][0] = min(sh[2][0], sh[2][q]); sh[3][0] = max(sh[3][0], sh[3][q]);
        shbad[0] |= shbad[q];
      }
      int w = sh[1][0] - sh[0][0] + 1, h = sh[3][0] - sh[2][0] + 1;
      int4 b = make_int4(0, 0, 0, 0);
      if (!shbad[0] && sh[1][0] >= 0 && (long long)w * h <= BOXMAX) b = make_int4(sh[0][0], sh[2][0], w, h);
      boxes[tile] = b;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256, 4)
k_apply_staged(const int4* __restrict__ pidx, const float4* __restrict__ pw, size_t nt, SlabView v, const float* __restrict__ Z1,
               float* __restrict__ Z2, int nslab, int slabs_per_chunk, PostOps po, TileGeom tgm, const int4* __restrict__ boxes) {
  __shared__ float buf[NSTAGE][BOXMAX];
  const size_t src_stride = (size_t)v.nreal;
  const int s0 = blockIdx.y * slabs_per_chunk;
  const int s1 = min(s0 + slabs_per_chunk, nslab);
  const int nsl = s1 - s0;
  for (int tile = blockIdx.x; tile < tgm.ntiles; tile += gridDim.x) {
    const int ti = tile % tgm.tiles_x, tj = tile / tgm.tiles_x;
    const int i = ti * 32 + (threadIdx.x & 31), j = tgm.j0 + tj * 8 + (threadIdx.x >> 5);
    const bool valid = (i < tgm.nx && j <= tgm.j1);
    const size_t k = valid ? (size_t)i + (size_t)j * tgm.nx : 0;
    int4 id = make_int4(REC_CONST, 0, 0, 0);
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) { id = pidx[k]; w = pw[k]; }
    const float wup = __fadd_rn(__fadd_rn(__fadd_rn(w.x, w.y), w.z), w.w);
    const bool masked = (valid && po.mask_trg) ? (po.mask_trg[k] == 0) : false;
    const int4 box = boxes[tile];
    if (box.z == 0) {  // not stageable (special records, pole rows, E-W wrap, huge footprint): direct gathers
      if (valid)
        for (int s = s0; s < s1; s++) {
          float r = interp_one(Z1 + (size_t)s * src_stride, v, id, w, wup);
          if (po.use_clamp) { r = fmaxf(r, po.vmin); r = fminf(r, po.vmax); }
          if (masked) r = po.rmiss;
          __stcs(Z2 + (size_t)s * nt + k, r);
        }
      continue;  // block-uniform
    }
    const int bw = box.z, nelem = box.z * box.w;
    int o1 = 0, o2 = 0, o3 = 0, o4 = 0;
    if (valid) {
      o1 = (id.x / v.nx - box.y) * bw + (id.x % v.nx - box.x);
      o2 = (id.y / v.nx - box.y) * bw + (id.y % v.nx - box.x);
      o3 = (id.z / v.nx - box.y) * bw + (id.z % v.nx - box.x);
      o4 = (id.w / v.nx - box.y) * bw + (id.w % v.nx - box.x);
    }
    // per-thread running pointers: source cells this thread stages (advance by one slab per issue), output cell
    const float* gsrc[BOX_PER_THREAD];
    bool gok[BOX_PER_THREAD];
#pragma unroll
    for (int m = 0; m < BOX_PER_THREAD; m++) {
      int e = threadIdx.x + 256 * m;
      gok[m] = e < nelem;
      gsrc[m] = Z1 + (size_t)s0 * src_stride + (gok[m] ? (size_t)(box.y + e / bw) * v.nx + box.x + e % bw : 0);
    }
    float* outp = Z2 + (size_t)s0 * nt + k;
    float* const sbase = &buf[0][0] + threadIdx.x;
    auto issue = [&](int stage) {
#pragma unroll
      for (int m = 0; m < BOX_PER_THREAD; m++) {
        if (gok[m]) cp_async4(sbase + stage * BOXMAX + 256 * m, gsrc[m]);
        gsrc[m] += src_stride;
      }
    };
#pragma unroll
    for (int q = 0; q < NSTAGE - 1; q++) {
      if (q < nsl) issue(q);
      cp_async_commit();
    }
    int nissued = NSTAGE - 1;
    for (int q0 = 0; q0 < nsl; q0 += NSTAGE) {
#pragma unroll
      for (int u = 0; u < NSTAGE; u++) {  // stage index == u: compile-time shared-memory offsets
        const int q = q0 + u;
        if (q < nsl) {
          cp_async_wait<NSTAGE - 2>();  // this thread's copies of slab q have landed
          __syncthreads();              // ... everybody's have, and everybody is done reading slab q-1's buffer
          if (nissued < nsl) issue((u + NSTAGE - 1) % NSTAGE);  // refills the buffer slab q-1 used
          nissued++;
          cp_async_commit();
          if (valid) {
            const float* b = &buf[u][0];
            float r = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(b[o1], w.x), __fmul_rn(b[o2], w.y)), __fmul_rn(b[o3], w.z)), __fmul_rn(b[o4], w.w)), wup);
            if (po.use_clamp) { r = fmaxf(r, po.vmin); r = fminf(r, po.vmax); }
            if (masked) r = po.rmiss;
            __stcs(outp, r);
            outp += nt;
          }
        }
      }
    }
    __syncthreads();  // the next tile's prologue overwrites the buffers
    cp_async_wait<0>();
  }
}

// (U,V) pair + rotation on the target grid (src/corr_vect.f90:622-624,630,632)
__global__ void __launch_bounds__(256)
k_apply_uv(const int4* __restrict__ pidx, const float4* __restrict__ pw, size_t nt, SlabView v, const float* __restrict__ U1,
           const float* __restrict__ V1, float* __restrict__ Uo, float* __restrict__ Vo, const double* __restrict__ xcos,
           const double* __restrict__ xsin, int nslab, int slabs_per_chunk, TileGeom tgm) {
  const size_t src_stride = (size_t)v.nreal;
  const int s0 = blockIdx.y * slabs_per_chunk;
  const int s1 = min(s0 + slabs_per_chunk, nslab);
  for (int tile = blockIdx.x; tile < tgm.ntiles; tile += gridDim.x) {
    const int ti = tile % tgm.tiles_x, tj = tile / tgm.tiles_x;
    const int i = ti * 32 + (threadIdx.x & 31), j = tgm.j0 + tj * 8 + (threadIdx.x >> 5);
    if (i >= tgm.nx || j > tgm.j1) continue;
    const size_t k = (size_t)i + (size_t)j * tgm.nx;
    const int4 id = pidx[k];
    const float4 w = pw[k];
    const float wup = __fadd_rn(__fadd_rn(__fadd_rn(w.x, w.y), w.z), w.w);
    const double c = xcos[k], sn = xsin[k];
    for (int s = s0; s < s1; s++) {
      double zU = (double)interp_one(U1 + (size_t)s * src_stride, v, id, w, wup);
      double zV = (double)interp_one(V1 + (size_t)s * src_stride, v, id, w, wup);
      float uo = (float)__dadd_rn(__dmul_rn(c, zU), __dmul_rn(sn, zV));
      float vo = (float)__dsub_rn(__dmul_rn(c, zV), __dmul_rn(sn, zU));
      __stcs(Uo + (size_t)s * nt + k, uo);
      __stcs(Vo + (size_t)s * nt + k, vo);
    }
  }
}

// rmeanv = SUM(Z2(:,ny2))/nx2 sequentially (:244)
__global__ void k_lastrow_mean(const float* __restrict__ Z2, int nx2, int ny2, float* out) {
  float s = 0.f;
  const float* row = Z2 + (size_t)(ny2 - 1) * nx2;
  for (int i = 0; i < nx2; i++) s = __fadd_rn(s, row[i]);
  *out = __fdiv_rn(s, (float)nx2);
}
// section copy dst(l0+k*ls, jl) = src(r0+k*rs, jr) (1-based), RHS read before any store (two kernels)
__global__ void k_sect_read(const float* Z, int nx, int n, int r0, int rs, int jr, float* tmp) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    int ir = r0 + k * rs;
    tmp[k] = (ir >= 1 && ir <= nx) ? Z[(size_t)(ir - 1) + (size_t)(jr - 1) * nx] : 0.f;
  }
}
__global__ void k_sect_write(float* Z, int nx, int n, int l0, int ls, int jl, const float* tmp) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    Z[(size_t)(l0 + k * ls - 1) + (size_t)(jl - 1) * nx] = tmp[k];
}

// ---------------------------------------------------------------------------------------------------
int bilin_pack_range(sosie_ctx* ctx, size_t k0, size_t kcnt) {
  const size_t nt = (size_t)ctx->tg.nx * ctx->tg.ny;
  if ((size_t)ctx->sg.nx * ctx->sg.nyw > 2147483647ULL) { ctx->err = "source slab too large for 32-bit offsets"; return SOSIE_ERR_ARG; }
  CK(ctx->d_pidx.alloc(nt)); CK(ctx->d_pw.alloc(nt));
  if (kcnt == 0) return SOSIE_OK;
  k_pack<<<grid_for(ctx, kcnt, 256), 256, 0, ctx->stream>>>(ctx->sg, ctx->tg, ctx->k_ew, ctx->d_metrics.p, ctx->d_ab.p, ctx->d_pidx.p, ctx->d_pw.p, k0, kcnt);
  KLAUNCH_CHECK();
  return SOSIE_OK;
}

int bilin_pack_impl(sosie_ctx* ctx) {
  if (!ctx->map_ready) { ctx->err = "bilinear mapping neither built nor loaded"; return SOSIE_ERR_STATE; }
  size_t k0, kcnt;
  shard_range(ctx, &k0, &kcnt);
  prof_begin(ctx, SOSIE_T_PACK);
  if (int rc = bilin_pack_range(ctx, k0, kcnt)) return rc;
  prof_end(ctx, SOSIE_T_PACK);
  ctx->pack_ready = true;
  ctx->pack_gen++;
  ctx->boxes_ready = false;
  return SOSIE_OK;
}

// rows (0-based, physical) of the source slab that the packed records [k0, k0+kcnt) read: the host-pointer paths
// upload only those (a regional target touches a fraction of a global slab).  The virtual pole rows read rows
// ny-2 and ny-1 (zget).  row_lo > row_hi: nothing is read (all records constant).
__global__ void k_src_rowrange(const int4* __restrict__ pidx, size_t k0, size_t kcnt, int nx, int ny, int* out) {
  int lo = 2147483647, hi = -1;
  const int nreal = nx * ny;
  for (size_t kk = blockIdx.x * (size_t)blockDim.x + threadIdx.x; kk < kcnt; kk += (size_t)gridDim.x * blockDim.x) {
    const int4 id = pidx[k0 + kk];
    int c[4] = {id.x, id.y, id.z, id.w};
    int n = 4;
    if (id.x == REC_COPY) { c[0] = id.y; n = 1; }
    else if (id.x < 0) n = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if (q >= n) break;
      if (c[q] >= nreal) { lo = min(lo, ny - 2); hi = max(hi, ny - 1); }
      else { int r = c[q] / nx; lo = min(lo, r); hi = max(hi, r); }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&out[0], lo); atomicMax(&out[1], hi); }
}

__global__ void k_rowrange_init(int* out) { out[0] = 2147483647; out[1] = -1; }

int bilin_src_rowrange(sosie_ctx* ctx, size_t k0, size_t kcnt, int* row_lo, int* row_hi) {
  CK(ctx->d_iscratch.alloc(SOSIE_ISCRATCH));
  int* d = ctx->d_iscratch.p + 2;
  k_rowrange_init<<<1, 1, 0, ctx->stream>>>(d); KLAUNCH_CHECK();
  if (kcnt > 0) {
    k_src_rowrange<<<grid_for(ctx, kcnt, 256), 256, 0, ctx->stream>>>(ctx->d_pidx.p, k0, kcnt, ctx->sg.nx, ctx->sg.ny, d);
    KLAUNCH_CHECK();
  }
  int h[2];
  if (int rc = fetch_small(ctx, d, h, sizeof(h))) return rc;
  *row_lo = h[0]; *row_hi = h[1];
  return SOSIE_OK;
}

static void apply_grid(const sosie_ctx* ctx, int nslab, dim3& grid, int& spc, TileGeom& tgm) {
  size_t k0, kcnt;
  shard_range(ctx, &k0, &kcnt);
  tgm.nx = ctx->tg.nx;
  tgm.j0 = (int)(k0 / (size_t)ctx->tg.nx);
  tgm.j1 = tgm.j0 + (int)(kcnt / (size_t)ctx->tg.nx) - 1;
  tgm.tiles_x = (tgm.nx + 31) / 32;
  int tiles_y = (tgm.j1 - tgm.j0 + 1 + 7) / 8;
  tgm.ntiles = tgm.tiles_x * tiles_y;
  // x covers the tiles once (fastest-varying in the block schedule), y the slab chunks: blocks that run together
  // work on neighbouring tiles of the SAME few slabs, so HBM is streamed slab by slab in whole DRAM pages instead
  // of 1460 concurrent 13 MB-strided trickles; the 32 B/target records are re-read per chunk from L2, not HBM.
  int gx = tgm.ntiles;
  spc = nslab < APPLY_SLABS_PER_CHUNK ? nslab : APPLY_SLABS_PER_CHUNK;
  int chunks = (nslab + spc - 1) / spc;
  grid = dim3(gx, chunks, 1);
}

static int orca_lastrow_patch(sosie_ctx* ctx, float* dZ2s) {
  const int nx2 = ctx->tg.nx, ny2 = ctx->tg.ny, io = ctx->i_orca_trg;
  cudaStream_t st = ctx->stream;
  auto sect = [&](int l0, int ls, int n, int jl, int r0, int rs, int jr) -> int {
    if (n <= 0) return 0;
    CK(ctx->d_stage2.alloc((size_t)n > ctx->d_stage2.n ? (size_t)n : ctx->d_stage2.n));
    k_sect_read<<<cdiv(n, 256), 256, 0, st>>>(dZ2s, nx2, n, r0, rs, jr, ctx->d_stage2.p); KLAUNCH_CHECK();
    k_sect_write<<<cdiv(n, 256), 256, 0, st>>>(dZ2s, nx2, n, l0, ls, jl, ctx->d_stage2.p); KLAUNCH_CHECK();
    return 0;
  };
  if (io == 4) {  // non-conforming sections: left-hand extent governs
    int nA = nx2 / 2 - 1, nB = nx2 / 2 + 3;
    if (int rc = sect(2, 1, nA, ny2, nx2, -1, ny2 - 2)) return rc;
    if (int rc = sect(nx2, -1, nB, ny2, 2, 1, ny2 - 2)) return rc;
  }
  if (io == 6) {
    int nA = nx2 / 2 - 1;
    if (int rc = sect(2, 1, nA, ny2, nx2 - 1, -1, ny2 - 1)) return rc;
    if (int rc = sect(nx2 - 1, -1, nA, ny2, 2, 1, ny2 - 1)) return rc;
  }
  return 0;
}

int bilin_apply_impl(sosie_ctx* ctx, int nslab, const float* dZ1, float* dZ2, int first_call, int use_clamp, float vmin,
                     float vmax, const int32_t* dmask_trg, float rmiss) {
  // one slab and no records yet (the first call after the mapping was built): pack inside the apply kernel
  const bool fuse = !ctx->pack_ready && nslab == 1 && ctx->map_ready;
  if (!ctx->pack_ready && !fuse) { if (int rc = bilin_pack_impl(ctx)) return rc; }
  if (nslab <= 0) return SOSIE_OK;
  const size_t nt = (size_t)ctx->tg.nx * ctx->tg.ny;
  SlabView v;
  v.nx = ctx->sg.nx; v.ny = ctx->sg.ny; v.nreal = ctx->sg.nx * ctx->sg.ny; v.im = ctx->d_im.p;
  PostOps po;
  po.use_clamp = use_clamp; po.vmin = vmin; po.vmax = vmax; po.mask_trg = dmask_trg; po.rmiss = rmiss;
  dim3 grid;
  int spc;
  TileGeom tgm;
  apply_grid(ctx, nslab, grid, spc, tgm);
  if (nslab >= 2) {
    // batched slabs: shared-memory staged gathers; the per-tile source boxes are cached with the packed records
    if (!ctx->boxes_ready) {
      CK(ctx->d_boxes.alloc(tgm.ntiles));
      k_tilebox<<<grid_for(ctx, (size_t)tgm.ntiles * 256, 256), 256, 0, ctx->stream>>>(ctx->d_pidx.p, v, tgm, ctx->d_boxes.p);
      KLAUNCH_CHECK();
      ctx->boxes_ready = true;
    }
    spc = nslab < APPLY_SLABS_PER_CHUNK_STAGED ? nslab : APPLY_SLABS_PER_CHUNK_STAGED;
    grid = dim3(tgm.ntiles, (nslab + spc - 1) / spc, 1);
    prof_begin(ctx, SOSIE_T_APPLY);
    k_apply_staged<<<grid, 256, 0, ctx->stream>>>(ctx->d_pidx.p, ctx->d_pw.p, nt, v, dZ1, dZ2, nslab, spc, po, tgm, ctx->d_boxes.p);
    KLAUNCH_CHECK();
    prof_end(ctx, SOSIE_T_APPLY);
  } else {
    prof_begin(ctx, SOSIE_T_APPLY);
    if (fuse) {
      if ((size_t)ctx->sg.nx * ctx->sg.nyw > 2147483647ULL) { ctx->err = "source slab too large for 32-bit offsets"; return SOSIE_ERR_ARG; }
      // The very first single-slab apply of a mapping only uses the records (a one-slab job -- a bathymetry -- never needs
      // them again: 32 B per target not written); a second one shows the mapping is reused and stores them on its way.
      const bool store = ctx->unpacked_applies >= 1 || ctx->always_store_records;
      if (store) { CK(ctx->d_pidx.alloc(nt)); CK(ctx->d_pw.alloc(nt)); }
      k_apply1<true><<<grid.x, 256, 0, ctx->stream>>>(nullptr, nullptr, v, dZ1, dZ2, po, tgm, ctx->sg, ctx->tg, ctx->k_ew, ctx->d_metrics.p,
                                                      ctx->d_ab.p, store ? ctx->d_pidx.p : nullptr, store ? ctx->d_pw.p : nullptr);
      KLAUNCH_CHECK();
      if (store) { ctx->pack_ready = true; ctx->pack_gen++; ctx->boxes_ready = false; }
      else ctx->unpacked_applies++;
    } else {
      k_apply1<false><<<grid.x, 256, 0, ctx->stream>>>(ctx->d_pidx.p, ctx->d_pw.p, v, dZ1, dZ2, po, tgm, ctx->sg, ctx->tg, ctx->k_ew, nullptr,
                                                       nullptr, nullptr, nullptr);   // nslab == 1 here
      KLAUNCH_CHECK();
    }
    prof_end(ctx, SOSIE_T_APPLY);
  }
  // l_last_y_row_missing (:241-247) is evaluated on the first call only, the patch (:254-263) on every call
  if (first_call) {
    ctx->l_last_y_row_missing = 0;
    // ORCA last-row test and patch read rows ny2-2 .. ny2 of Z2: on a row-sharded context only the rank whose shard holds
    // all three can evaluate them; a shard that splits them is refused (the other ranks neither test nor patch: their rows
    // do not include the last one)
    bool owns_last = true;
    if (ctx->i_orca_trg >= 4 && ctx->shard_j0 >= 1) {
      const int ny2 = ctx->tg.ny;
      owns_last = ctx->shard_j1 == ny2;
      if (ctx->shard_j1 >= ny2 - 2 && !(ctx->shard_j0 <= ny2 - 2 && ctx->shard_j1 == ny2)) {
        ctx->err = "bilin_apply: with an ORCA target (i_orca_trg >= 4) a shard must hold the last three target rows or none of them";
        return SOSIE_ERR_ARG;
      }
    }
    if (ctx->i_orca_trg >= 4 && owns_last) {
      DBuf<float> dm; CK(dm.alloc(1));
      k_lastrow_mean<<<1, 1, 0, ctx->stream>>>(dZ2, ctx->tg.nx, ctx->tg.ny, dm.p); KLAUNCH_CHECK();
      float hm;
      if (int rc = fetch_small(ctx, dm.p, &hm, sizeof(float))) return rc;
      double rmeanv = (double)hm;
      ctx->l_last_y_row_missing = ((rmeanv < G_RMV + (double)0.1f) && (rmeanv > G_RMV - (double)0.1f)) ? 1 : 0;
    }
  }
  if (ctx->l_last_y_row_missing && (use_clamp || dmask_trg)) {
    // the reference patches the row inside BILIN_2D, BEFORE INTERP_2D's clamp/mask: not fusable
    ctx->err = "bilin_apply_ex: fused post-ops cannot be combined with the ORCA last-row patch (l_last_y_row_missing)";
    return SOSIE_ERR_STATE;
  }
  if (ctx->l_last_y_row_missing) {
    for (int s = 0; s < nslab; s++)
      if (int rc = orca_lastrow_patch(ctx, dZ2 + (size_t)s * nt)) return rc;
  }
  return SOSIE_OK;
}

int bilin_apply_uv_impl(sosie_ctx* ctx, int nslab, const float* dU1, const float* dV1, float* dUo, float* dVo,
                        const double* dcos, const double* dsin) {
  if (ctx->l_last_y_row_missing) {
    // the reference rotates the PATCHED last row (BILIN_2D :254-263, then corr_vect): not fusable, as in apply_ex
    ctx->err = "bilin_apply_uv: the fused rotation cannot be combined with the ORCA last-row patch (l_last_y_row_missing): "
               "use sosie_bilin_apply_dev + sosie_rotate_uv_dev";
    return SOSIE_ERR_STATE;
  }
  if (!ctx->pack_ready) { if (int rc = bilin_pack_impl(ctx)) return rc; }
  if (nslab <= 0) return SOSIE_OK;
  const size_t nt = (size_t)ctx->tg.nx * ctx->tg.ny;
  SlabView v;
  v.nx = ctx->sg.nx; v.ny = ctx->sg.ny; v.nreal = ctx->sg.nx * ctx->sg.ny; v.im = ctx->d_im.p;
  dim3 grid;
  int spc;
  TileGeom tgm;
  apply_grid(ctx, nslab, grid, spc, tgm);
  prof_begin(ctx, SOSIE_T_APPLY);
  k_apply_uv<<<grid, 256, 0, ctx->stream>>>(ctx->d_pidx.p, ctx->d_pw.p, nt, v, dU1, dV1, dUo, dVo, dcos, dsin, nslab, spc, tgm);
  KLAUNCH_CHECK();
  prof_end(ctx, SOSIE_T_APPLY);
  return SOSIE_OK;
}

// ---------------------------------------------------------------------------------------------------
// scalar INTERP_BL for trajectory points (src/mod_bilin_2d.f90:275-361); FZ(i, j) returns Z_in(i,j), 1-based
template <class FZ>
__device__ __forceinline__ float interp_bl_point(int k_ew, int nxi, int nyi, int iP, int jP, int iqd, double xa, double xb, const FZ& Z) {
  int jiPm1 = iP - 1, jiPp1 = iP + 1;
  if ((jiPm1 == 0) && (k_ew >= 0)) jiPm1 = nxi - k_ew;
  if ((jiPp1 == nxi + 1) && (k_ew >= 0)) jiPp1 = 1 + k_ew;
  int i1 = 0, j1 = 0, i2 = 0, j2 = 0, i3 = 0, j3 = 0, i4 = 0, j4 = 0;
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
  float wup = __fadd_rn(__fadd_rn(__fadd_rn(w1, w2), w3), w4);
  if (wup == 0.f) return -9998.f;
  if ((i1 < 1) || (i2 < 1) || (i3 < 1) || (i4 < 1)) return -9997.f;
  if ((j1 < 1) || (j2 < 1) || (j3 < 1) || (j4 < 1)) return -9996.f;
  if ((j1 > nyi) || (j2 > nyi) || (j3 > nyi) || (j4 > nyi)) return -9995.f;
  if ((i1 > nxi) || (i2 > nxi) || (i3 > nxi) || (i4 > nxi)) return -9994.f;
  const float z1 = Z(i1, j1), z2 = Z(i2, j2), z3 = Z(i3, j3), z4 = Z(i4, j4);
  return __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(z1, w1), __fmul_rn(z2, w2)), __fmul_rn(z3, w3)), __fmul_rn(z4, w4)), wup);
}

__global__ void k_interp_bl(int k_ew, int nxi, int nyi, const float* __restrict__ Z, int n, const int32_t* __restrict__ jiP,
                            const int32_t* __restrict__ jjP, const int32_t* __restrict__ iqdv, const double* __restrict__ xav,
                            const double* __restrict__ xbv, float* out) {
  auto fz = [&](int i, int j) { return Z[(size_t)(i - 1) + (size_t)(j - 1) * nxi]; };
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    out[k] = interp_bl_point(k_ew, nxi, nyi, jiP[k], jjP[k], iqdv[k], xav[k], xbv[k], fz);
}

// The per-point block of the trajectory front ends: src/interp_to_ground_track.f90:718-779 (nk = 1,
// linear interpolation in time between the two surrounding model records) and src/interp_to_hydro_section.f90:567-607
// (nk levels of one record, F2 == NULL).  One thread per (level, point).  The reference rebuilds the whole field
// xvar = xvar1 + xdum_r4*(rt - vt_mod(jtm_1)) for EVERY track point; only the five elements a point reads are formed here,
// with the reference's roundings (slope REAL(4) <- REAL(8) quotient; xvar REAL(4) <- REAL(8) sum).
__global__ void __launch_bounds__(256)
k_track_interp(int ni, int nj, int nk, const float* __restrict__ F1, const float* __restrict__ F2, double vt1, double vt2,
               const int8_t* __restrict__ imask, long long n, const double* __restrict__ rt, const int32_t* __restrict__ jiP,
               const int32_t* __restrict__ jjP, const int32_t* __restrict__ iqdv, const double* __restrict__ xav,
               const double* __restrict__ xbv, const double* __restrict__ r_obs_v, double obs_lo, double obs_hi, int clip_missing,
               double rmissval, double* f_bl, double* f_np, int8_t* fmask) {
  const long long tot = n * nk;
  const size_t nij = (size_t)ni * nj;
  const double dvt = vt2 - vt1;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < tot; q += (long long)gridDim.x * blockDim.x) {
    const long long p = q / nk;
    const int jk = (int)(q - p * nk);
    const int iP = jiP[p], jP = jjP[p];
    if (!((iP > 0) && (jP > 0)) || iP > ni || jP > nj) continue;
    const size_t off = (size_t)jk * nij;
    auto M = [&](int i, int j) { return (int)imask[off + (size_t)(i - 1) + (size_t)(j - 1) * ni]; };
    if (M(iP, jP) != 1) continue;
    const double r_obs = r_obs_v[p];
    if (!((r_obs > obs_lo) && (r_obs < obs_hi))) continue;
    const int ip1 = min(iP + 1, ni), jp1 = min(jP + 1, nj), im1 = max(iP - 1, 1), jm1 = max(jP - 1, 1);
    const int idot = M(ip1, jP) + M(ip1, jp1) + M(iP, jp1) + M(im1, jp1) + M(im1, jP) + M(im1, jm1) + M(iP, jm1) + M(ip1, jm1);
    if (idot != 8) continue;
    const double dt = F2 ? __dsub_rn(rt[p], vt1) : 0.0;
    auto xvar = [&](int i, int j) -> float {
      const size_t e = off + (size_t)(i - 1) + (size_t)(j - 1) * ni;
      const float a = F1[e];
      if (!F2) return a;
      const float slope = (float)__ddiv_rn((double)__fsub_rn(F2[e], a), dvt);
      return (float)__dadd_rn((double)a, __dmul_rn((double)slope, dt));
    };
    f_np[q] = (double)xvar(iP, jP);
    double v = (double)interp_bl_point(-1, ni, nj, iP, jP, iqdv[p], xav[p], xbv[p], xvar);
    if (clip_missing && (v < (double)-9990.f)) v = rmissval;
    f_bl[q] = v;
    fmask[q] = 1;
  }
}

int track_interp_impl(sosie_ctx* ctx, int ni, int nj, int nk, const float* F1, const float* F2, double vt1, double vt2, const int8_t* imask,
                      long long n, const double* rt, const int32_t* jiP, const int32_t* jjP, const int32_t* iqd, const double* xa,
                      const double* xb, const double* r_obs, double obs_lo, double obs_hi, int clip_missing, double rmissval, double* f_bl,
                      double* f_np, int8_t* fmask) {
  if (n <= 0) return SOSIE_OK;
  if (ni < 1 || nj < 1 || nk < 1 || !F1 || !imask || !jiP || !jjP || !iqd || !xa || !xb || !r_obs || !f_bl || !f_np || !fmask || (F2 && !rt)) {
    ctx->err = "sosie_track_interp: bad arguments"; return SOSIE_ERR_ARG;
  }
  if (F2 && !(vt2 != vt1)) { ctx->err = "sosie_track_interp: the two model records have the same time"; return SOSIE_ERR_ARG; }
  k_track_interp<<<grid_for(ctx, (size_t)n * nk, 256), 256, 0, ctx->stream>>>(ni, nj, nk, F1, F2, vt1, vt2, imask, n, rt, jiP, jjP, iqd, xa, xb,
                                                                                 r_obs, obs_lo, obs_hi, clip_missing, rmissval, f_bl, f_np, fmask);
  KLAUNCH_CHECK();
  return SOSIE_OK;
}

int interp_bl_impl(sosie_ctx* ctx, int k_ew, int nxi, int nyi, const float* dZ, int n, const int32_t* ji, const int32_t* jj,
                   const int32_t* iq, const double* xa, const double* xb, float* out) {
  if (n <= 0) return SOSIE_OK;
  k_interp_bl<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(k_ew, nxi, nyi, dZ, n, ji, jj, iq, xa, xb, out);
  KLAUNCH_CHECK();
  return SOSIE_OK;
}
