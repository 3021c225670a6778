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
#include <algorithm>

#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no -lcuda)

#include "pack.cuh"

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pack(SrcGeom sg, TrgGeom tg, int k_ew, const int32_t* __restrict__ MT, const double* __restrict__ AB, const float* __restrict__ DNP,
       int4* pidx, float4* pw, size_t k0, size_t kcnt) {
  const size_t nt = (size_t)tg.nx * tg.ny;
  for (size_t kk = blockIdx.x * (size_t)blockDim.x + threadIdx.x; kk < kcnt; kk += (size_t)gridDim.x * blockDim.x) {
    const size_t k = k0 + kk;
    int jj = (int)(k / tg.nx) + 1, ji = (int)(k % tg.nx) + 1;
    int4 id;
    float4 w;
    const bool far = DNP[k] > PACK_FAR_KM;   // (false for NaN)
    double lon_t = 0.0, lat_t = 0.0;
    if (!far) { lon_t = trg_lon(tg, ji, jj); lat_t = trg_lat(tg, ji, jj); }
    pack_record(sg, k_ew, MT[k], MT[k + nt], MT[k + 2 * nt], AB[k], AB[k + nt], lon_t, lat_t, id, w, far);
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
  FastDiv dtx;   // division by tiles_x (the single-slab kernel splits its tile number once per thread: a fifth of its set-up)
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
         const float* __restrict__ DNP, int4* pidx_out, float4* pw_out) {
  const size_t nt = FUSE ? (size_t)tg.nx * tg.ny : 0;
  for (int tile = blockIdx.x; tile < tgm.ntiles; tile += gridDim.x) {
    const int tj = (int)fdiv((unsigned)tile, tgm.dtx), ti = tile - tj * tgm.tiles_x;
    const int i = ti * 32 + (threadIdx.x & 31), j = tgm.j0 + tj * 8 + (threadIdx.x >> 5);  // 0-based
    if (i >= tgm.nx || j > tgm.j1) continue;
    const size_t k = (size_t)i + (size_t)j * tgm.nx;
    int4 id;
    float4 w;
    if (FUSE) {
      const bool far = __ldcs(DNP + k) > PACK_FAR_KM;   // (false for NaN): no identical-point copy possible, coordinates not read
      double lon_t = 0.0, lat_t = 0.0;
      if (!far) { lon_t = trg_lon(tg, i + 1, j + 1); lat_t = trg_lat(tg, i + 1, j + 1); }
      pack_record(sg, k_ew, __ldcs(MT + k), __ldcs(MT + k + nt), __ldcs(MT + k + 2 * nt), __ldcs(AB + k), __ldcs(AB + k + nt), lon_t, lat_t, id, w, far);
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

// ---- TMA staging (tiles whose source box fits TMA_NX sub-boxes of TMA_BW x TMA_BH cells) -----------------------------
// A 3-D tensor map describes the slab stack Z1(nx, ny, nslab); ONE cp.async.bulk.tensor instruction, issued by one thread,
// brings the box {TMA_BW, TMA_BH, TMA_SD} at the tile's box origin -- TMA_SD consecutive slabs at once -- into shared
// memory and signals an mbarrier; cells beyond the slab edges are zero-filled by the hardware and never read.  The
// sub-box is 32 floats = 128 B wide, the span of SWIZZLE_128B: the 16-byte chunk index is XOR-ed with the row index mod 8,
// so the lanes of a warp whose corners share a column but sit in neighbouring rows (a target row cutting obliquely through
// the source grid: the polar-stereographic target of config D) hit different banks.
#define TMA_BW 32
#define TMA_BH 16
#define TMA_SD 4
#define TMA_NX 2
#define TMA_NSTG 3
#define TMA_PLANE (TMA_BW * TMA_BH)                 // floats of one slab of one sub-box
#define TMA_SUB (TMA_PLANE * TMA_SD)                // floats one TMA instruction writes (8 KB)
#define TMA_STAGE (TMA_SUB * TMA_NX)                // floats per pipeline stage (16 KB)
#define TMA_SMEM_BYTES (TMA_STAGE * TMA_NSTG * 4 + 1024)

__device__ __forceinline__ bool box_is_tma(const int4& box) { return box.z > 0 && box.z <= TMA_NX * TMA_BW && box.w <= TMA_BH; }

// ---- group staging with 16-byte cp.async (k_apply_grp) ---------------------------------------------------------------
// The box is widened to 16-byte boundaries of the source rows (its first column rounded down to a multiple of 4: the source
// width must be a multiple of 4 and the slab stack 16-byte aligned), so a slab's box is h rows of `cpr` 16-byte chunks; a
// GROUP of GRP_SD consecutive slabs is staged at once, each thread copying at most GRP_CPT chunks with cp.async.cg 16, and
// the block synchronises once per group instead of once per slab.  Row pitch in shared memory: cpr chunks, plus one when
// cpr is even, so that the same column of neighbouring rows falls into different banks.
#ifndef GRP_NSTG
#define GRP_NSTG 4
#endif
#ifndef GRP_MINB
#define GRP_MINB 4                                      // resident blocks per SM the kernel is compiled for
#endif
#define GRP_STAGE 3072                                  // floats per pipeline stage: SD slabs x (GRP_STAGE / SD) floats of box
#define GRP_CPT (GRP_STAGE / 4 / 256)                   // 16-byte chunks per thread and group at most (3)
// The kernel is instantiated for groups of SD = 8, 4, 2 and 1 slabs: a tile takes the deepest group its (padded) box allows --
// 384, 768, 1536 or 3072 floats per slab (the wide, short boxes next to the pole of a polar-stereographic target are the big ones)
__host__ __device__ __forceinline__ int grp_cpr(const int4& box) { return ((box.x & 3) + box.z + 3) >> 2; }
__host__ __device__ __forceinline__ int grp_pitch(const int4& box) { const int c = grp_cpr(box); return 4 * (c + ((c & 1) ? 0 : 1)); }
__device__ __forceinline__ int grp_class(const int4& box) {   // slabs per group, 0: the box fits no instantiation
  if (box.z <= 0) return 0;
  const long long need = (long long)box.w * grp_pitch(box);
  return need <= GRP_STAGE / 8 ? 8 : need <= GRP_STAGE / 4 ? 4 : need <= GRP_STAGE / 2 ? 2 : need <= GRP_STAGE ? 1 : 0;
}
__device__ __forceinline__ bool box_fits_grp(const int4& box) { return grp_class(box) != 0; }
// which kernel of a batched apply takes a tile: 0 the 4-byte cp.async kernel (also direct gathers), 1 k_apply_grp,
// 2 k_apply_tma.  flags bit 0: the group kernel runs, bit 1: the TMA kernel runs
#define OWN_OLD 0
#define OWN_GRP 1
#define OWN_TMA 2
__device__ __forceinline__ int tile_owner(const int4& box, int flags) {
  if ((flags & 2) && box_is_tma(box)) return OWN_TMA;
  if ((flags & 1) && box_fits_grp(box)) return OWN_GRP;
  return OWN_OLD;
}

// one block per tile: bounding box (0-based i0, j0, w, h) of the tile's source cells; w == 0 -> not stageable.
// counts[0..2]: tiles that fit the TMA box / the group-staged kernel / neither; counts[3]: tiles the group kernel does not
// take, counts[4..6]: tiles of its instantiations with 8 or 4 / 2 / 1 slabs per group; lists: four compact tile lists of ntiles
// entries each (groups of 8 or 4, of 2, of 1, the rest), so that every kernel is launched over its own tiles only
__global__ void __launch_bounds__(256)
k_tilebox(const int4* __restrict__ pidx, SlabView v, TileGeom tgm, int4* boxes, int* counts, int* lists) {
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
      if (!shbad[0] && sh[1][0] >= 0 && (long long)w * h <= 65536) {
        const int4 c = make_int4(sh[0][0], sh[2][0], w, h);
        if ((long long)w * h <= BOXMAX || box_is_tma(c) || box_fits_grp(c)) b = c;
      }
      boxes[tile] = b;
      if (box_is_tma(b)) atomicAdd(&counts[0], 1);
      const int gc = grp_class(b), li = gc >= 4 ? 0 : gc == 2 ? 1 : gc == 1 ? 2 : 3;
      if (gc) atomicAdd(&counts[1], 1);
      lists[(size_t)li * tgm.ntiles + atomicAdd(&counts[li == 3 ? 3 : 4 + li], 1)] = tile;
      if (!box_is_tma(b) && !box_fits_grp(b)) atomicAdd(&counts[2], 1);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256, 4)
k_apply_staged(const int4* __restrict__ pidx, const float4* __restrict__ pw, size_t nt, SlabView v, const float* __restrict__ Z1,
               float* __restrict__ Z2, int nslab, int slabs_per_chunk, PostOps po, TileGeom tgm, const int4* __restrict__ boxes,
               int own_flags, const int* __restrict__ tile_list, int nlist) {
  __shared__ float buf[NSTAGE][BOXMAX];
  const size_t src_stride = (size_t)v.nreal;
  const int s0 = blockIdx.y * slabs_per_chunk;
  const int s1 = min(s0 + slabs_per_chunk, nslab);
  const int nsl = s1 - s0;
  for (int tl = blockIdx.x; tl < nlist; tl += gridDim.x) {
    const int tile = tile_list ? tile_list[tl] : tl;
    const int4 box = boxes[tile];
    if (tile_owner(box, own_flags) != OWN_OLD) continue;   // another kernel of this apply takes the tile (block-uniform)
    const int ti = tile % tgm.tiles_x, tj = tile / tgm.tiles_x;
    const int i = ti * 32 + (threadIdx.x & 31), j = tgm.j0 + tj * 8 + (threadIdx.x >> 5);
    const bool valid = (i < tgm.nx && j <= tgm.j1);
    const size_t k = valid ? (size_t)i + (size_t)j * tgm.nx : 0;
    int4 id = make_int4(REC_CONST, 0, 0, 0);
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) { id = pidx[k]; w = pw[k]; }
    const float wup = __fadd_rn(__fadd_rn(__fadd_rn(w.x, w.y), w.z), w.w);
    const bool masked = (valid && po.mask_trg) ? (po.mask_trg[k] == 0) : false;
    if (box.z == 0 || (long long)box.z * box.w > BOXMAX) {  // not stageable (special records, pole rows, E-W wrap, huge footprint): direct gathers
      if (valid) {
        int s = s0;
        if (id.x >= 0 && max(max(id.x, id.y), max(id.z, id.w)) < v.nreal) {
          // ordinary record, no pole rows: the 16 gathers of four slabs are in flight together (the loop below is a chain of
          // dependent latencies per slab; these tiles -- wide boxes next to the pole -- are few but were 4x slower per point)
          for (; s + 4 <= s1; s += 4) {
            float z[4][4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const float* Z = Z1 + (size_t)(s + u) * src_stride;
              z[u][0] = __ldg(Z + id.x); z[u][1] = __ldg(Z + id.y); z[u][2] = __ldg(Z + id.z); z[u][3] = __ldg(Z + id.w);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
              float r = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(z[u][0], w.x), __fmul_rn(z[u][1], w.y)), __fmul_rn(z[u][2], w.z)), __fmul_rn(z[u][3], w.w)), wup);
              if (po.use_clamp) { r = fmaxf(r, po.vmin); r = fminf(r, po.vmax); }
              if (masked) r = po.rmiss;
              __stcs(Z2 + (size_t)(s + u) * nt + k, r);
            }
          }
        }
        for (; s < s1; s++) {
          float r = interp_one(Z1 + (size_t)s * src_stride, v, id, w, wup);
          if (po.use_clamp) { r = fmaxf(r, po.vmin); r = fminf(r, po.vmax); }
          if (masked) r = po.rmiss;
          __stcs(Z2 + (size_t)s * nt + k, r);
        }
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

// ---- mbarrier / TMA primitives (PTX) ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// spin until the phase with parity `par` of `bar` has completed; a transfer that never lands (a descriptor the hardware
// rejects) would hang the GPU: after ~2^28 probes the kernel traps instead, which the host sees as a launch failure
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned par) {
  const unsigned a = smem_u32(bar);
  unsigned done = 0;
  for (unsigned spin = 0; !done; spin++) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(a), "r"(par) : "memory");
    if (spin > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
               ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// x / wup, correctly rounded, through the reciprocal rcp = RN(1/wup) computed once per target: q = RN(x*rcp) is a faithful
// quotient, r = RN(x - q*wup) is exact (one FMA), RN(q + r*rcp) is the correctly rounded quotient (Markstein's theorem, no
// over/underflow: guarded by the range test; callers check 0.5 <= wup <= 2).  Zeros keep their sign through q alone.
// Outside the guarded range the IEEE division runs.  3 instructions instead of the ~12 of __fdiv_rn with its slow path;
// tests: sosie_selftest_div (2^28 random and edge operands against __fdiv_rn) and every apply parity test.
__device__ __forceinline__ float div_by_rcp(float x, float wup, float rcp) {
  const float ax = fabsf(x);
  const float q = __fmul_rn(x, rcp);
  if (ax >= 0x1p-90f && ax <= 0x1p90f) {
    const float r = __fmaf_rn(-q, wup, x);
    return __fmaf_rn(r, rcp, q);
  }
  if (ax == 0.f) return q;
  return __fdiv_rn(x, wup);
}

__global__ void k_selftest_div(unsigned long long n, unsigned long long seed, unsigned long long* bad) {
  unsigned long long nb = 0;
  for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < n; k += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned long long z = (k + seed) * 0x9E3779B97F4A7C15ULL;   // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
    const unsigned xb = (unsigned)z, wb = (unsigned)(z >> 32);
    float x = __uint_as_float(xb);                                              // any bit pattern (NaN, Inf, denormals, zeros included)
    float wup = __uint_as_float(0x3F000000u + (wb % 0x01000001u));              // [0.5, 2.0] inclusive, every REAL(4) in it reachable
    const unsigned sel = (unsigned)(k & 7), e0 = sel == 0 ? 27u : sel == 1 ? 207u : 117u;   // around the guards 2^-90, 2^90 and around 1
    if (sel < 3) x = __uint_as_float((xb & 0x807FFFFFu) | ((e0 + (xb >> 23) % 20u) << 23));
    const float a = div_by_rcp(x, wup, __frcp_rn(wup)), b = __fdiv_rn(x, wup);
    if (__float_as_uint(a) != __float_as_uint(b) && !(a != a && b != b)) nb++;
  }
  if (nb) atomicAdd(bad, nb);
}

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc));
}


// Batched slabs, group-staged (see GRP_* above).  Per tile and chunk of slabs: groups of GRP_SD slabs travel through
// GRP_NSTG stages; per group a thread issues its (at most GRP_CPT, usually one) 16-byte copies, the block waits and
// synchronises ONCE, and every thread interpolates its target in the GRP_SD slabs from shared memory with compile-time plane
// offsets.  Instruction diet (ncu r02a: the first version was issue bound at 57 instructions per output, 20 % of them tile
// set-up): the set-up is amortised over a long run of slabs (the host sizes the chunks for ~8 waves of blocks, not 32
// slabs), its divisions go through multiply-high reciprocals, the kernel is specialised on the presence of post-ops, full
// groups run without per-slab predicates, and the quotient by the weight sum is three instructions (div_by_rcp).
// the quotient outside the guarded range of the three-instruction form (zeros, denormals, huge values, Inf, NaN, or a weight sum
// outside [0.5, 2]): rare, and kept out of line -- inlined, its FCHK / MUFU / fix-up call sequence was 2/3 of the kernel's code
static __device__ __noinline__ float quot_general(float x, float wup, float rcp, bool fast) {
  return fast ? div_by_rcp(x, wup, rcp) : __fdiv_rn(x, wup);
}
// SD slabs, PLANE floats apart, starting at b (a compile-time offset into the stage buffers once the caller's loops are unrolled:
// the four corner loads are LDS [corner register + immediate])
template <int SD, int PLANE, bool FULL, bool POST>
__device__ __forceinline__ void grp_compute(const float* __restrict__ b, int left, int o1, int o2, int o3, int o4, const float4& w, float wup, float rcp,
                                            bool fast, bool masked, const PostOps& po, float*& outp, size_t nt) {
  float acc[SD];
#pragma unroll
  for (int d = 0; d < SD; d++) {
    const float* bd = b + d * PLANE;
    acc[d] = (FULL || d < left) ? __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(bd[o1], w.x), __fmul_rn(bd[o2], w.y)), __fmul_rn(bd[o3], w.z)), __fmul_rn(bd[o4], w.w)) : 1.0f;
  }
  // one range test for the group: every |acc| inside the guarded range of div_by_rcp.  The smallest magnitude through
  // fminf (which drops a NaN), the largest through the SUM of the magnitudes (a NaN or an Inf poisons it, and a sum below
  // 2^90 bounds every term): zeros, denormals, huge values, Inf and NaN all fail the test and take the general path
  float lo = fabsf(acc[0]), sum = lo;
#pragma unroll
  for (int d = 1; d < SD; d++) { lo = fminf(lo, fabsf(acc[d])); sum = __fadd_rn(sum, fabsf(acc[d])); }
  const bool all_fast = fast && lo >= 0x1p-90f && sum <= 0x1p90f;
  // ONE branch per batch (not per slab): the general quotients of a batch that fails the test replace the accumulators
  if (!all_fast) {
#pragma unroll
    for (int d = 0; d < SD; d++) acc[d] = quot_general(acc[d], wup, rcp, fast);
  }
#pragma unroll
  for (int d = 0; d < SD; d++) {
    if (FULL || d < left) {
      float r = acc[d];
      if (all_fast) {
        const float q = __fmul_rn(acc[d], rcp);
        r = __fmaf_rn(__fmaf_rn(-q, wup, acc[d]), rcp, q);
      }
      if (POST) {
        if (po.use_clamp) { r = fmaxf(r, po.vmin); r = fminf(r, po.vmax); }
        if (masked) r = po.rmiss;
      }
      __stcs(outp + (size_t)d * nt, r);
    }
  }
  outp += (size_t)SD * nt;
}

// one tile, groups of SD slabs (block-uniform call)
template <int SD, bool POST>
__device__ __forceinline__ void grp_tile(float (*buf)[GRP_STAGE], const int4& box, int tile, const int4* __restrict__ pidx, const float4* __restrict__ pw,
                                         size_t nt, const SlabView& v, const FastDiv& dnx, const float* __restrict__ Z1, float* __restrict__ Z2,
                                         int s0, int s1, const PostOps& po, const TileGeom& tgm) {
  constexpr int PLANE = GRP_STAGE / SD;
  const size_t src_stride = (size_t)v.nreal;
  const int ngrp = (s1 - s0 + SD - 1) / SD, nfull = (s1 - s0) / SD;
  const int tj = tile / tgm.tiles_x, ti = tile - tj * tgm.tiles_x;
  const int i = ti * 32 + (threadIdx.x & 31), j = tgm.j0 + tj * 8 + (threadIdx.x >> 5);
  const bool valid = (i < tgm.nx && j <= tgm.j1);
  const size_t k = valid ? (size_t)i + (size_t)j * tgm.nx : 0;
  int4 id = make_int4(0, 0, 0, 0);
  float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) { id = pidx[k]; w = pw[k]; }
  const float wup = __fadd_rn(__fadd_rn(__fadd_rn(w.x, w.y), w.z), w.w);
  const bool fast = wup >= 0.5f && wup <= 2.0f;
  const float rcp = fast ? __frcp_rn(wup) : 0.f;
  const bool masked = (POST && valid && po.mask_trg) ? (po.mask_trg[k] == 0) : false;
  const int x0a = box.x & ~3, cpr = grp_cpr(box), pitch = grp_pitch(box), nch = cpr * box.w;   // chunks per slab
  auto off = [&](int lin) -> int {
    const int r = (int)fdiv((unsigned)lin, dnx);
    return (r - box.y) * pitch + (lin - r * v.nx - x0a);
  };
  int o1 = 0, o2 = 0, o3 = 0, o4 = 0;
  if (valid) { o1 = off(id.x); o2 = off(id.y); o3 = off(id.z); o4 = off(id.w); }
  // this thread's chunks of a group: slab d of the group, chunk c of the slab's box; ncp = chunks per thread in use
  // (block-uniform).  e < 768: the quotients by nch and cpr come exactly from float reciprocals
  const int ncp = (nch * SD + 255) >> 8;
  const float rnch = 1.0f / (float)nch, rcpr = 1.0f / (float)cpr;
  const float* gsrc[GRP_CPT];
  int soff[GRP_CPT], sd_[GRP_CPT];
#pragma unroll
  for (int m = 0; m < GRP_CPT; m++) {
    const int e = threadIdx.x + 256 * m;
    const bool ok = e < nch * SD;
    const int d = ok ? (int)(((float)e + 0.5f) * rnch) : 0;
    const int c = ok ? e - d * nch : 0;
    const int row = (int)(((float)c + 0.5f) * rcpr);
    const int cc = c - row * cpr;
    sd_[m] = ok ? d : SD;    // SD: no chunk
    soff[m] = d * PLANE + row * pitch + cc * 4;
    gsrc[m] = Z1 + (size_t)(s0 + d) * src_stride + (size_t)(box.y + row) * v.nx + x0a + cc * 4;
  }
  float* outp = Z2 + (size_t)s0 * nt + k;
  auto issue = [&](float* sb, int g, bool full) {   // group g -> the stage at sb
    const int left = full ? SD : s1 - (s0 + g * SD);   // slabs of this group that exist
#pragma unroll
    for (int m = 0; m < GRP_CPT; m++) {
      if (m < ncp) {
        if (sd_[m] < left) cp_async16(sb + soff[m], gsrc[m]);
        gsrc[m] += (size_t)SD * src_stride;
      }
    }
  };
#pragma unroll
  for (int q = 0; q < GRP_NSTG - 1; q++) {
    if (q < ngrp) issue(&buf[q][0], q, q < nfull);
    cp_async_commit();
  }
  // a group of more than 4 slabs is interpolated 4 at a time (the accumulators of a batch live in registers)
  constexpr int SC = SD > 4 ? 4 : SD, NH = SD / SC;
  // GRP_NSTG groups per trip, so that every stage address below is a compile-time constant
  int g = 0;
  bool more = ngrp > 0;
  while (more) {
#pragma unroll
    for (int st = 0; st < GRP_NSTG; st++, g++) {
      if (g >= ngrp) { more = false; break; }   // (block-uniform)
      cp_async_wait<GRP_NSTG - 2>();   // this thread's copies of group g have landed
      __syncthreads();                 // ... everybody's have, and everybody is done reading the stage of group g-1
      const int gn = g + GRP_NSTG - 1;
      if (gn < ngrp) issue(&buf[(st + GRP_NSTG - 1) % GRP_NSTG][0], gn, gn < nfull);   // refills the stage group g-1 used
      cp_async_commit();
      if (valid) {
        const float* b = &buf[st][0];
        if (g < nfull) {
#pragma unroll
          for (int h = 0; h < NH; h++) grp_compute<SC, PLANE, true, POST>(b + h * SC * PLANE, SC, o1, o2, o3, o4, w, wup, rcp, fast, masked, po, outp, nt);
        } else {
          int left = s1 - (s0 + g * SD);
#pragma unroll
          for (int h = 0; h < NH; h++, left -= SC)
            if (left > 0) grp_compute<SC, PLANE, false, POST>(b + h * SC * PLANE, left, o1, o2, o3, o4, w, wup, rcp, fast, masked, po, outp, nt);
        }
      }
    }
  }
  __syncthreads();   // the next tile's prologue overwrites the stages
  cp_async_wait<0>();
}

// a tile whose box fits no instantiation (special records, pole rows, E-W wrap, huge footprint): direct gathers
template <bool POST>
__device__ __forceinline__ void direct_tile(int tile, const int4* __restrict__ pidx, const float4* __restrict__ pw, size_t nt, const SlabView& v,
                                            const float* __restrict__ Z1, float* __restrict__ Z2, int s0, int s1, const PostOps& po, const TileGeom& tgm) {
  const size_t src_stride = (size_t)v.nreal;
  const int tj = tile / tgm.tiles_x, ti = tile - tj * tgm.tiles_x;
  const int i = ti * 32 + (threadIdx.x & 31), j = tgm.j0 + tj * 8 + (threadIdx.x >> 5);
  if (!(i < tgm.nx && j <= tgm.j1)) return;
  const size_t k = (size_t)i + (size_t)j * tgm.nx;
  const int4 id = pidx[k];
  const float4 w = pw[k];
  const float wup = __fadd_rn(__fadd_rn(__fadd_rn(w.x, w.y), w.z), w.w);
  const bool masked = (POST && po.mask_trg) ? (po.mask_trg[k] == 0) : false;
  auto fin = [&](float r, int s) {
    if (POST) {
      if (po.use_clamp) { r = fmaxf(r, po.vmin); r = fminf(r, po.vmax); }
      if (masked) r = po.rmiss;
    }
    __stcs(Z2 + (size_t)s * nt + k, r);
  };
  int s = s0;
  if (id.x >= 0 && max(max(id.x, id.y), max(id.z, id.w)) < v.nreal) {
    // ordinary record, no pole rows: the 16 gathers of four slabs are in flight together
    for (; s + 4 <= s1; s += 4) {
      float z[4][4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const float* Z = Z1 + (size_t)(s + u) * src_stride;
        z[u][0] = __ldg(Z + id.x); z[u][1] = __ldg(Z + id.y); z[u][2] = __ldg(Z + id.z); z[u][3] = __ldg(Z + id.w);
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
        fin(__fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(z[u][0], w.x), __fmul_rn(z[u][1], w.y)), __fmul_rn(z[u][2], w.z)), __fmul_rn(z[u][3], w.w)), wup), s + u);
    }
  }
  for (; s < s1; s++) fin(interp_one(Z1 + (size_t)s * src_stride, v, id, w, wup), s);
}

// The batched apply of a 16-byte aligned slab stack in ONE launch.  blockIdx.x walks the compact tile lists in the order
// "direct" tiles, groups of 1, of 2, of 4 or 8 slabs: the slow tiles (huge or odd boxes) start first and the many cheap ones fill
// the tail of the launch.  cnt = lengths of the four lists (8 or 4 / 2 / 1 slabs per group, direct), lists = k_tilebox's.
// tile_list == nullptr (the TMA kernel also runs): every tile is visited and the owner test decides.
template <bool POST>
__global__ void __launch_bounds__(256, GRP_MINB)
k_apply_grp(const int4* __restrict__ pidx, const float4* __restrict__ pw, size_t nt, SlabView v, FastDiv dnx, const float* __restrict__ Z1,
            float* __restrict__ Z2, int nslab, int slabs_per_chunk, const __grid_constant__ PostOps po, TileGeom tgm, const int4* __restrict__ boxes,
            int own_flags, const int* __restrict__ lists, int4 cnt) {
  __shared__ __align__(16) float buf[GRP_NSTG][GRP_STAGE];
  const int s0 = blockIdx.y * slabs_per_chunk;
  const int s1 = min(s0 + slabs_per_chunk, nslab);
  const int nlist = lists ? cnt.x + cnt.y + cnt.z + cnt.w : tgm.ntiles;
  for (int tl = blockIdx.x; tl < nlist; tl += gridDim.x) {
    int tile = tl;
    if (lists) {
      int t = tl;
      if (t < cnt.w) tile = lists[3 * (size_t)tgm.ntiles + t];
      else if ((t -= cnt.w) < cnt.z) tile = lists[2 * (size_t)tgm.ntiles + t];
      else if ((t -= cnt.z) < cnt.y) tile = lists[(size_t)tgm.ntiles + t];
      else tile = lists[t - cnt.y];
    }
    const int4 box = boxes[tile];
    if (tile_owner(box, own_flags) == OWN_TMA) continue;   // block-uniform
    const int cls = grp_class(box);
    if (cls == 8) grp_tile<8, POST>(buf, box, tile, pidx, pw, nt, v, dnx, Z1, Z2, s0, s1, po, tgm);
    else if (cls == 4) grp_tile<4, POST>(buf, box, tile, pidx, pw, nt, v, dnx, Z1, Z2, s0, s1, po, tgm);
    else if (cls == 2) grp_tile<2, POST>(buf, box, tile, pidx, pw, nt, v, dnx, Z1, Z2, s0, s1, po, tgm);
    else if (cls == 1) grp_tile<1, POST>(buf, box, tile, pidx, pw, nt, v, dnx, Z1, Z2, s0, s1, po, tgm);
    else direct_tile<POST>(tile, pidx, pw, nt, v, Z1, Z2, s0, s1, po, tgm);
  }
}

// Batched slabs, TMA-staged.  Pipeline per tile and chunk of slabs: groups of TMA_SD slabs travel through TMA_NSTG stages;
// thread 0 arms the stage's mbarrier with the byte count and issues one TMA instruction per sub-box; every thread waits on
// the barrier, interpolates its target in the TMA_SD slabs of the group from shared memory (compile-time plane offsets,
// swizzled corner offsets computed once per tile) and stores; a block barrier frees the stage for the group TMA_NSTG ahead.
__global__ void __launch_bounds__(256, 4)
k_apply_tma(const __grid_constant__ CUtensorMap tm, const int4* __restrict__ pidx, const float4* __restrict__ pw, size_t nt, SlabView v,
            float* __restrict__ Z2, int nslab, int slabs_per_chunk, PostOps po, TileGeom tgm, const int4* __restrict__ boxes, int own_flags) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ uint64_t full[TMA_NSTG];
  float* const stg = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);   // swizzle atoms: 1024-byte aligned
  const int s0 = blockIdx.y * slabs_per_chunk;
  const int s1 = min(s0 + slabs_per_chunk, nslab);
  const int ngrp = (s1 - s0 + TMA_SD - 1) / TMA_SD;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int st = 0; st < TMA_NSTG; st++) mbar_init(&full[st], 1);
    mbar_fence_init();
  }
  __syncthreads();
  unsigned phase = 0;   // bit st: parity of the next completion of full[st] (identical in every thread)
  for (int tile = blockIdx.x; tile < tgm.ntiles; tile += gridDim.x) {
    const int4 box = boxes[tile];
    if (tile_owner(box, own_flags) != OWN_TMA) continue;   // block-uniform
    const int nsub = box.z <= TMA_BW ? 1 : 2;
    const int ti = tile % tgm.tiles_x, tj = tile / tgm.tiles_x;
    const int i = ti * 32 + (threadIdx.x & 31), j = tgm.j0 + tj * 8 + (threadIdx.x >> 5);
    const bool valid = (i < tgm.nx && j <= tgm.j1);
    const size_t k = valid ? (size_t)i + (size_t)j * tgm.nx : 0;
    int4 id = make_int4(0, 0, 0, 0);
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) { id = pidx[k]; w = pw[k]; }
    const float wup = __fadd_rn(__fadd_rn(__fadd_rn(w.x, w.y), w.z), w.w);
    const bool fast = wup >= 0.5f && wup <= 2.0f;
    const float rcp = fast ? __frcp_rn(wup) : 0.f;
    const bool masked = (valid && po.mask_trg) ? (po.mask_trg[k] == 0) : false;
    // corner -> float offset inside one slab plane pair: sub-box, row, swizzled column
    auto off = [&](int lin) -> int {
      const int c = lin % v.nx - box.x, r = lin / v.nx - box.y;
      const int sub = c >> 5, cc = c & 31;
      return sub * TMA_SUB + r * TMA_BW + ((((cc >> 2) ^ (r & 7)) << 2) | (cc & 3));
    };
    int o1 = 0, o2 = 0, o3 = 0, o4 = 0;
    if (valid) { o1 = off(id.x); o2 = off(id.y); o3 = off(id.z); o4 = off(id.w); }
    float* outp = Z2 + (size_t)s0 * nt + k;
    auto issue = [&](int g) {
      const int st = g % TMA_NSTG;
      mbar_expect_tx(&full[st], (unsigned)(nsub * TMA_SUB * 4));
      for (int sx = 0; sx < nsub; sx++)
        tma_load_3d(stg + st * TMA_STAGE + sx * TMA_SUB, &tm, box.x + sx * TMA_BW, box.y, s0 + g * TMA_SD, &full[st]);
    };
    if (threadIdx.x == 0)
      for (int g = 0; g < TMA_NSTG && g < ngrp; g++) issue(g);
    for (int g = 0; g < ngrp; g++) {
      const int st = g % TMA_NSTG;
      mbar_wait(&full[st], (phase >> st) & 1u);
      phase ^= 1u << st;
      const float* b = stg + st * TMA_STAGE;
      if (valid) {
        const int sb = s0 + g * TMA_SD;
#pragma unroll
        for (int d = 0; d < TMA_SD; d++) {
          if (sb + d < s1) {
            const float* bd = b + d * TMA_PLANE;
            const float acc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(bd[o1], w.x), __fmul_rn(bd[o2], w.y)), __fmul_rn(bd[o3], w.z)), __fmul_rn(bd[o4], w.w));
            float r = fast ? div_by_rcp(acc, wup, rcp) : __fdiv_rn(acc, wup);
            if (po.use_clamp) { r = fmaxf(r, po.vmin); r = fminf(r, po.vmax); }
            if (masked) r = po.rmiss;
            __stcs(outp, r);
            outp += nt;
          }
        }
      }
      __syncthreads();   // everybody is done reading this stage
      if (threadIdx.x == 0 && g + TMA_NSTG < ngrp) issue(g + TMA_NSTG);
    }
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
  k_pack<<<grid_for(ctx, kcnt, 256), 256, 0, ctx->stream>>>(ctx->sg, ctx->tg, ctx->k_ew, ctx->d_metrics.p, ctx->d_ab.p, ctx->d_dnp.p, ctx->d_pidx.p, ctx->d_pw.p, k0, kcnt);
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

// 3-D tensor map of the slab stack Z1(nx, ny, nslab) REAL(4) with box {TMA_BW, TMA_BH, TMA_SD}, SWIZZLE_128B, zero fill
// outside the slab.  cuTensorMapEncodeTiled is a driver entry point: resolved at run time through the runtime API.
static bool apply_tensor_map(CUtensorMap* tm, const float* dZ1, int nx, int ny, int nslab) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn enc = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess) enc = (encode_fn)fn;
    else cudaGetLastError();
  }
  if (!enc) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nslab};
  const cuuint64_t strides[2] = {(cuuint64_t)nx * 4, (cuuint64_t)nx * ny * 4};
  const cuuint32_t box[3] = {TMA_BW, TMA_BH, TMA_SD};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(dZ1), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int selftest_div_impl(sosie_ctx* ctx, unsigned long long n, unsigned long long seed, unsigned long long* mismatches) {
  DBuf<unsigned long long> d;
  CK(d.alloc(1));
  CK(cudaMemsetAsync(d.p, 0, sizeof(unsigned long long), ctx->stream));
  k_selftest_div<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(n, seed, d.p);
  KLAUNCH_CHECK();
  return fetch_small(ctx, d.p, mismatches, sizeof(unsigned long long));
}

static void apply_grid(const sosie_ctx* ctx, int nslab, dim3& grid, int& spc, TileGeom& tgm) {
  size_t k0, kcnt;
  shard_range(ctx, &k0, &kcnt);
  tgm.nx = ctx->tg.nx;
  tgm.j0 = (int)(k0 / (size_t)ctx->tg.nx);
  tgm.j1 = tgm.j0 + (int)(kcnt / (size_t)ctx->tg.nx) - 1;
  tgm.tiles_x = (tgm.nx + 31) / 32;
  tgm.dtx = make_fastdiv((unsigned)tgm.tiles_x);
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
      CK(ctx->d_iscratch.alloc(SOSIE_ISCRATCH));
      CK(ctx->d_tilecnt.alloc(8));
      int* dcnt = ctx->d_tilecnt.p;
      CK(cudaMemsetAsync(dcnt, 0, 8 * sizeof(int), ctx->stream));
      CK(ctx->d_tilelist.alloc(4 * (size_t)tgm.ntiles));
      k_tilebox<<<grid_for(ctx, (size_t)tgm.ntiles * 256, 256), 256, 0, ctx->stream>>>(ctx->d_pidx.p, v, tgm, ctx->d_boxes.p, dcnt, ctx->d_tilelist.p);
      KLAUNCH_CHECK();
      int hc[8];
      if (int rc = fetch_small(ctx, dcnt, hc, sizeof(hc))) return rc;   // once per mapping
      ctx->tiles_tma = hc[0]; ctx->tiles_grp = hc[1]; ctx->tiles_other = hc[2];
      ctx->tiles_cls[0] = hc[4]; ctx->tiles_cls[1] = hc[5]; ctx->tiles_cls[2] = hc[6]; ctx->tiles_cls[3] = hc[3];
      ctx->boxes_ready = true;
    }
    spc = nslab < APPLY_SLABS_PER_CHUNK_STAGED ? nslab : APPLY_SLABS_PER_CHUNK_STAGED;
    grid = dim3(tgm.ntiles, (nslab + spc - 1) / spc, 1);
    // 16-byte staging (group kernel, TMA) needs a 16-byte aligned slab stack and row pitch
    const bool aligned = (v.nx % 4) == 0 && ((uintptr_t)dZ1 % 16) == 0;
    const bool use_grp = ctx->grp_mode != 0 && aligned;
    // TMA staging: a tensor map of the slab stack, at least one full box depth (off unless SOSIE_APPLY_TMA=1 / sosie_apply_tma:
    // cp.async.bulk.tensor traps as an illegal instruction under the gVisor nvproxy driver of this GPU pool -- profiles/r02_tma_nvproxy.md)
    CUtensorMap tmap;
    const bool use_tma = ctx->tma_mode != 0 && aligned && ctx->tiles_tma > 0 && nslab >= TMA_SD && apply_tensor_map(&tmap, dZ1, v.nx, v.ny, nslab);
    const int own = (use_grp ? 1 : 0) | (use_tma ? 2 : 0);
    const int n_tma = use_tma ? ctx->tiles_tma : 0;
    // (a tile that fits both goes to the TMA kernel: the counts overlap, so the other two kernels are launched whenever
    // they could own a tile)
    prof_begin(ctx, SOSIE_T_APPLY);
    if (use_tma) {
      // (function attributes are per device: a flag per context, not per process)
      if (!ctx->attr_tma) { CK(cudaFuncSetAttribute(k_apply_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_BYTES)); ctx->attr_tma = true; }
      k_apply_tma<<<grid, 256, TMA_SMEM_BYTES, ctx->stream>>>(tmap, ctx->d_pidx.p, ctx->d_pw.p, nt, v, dZ2, nslab, spc, po, tgm, ctx->d_boxes.p, own);
      KLAUNCH_CHECK();
    }
    if (use_grp && n_tma < tgm.ntiles) {
      const FastDiv dnx = make_fastdiv((unsigned)v.nx);
      const int* lst = use_tma ? nullptr : ctx->d_tilelist.p;
      const int4 cnt = make_int4(ctx->tiles_cls[0], ctx->tiles_cls[1], ctx->tiles_cls[2], ctx->tiles_cls[3]);
      const int nl = tgm.ntiles;
      // slabs per block: long runs amortise the per-tile set-up (records, offsets, copy plan); enough blocks for ~16 waves
      const int waves_blocks = 16 * 4 * ctx->num_sms;
      // ... but not much longer than ~128 slabs: blocks that run together then work on the same slabs, neighbouring tiles
      // find their shared source rows in L2 (measured on D: 292 slabs per block move 1.6x the bytes of 128 per block)
      int nchunk_g = std::max(1, std::min((nslab + 3) / 4, std::max((waves_blocks + nl - 1) / nl, (nslab + 127) / 128)));
      if (ctx->grp_spc > 0) nchunk_g = std::max(1, (nslab + ctx->grp_spc - 1) / ctx->grp_spc);    // tuning hook (SOSIE_GRP_SPC)
      int spc_g = (nslab + nchunk_g - 1) / nchunk_g;
      spc_g = (spc_g + 7) / 8 * 8;   // whole groups of 8 (and of 4, 2): only a block's last group can be incomplete
      const dim3 gg(nl, (nslab + spc_g - 1) / spc_g, 1);
      if (use_clamp || dmask_trg)
        k_apply_grp<true><<<gg, 256, 0, ctx->stream>>>(ctx->d_pidx.p, ctx->d_pw.p, nt, v, dnx, dZ1, dZ2, nslab, spc_g, po, tgm, ctx->d_boxes.p, own, lst, cnt);
      else
        k_apply_grp<false><<<gg, 256, 0, ctx->stream>>>(ctx->d_pidx.p, ctx->d_pw.p, nt, v, dnx, dZ1, dZ2, nslab, spc_g, po, tgm, ctx->d_boxes.p, own, lst, cnt);
      KLAUNCH_CHECK();
    }
    if (!use_grp) {   // unaligned slab stack (or the test hook): the 4-byte cp.async kernel takes everything the TMA kernel does not
      k_apply_staged<<<grid, 256, 0, ctx->stream>>>(ctx->d_pidx.p, ctx->d_pw.p, nt, v, dZ1, dZ2, nslab, spc, po, tgm, ctx->d_boxes.p, own, nullptr, tgm.ntiles);
      KLAUNCH_CHECK();
    }
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
                                                      ctx->d_ab.p, ctx->d_dnp.p, store ? ctx->d_pidx.p : nullptr, store ? ctx->d_pw.p : nullptr);
      KLAUNCH_CHECK();
      if (store) { ctx->pack_ready = true; ctx->pack_gen++; ctx->boxes_ready = false; }
      else ctx->unpacked_applies++;
    } else {
      k_apply1<false><<<grid.x, 256, 0, ctx->stream>>>(ctx->d_pidx.p, ctx->d_pw.p, v, dZ1, dZ2, po, tgm, ctx->sg, ctx->tg, ctx->k_ew, nullptr,
                                                       nullptr, nullptr, nullptr, nullptr);   // nslab == 1 here
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
