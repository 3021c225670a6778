// bilin_pipeline.cu -- the FIRST call of BILIN_2D from host arrays, as one software pipeline.
//
// The reference's first call (src/mod_bilin_2d.f90:173-206) builds the mapping (MAPPING_BL), writes it to the
// mapping file (P2D_MAPPING_AB, src/io_ezcdf.f90:2195), and applies it (:209-238).  From host arrays that is, on
// the headline shape (21600x10800 -> 8354x4729): 0.63 GB of target coordinates + a 0.93 GB slab in, 1.19 GB of
// mapping + a 0.16 GB field out, around 8 ms of kernels.  Done one after the other the copies cost ~60 ms; PCIe
// is full duplex and the copy engines run beside the SMs, so here the three overlap:
//
//   H2D stream : [src axes][Y,X rows of chunk 0][chunk 1] ... [rest of Y (and X row 1)]        [Z1 rows lo..hi]
//   compute    :           [map+pack chunk 0][map+pack chunk 1] ...        [prelude][row range]      [apply]
//   D2H stream :                      [mapping rows chunk 0][chunk 1] ...                              [Z2]
//
// * Chunks are blocks of target rows (of this context's shard).  Their search runs SPECULATIVELY with every row
//   searchable; the FIND_NEAREST_POINT prelude (src/mod_manip.f90:888-947) needs the whole latitude array and runs
//   when it has arrived.  If it restricts the searched rows (target wider than the source in latitude) the chunks
//   not entirely inside [j_strt, j_stop] are recomputed with the true range and copied out again, and the error
//   flags of the speculative pass are discarded for them: the result is always the sequential one.
// * The source slab is uploaded last, and only the rows the packed records read (bilin_src_rowrange): a regional
//   target touches a fraction of a global slab (37 % of the rows on the headline shape).
// * Eligibility: regular source given as 1-D axes (=> EASY search, no carried state), 2-D target.  Everything else
//   takes the sequential path (sosie_bilin_set_grids + bilin_map_impl + apply).
#include <algorithm>
#include <cstdlib>
#include <utility>

#include "geom.cuh"

int fill_i32_dev(sosie_ctx* ctx, int32_t* p, size_t n, int32_t v);

namespace {

struct Pipe {
  cudaStream_t h2d = nullptr, d2h = nullptr;
  std::vector<cudaEvent_t> evs;
  // SOSIE_PIPE_TRACE=1: timestamps (ms since the first mark) of the pipeline's milestones on their streams
  bool trace = getenv("SOSIE_PIPE_TRACE") != nullptr;
  std::vector<std::pair<std::string, cudaEvent_t>> marks;
  void mark(const char* name, cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    marks.emplace_back(name, e);
  }
  void report() {
    if (!trace || marks.empty()) return;
    for (auto& m : marks) {
      float ms = 0.f;
      cudaEventSynchronize(m.second);
      cudaEventElapsedTime(&ms, marks[0].second, m.second);
      fprintf(stderr, "[sosie pipe] %-28s %8.3f ms\n", m.first.c_str(), ms);
    }
  }
  ~Pipe() {
    for (auto& m : marks) cudaEventDestroy(m.second);
    for (auto e : evs) cudaEventDestroy(e);
    if (h2d) cudaStreamDestroy(h2d);
    if (d2h) cudaStreamDestroy(d2h);
  }
  cudaEvent_t ev() {
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    evs.push_back(e);
    return e;
  }
};

}  // namespace

// copy rows j0..j1 (0-based, inclusive) of the cached mapping to the host arrays (any may be NULL)
int bilin_map_rows_d2h(sosie_ctx* ctx, cudaStream_t st, int j0, int j1, int32_t* metrics, double* alphabeta,
                       int16_t* id_problem, float* dist_np, int32_t* mask_out) {
  if (j1 < j0) return SOSIE_OK;
  const size_t nx = (size_t)ctx->tg.nx, nt = nx * ctx->tg.ny;
  const size_t off = (size_t)j0 * nx, cnt = (size_t)(j1 - j0 + 1) * nx;
  if (metrics)
    for (int pl = 0; pl < 3; pl++)
      CK(cudaMemcpyAsync(metrics + pl * nt + off, ctx->d_metrics.p + pl * nt + off, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (alphabeta)
    for (int pl = 0; pl < 2; pl++)
      CK(cudaMemcpyAsync(alphabeta + pl * nt + off, ctx->d_ab.p + pl * nt + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (id_problem) CK(cudaMemcpyAsync(id_problem + off, ctx->d_ipb.p + off, cnt * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
  if (dist_np && ctx->d_dnp.p) CK(cudaMemcpyAsync(dist_np + off, ctx->d_dnp.p + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (mask_out && ctx->mask_is_ones) { if (int rc = bilin_materialize_mask(ctx)) return rc; }
  if (mask_out && ctx->t_mask.p) CK(cudaMemcpyAsync(mask_out + off, ctx->t_mask.p + off, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  return SOSIE_OK;
}

static int pipelined_run(sosie_ctx* ctx, Pipe& P, int k_ew, int nx1, int ny1, const double* X10, const double* Y10,
                         const float* Z1, int nx2, int ny2, const double* X20, const double* Y20, float* Z2,
                         const int32_t* mask, int l_reg_src, int i_orca_trg, int nslab, int32_t* metrics, double* alphabeta,
                         int16_t* id_problem, float* dist_np) {
  const size_t nt = (size_t)nx2 * ny2, n1 = (size_t)nx1 * ny1;
  cudaStream_t sc = ctx->stream;
  CK(cudaStreamSynchronize(sc));  // context buffers may still be read by earlier work
  CK(cudaStreamCreateWithFlags(&P.h2d, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&P.d2h, cudaStreamNonBlocking));

  // ---- state of set_grids (BILIN_2D prologue, :97-171)
  ctx->grids_set = false; ctx->map_ready = false; ctx->pack_ready = false; ctx->unpacked_applies = 0; ctx->boxes_ready = false;
  ctx->k_ew = k_ew; ctx->l_reg_src = l_reg_src; ctx->i_orca_trg = i_orca_trg;
  ctx->sg = SrcGeom();
  ctx->sg.nx = nx1; ctx->sg.ny = ny1; ctx->sg.separable = 1;
  ctx->tg = TrgGeom();
  ctx->tg.nx = nx2; ctx->tg.ny = ny2; ctx->tg.is_1d = 0;
  if (int rc = shard_check(ctx, "sosie_bilin_2d_first")) return rc;
  CK(ctx->s_in_X.alloc(nx1)); CK(ctx->s_in_Y.alloc(ny1)); CK(ctx->t_X.alloc(nt)); CK(ctx->t_Y.alloc(nt)); CK(ctx->t_mask.alloc(nt));
  ctx->tg.X = ctx->t_X.p; ctx->tg.Y = ctx->t_Y.p;
  if (int rc = bilin_map_alloc(ctx)) return rc;
  CK(ctx->d_pidx.alloc(nt)); CK(ctx->d_pw.alloc(nt));
  CK(ctx->d_stage1.alloc(n1 * nslab)); CK(ctx->d_out.alloc(nt * nslab));
  CK(ctx->d_iscratch.alloc(SOSIE_ISCRATCH));

  // ---- source axes -> working source (pole rows, trig / Mercator / heading tables), EASY axes
  cudaEvent_t e_src = P.ev();
  P.mark("start h2d", P.h2d);
  CK(cudaMemcpyAsync(ctx->s_in_X.p, X10, (size_t)nx1 * 8, cudaMemcpyHostToDevice, P.h2d));
  CK(cudaMemcpyAsync(ctx->s_in_Y.p, Y10, (size_t)ny1 * 8, cudaMemcpyHostToDevice, P.h2d));
  CK(cudaEventRecord(e_src, P.h2d));

  // ---- chunks of this shard's target rows
  size_t k0s, kcs;
  shard_range(ctx, &k0s, &kcs);
  const int r0 = (int)(k0s / nx2), r1 = r0 + (int)(kcs / nx2) - 1;  // 0-based, inclusive
  const int nrows = r1 - r0 + 1;
  const int max_chunks = SOSIE_ISCRATCH - 8;
  int rows_per = std::max(1, (int)(((size_t)2 << 20) / (size_t)nx2));  // ~2 M targets per chunk
  int nchunk = (nrows + rows_per - 1) / rows_per;
  if (nchunk > max_chunks) { nchunk = max_chunks; rows_per = (nrows + nchunk - 1) / nchunk; nchunk = (nrows + rows_per - 1) / rows_per; }
  int* derr = ctx->d_iscratch.p;
  CK(cudaMemsetAsync(derr, 0, SOSIE_ISCRATCH * sizeof(int), sc));

  std::vector<cudaEvent_t> e_in(nchunk), e_map(nchunk);
  for (int c = 0; c < nchunk; c++) {
    const int ja = r0 + c * rows_per, jb = std::min(r1, ja + rows_per - 1);
    const size_t off = (size_t)ja * nx2, cnt = (size_t)(jb - ja + 1) * nx2;
    CK(cudaMemcpyAsync(ctx->t_Y.p + off, Y20 + off, cnt * 8, cudaMemcpyHostToDevice, P.h2d));
    CK(cudaMemcpyAsync(ctx->t_X.p + off, X20 + off, cnt * 8, cudaMemcpyHostToDevice, P.h2d));
    if (mask) CK(cudaMemcpyAsync(ctx->t_mask.p + off, mask + off, cnt * 4, cudaMemcpyHostToDevice, P.h2d));
    e_in[c] = P.ev();
    CK(cudaEventRecord(e_in[c], P.h2d));
  }
  // What the prelude needs beyond this shard's rows.  With an exchange callback (sosie_set_exchange) the ranks reduce their
  // own rows and swap the results: only the row below the shard (row-to-row differences), the two rows of :921 and row 1
  // of the longitudes (L_IS_GRID_REGULAR compares every row with it) are uploaded.  Without it: the rest of the latitudes.
  const bool partial = (r0 > 0 || r1 < ny2 - 1);
  const bool xchg = partial && ctx->xchg_fn && ctx->xchg_nranks > 1 && nx2 > 2 &&
                    sizeof(double) * 8 + (size_t)24 * nx2 + 64 <= SOSIE_MAILBOX_BYTES;
  auto up_rows = [&](const double* src, double* dst, int ja, int jb) -> int {   // 0-based inclusive rows, skipped when inside the shard
    for (int j = ja; j <= jb; j++) {
      if (j < 0 || j >= ny2 || (j >= r0 && j <= r1)) continue;
      CK(cudaMemcpyAsync(dst + (size_t)j * nx2, src + (size_t)j * nx2, (size_t)nx2 * 8, cudaMemcpyHostToDevice, P.h2d));
    }
    return SOSIE_OK;
  };
  auto up_rest = [&]() -> int {
    if (r0 > 0) {
      CK(cudaMemcpyAsync(ctx->t_Y.p, Y20, (size_t)r0 * nx2 * 8, cudaMemcpyHostToDevice, P.h2d));
      CK(cudaMemcpyAsync(ctx->t_X.p, X20, (size_t)nx2 * 8, cudaMemcpyHostToDevice, P.h2d));
      if (mask) CK(cudaMemcpyAsync(ctx->t_mask.p, mask, (size_t)r0 * nx2 * 4, cudaMemcpyHostToDevice, P.h2d));
    }
    if (r1 < ny2 - 1) {
      const size_t off = (size_t)(r1 + 1) * nx2, cnt = (size_t)(ny2 - 1 - r1) * nx2;
      CK(cudaMemcpyAsync(ctx->t_Y.p + off, Y20 + off, cnt * 8, cudaMemcpyHostToDevice, P.h2d));
      if (mask) CK(cudaMemcpyAsync(ctx->t_mask.p + off, mask + off, cnt * 4, cudaMemcpyHostToDevice, P.h2d));
    }
    return SOSIE_OK;
  };
  if (xchg) {
    const int jm = ny2 / 2;                                    // 1-based row of :921; rows jm-1, jm -> 0-based jm-2, jm-1
    if (int rc = up_rows(Y20, ctx->t_Y.p, r0 - 1, r0 - 1)) return rc;
    if (int rc = up_rows(Y20, ctx->t_Y.p, jm - 2, jm - 1)) return rc;
    if (int rc = up_rows(X20, ctx->t_X.p, 0, 0)) return rc;
  } else if (int rc = up_rest()) return rc;
  cudaEvent_t e_coords = P.ev();
  CK(cudaEventRecord(e_coords, P.h2d));
  P.mark("coords uploaded (h2d)", P.h2d);

  // ---- compute stream: source preparation, then one (map, pack) pair per chunk as its rows arrive
  CK(cudaStreamWaitEvent(sc, e_src, 0));
  ctx->mask_is_ones = (mask == nullptr);
  if (int rc = bilin_prepare_src(ctx)) return rc;
  if (int rc = bilin_easy_prepare(ctx)) return rc;
  P.mark("source prepared (compute)", sc);
  ctx->map_info[5] = 0;
  for (int c = 0; c < nchunk; c++) {
    const int ja = r0 + c * rows_per, jb = std::min(r1, ja + rows_per - 1);
    const size_t off = (size_t)ja * nx2, cnt = (size_t)(jb - ja + 1) * nx2;
    CK(cudaStreamWaitEvent(sc, e_in[c], 0));
    if (c == nchunk - 1) prof_begin(ctx, SOSIE_T_MAP);
    if (int rc = bilin_easy_rows(ctx, 1, ny2, off, cnt, derr + 8 + c)) return rc;
    if (c == nchunk - 1) prof_end(ctx, SOSIE_T_MAP);
    if (int rc = bilin_pack_range(ctx, off, cnt)) return rc;
    e_map[c] = P.ev();
    CK(cudaEventRecord(e_map[c], sc));
    CK(cudaStreamWaitEvent(P.d2h, e_map[c], 0));
    if (int rc = bilin_map_rows_d2h(ctx, P.d2h, ja, jb, metrics, alphabeta, id_problem, dist_np, nullptr)) return rc;
    if (c == 0) { P.mark("chunk 0 mapped (compute)", sc); P.mark("chunk 0 downloaded (d2h)", P.d2h); }
  }
  P.mark("all chunks mapped (compute)", sc);
  P.mark("mapping downloaded (d2h)", P.d2h);

  // ---- prelude on the complete latitude array; validate the speculation
  CK(cudaStreamWaitEvent(sc, e_coords, 0));
  MapPlan pl;
  bool need_full_x = false, undecided = !xchg;
  ctx->xchg_used = 0;
  if (xchg) {
    if (int rc = bilin_map_prelude_exchange(ctx, &pl, r0 + 1, r1 + 1, &undecided)) return rc;
    if (undecided) {   // no rank could prove the target irregular: the ordinary prelude, on the whole arrays
      if (int rc = up_rest()) return rc;
      cudaEvent_t e_rest = P.ev();
      CK(cudaEventRecord(e_rest, P.h2d));
      CK(cudaStreamWaitEvent(sc, e_rest, 0));
    } else ctx->xchg_used = 1;
  }
  if (undecided) {
    if (int rc = bilin_map_prelude(ctx, &pl, partial ? r0 + 1 : 0, partial ? r1 + 1 : 0, &need_full_x)) return rc;
    if (need_full_x) {  // every resident row equals row 1: the grid may be regular, the test needs all rows
      CK(cudaMemcpyAsync(ctx->t_X.p, X20, nt * 8, cudaMemcpyHostToDevice, sc));
      if (int rc = bilin_map_prelude(ctx, &pl, 0, 0, nullptr)) return rc;
    }
  }
  if (!pl.reg_s) { ctx->err = "internal: pipelined first call with an irregular source"; return SOSIE_ERR_ARG; }
  std::vector<int> herr(SOSIE_ISCRATCH);
  bool redo_any = false;
  for (int c = 0; c < nchunk; c++) {
    const int ja = r0 + c * rows_per, jb = std::min(r1, ja + rows_per - 1);
    if (ja + 1 >= pl.jlo && jb + 1 <= pl.jhi) continue;  // speculation holds (rows are 1-based in the plan)
    const size_t off = (size_t)ja * nx2, cnt = (size_t)(jb - ja + 1) * nx2;
    CK(cudaMemsetAsync(derr + 8 + c, 0, sizeof(int), sc));
    if (int rc = bilin_easy_rows(ctx, pl.jlo, pl.jhi, off, cnt, derr + 8 + c)) return rc;
    if (int rc = bilin_pack_range(ctx, off, cnt)) return rc;
    cudaEvent_t e = P.ev();
    CK(cudaEventRecord(e, sc));
    CK(cudaStreamWaitEvent(P.d2h, e, 0));
    if (int rc = bilin_map_rows_d2h(ctx, P.d2h, ja, jb, metrics, alphabeta, id_problem, dist_np, nullptr)) return rc;
    redo_any = true;
  }
  (void)redo_any;
  P.mark("prelude done (compute)", sc);
  if (int rc = fetch_small(ctx, derr, herr.data(), SOSIE_ISCRATCH * sizeof(int))) return rc;
  int e_all = 0;
  for (int c = 0; c < nchunk; c++) e_all |= herr[8 + c];
  if (e_all & 1) { ctx->err = "ERROR (L_InPoly of mod_poly.f90): we only expect positive longitudes!"; return SOSIE_ERR_REF_STOP; }
  if (e_all & 6) { ctx->err = "FIND_NEAREST: The nearest point was not found! (ji_s or jj_s = 0)"; return SOSIE_ERR_REF_STOP; }
  ctx->grids_set = true; ctx->map_ready = true; ctx->map_nt = nt; ctx->trg_partial = partial;
  map_rows_from_shard(ctx);
  ctx->pack_ready = true; ctx->pack_gen++; ctx->boxes_ready = false;

  // ---- source slab rows the records read, apply, result
  int rlo = 0, rhi = -1;
  if (int rc = bilin_src_rowrange(ctx, k0s, kcs, &rlo, &rhi)) return rc;
  ctx->rows_gen = ctx->pack_gen; ctx->src_row_lo = rlo; ctx->src_row_hi = rhi;
  cudaEvent_t e_z = P.ev();
  if (rhi >= rlo) {
    const size_t off = (size_t)rlo * nx1, cnt = (size_t)(rhi - rlo + 1) * nx1;
    for (int s = 0; s < nslab; s++)
      CK(cudaMemcpyAsync(ctx->d_stage1.p + (size_t)s * n1 + off, Z1 + (size_t)s * n1 + off, cnt * sizeof(float), cudaMemcpyHostToDevice, P.h2d));
  }
  CK(cudaEventRecord(e_z, P.h2d));
  P.mark("row range known (compute)", sc);
  P.mark("slab rows uploaded (h2d)", P.h2d);
  CK(cudaStreamWaitEvent(sc, e_z, 0));
  if (int rc = bilin_apply_impl(ctx, nslab, ctx->d_stage1.p, ctx->d_out.p, 1, 0, 0.f, 0.f, nullptr, 0.f)) return rc;
  cudaEvent_t e_a = P.ev();
  CK(cudaEventRecord(e_a, sc));
  CK(cudaStreamWaitEvent(P.d2h, e_a, 0));
  {
    // ORCA last-row patch may rewrite row ny2 from other rows: only this shard's rows are meaningful either way
    const size_t off = (size_t)r0 * nx2, cnt = (size_t)nrows * nx2;
    for (int s = 0; s < nslab; s++)
      CK(cudaMemcpyAsync(Z2 + (size_t)s * nt + off, ctx->d_out.p + (size_t)s * nt + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, P.d2h));
  }
  P.mark("applied (compute)", sc);
  P.mark("field downloaded (d2h)", P.d2h);
  CK(cudaStreamSynchronize(P.d2h));
  CK(cudaStreamSynchronize(sc));
  P.report();
  return SOSIE_OK;
}

int bilin_first_pipelined(sosie_ctx* ctx, int k_ew, int nx1, int ny1, const double* X10, const double* Y10, const float* Z1,
                          int nx2, int ny2, const double* X20, const double* Y20, float* Z2, const int32_t* mask,
                          int l_reg_src, int i_orca_trg, int nslab, int32_t* metrics, double* alphabeta, int16_t* id_problem,
                          float* dist_np, bool* handled) {
  *handled = false;
  const size_t nt = (size_t)nx2 * ny2, n1 = (size_t)nx1 * ny1;
  if (!l_reg_src || nx1 < 3 || ny1 < 3 || nx2 < 1 || ny2 < 2 || nslab < 1) return SOSIE_OK;
  if (nt < ctx->pipeline_min_targets) return SOSIE_OK;
  if ((n1 + nt) * sizeof(float) * (size_t)nslab > ((size_t)24 << 30)) return SOSIE_OK;  // slabs are staged whole here
  *handled = true;
  Pipe P;
  int rc = pipelined_run(ctx, P, k_ew, nx1, ny1, X10, Y10, Z1, nx2, ny2, X20, Y20, Z2, mask, l_reg_src, i_orca_trg, nslab,
                         metrics, alphabeta, id_problem, dist_np);
  // whatever happened, no copy may still target the caller's arrays when we return
  if (P.h2d) cudaStreamSynchronize(P.h2d);
  if (P.d2h) cudaStreamSynchronize(P.d2h);
  cudaStreamSynchronize(ctx->stream);
  return rc;
}
