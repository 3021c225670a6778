// capi.cu -- extern "C" entry points declared in include/sosie_b200.h.
// Host pointers are staged through device buffers owned by the context; _dev entries work on the
// caller's device pointers.  No CPU compute path exists: without a CUDA device every call fails.
#include <cstdlib>
#include <algorithm>

#include "geom.cuh"

int akima_expand_axes(sosie_ctx* ctx, const double* dX1d, const double* dY1d, int nx, int ny, double* dX, double* dY);
int fill_i32_dev(sosie_ctx* ctx, int32_t* p, size_t n, int32_t v);

#define NEED(ctx_)                                   \
  if (!(ctx_)) return SOSIE_ERR_ARG;                 \
  sosie_ctx* ctx = (ctx_);                           \
  if (cudaSetDevice(ctx->device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return SOSIE_ERR_CUDA; } \
  if (ctx->q && queue_pending(ctx)) { if (int rcq__ = queue_flush(ctx)) return rcq__; }   /* a synchronous entry delivers the queued results first */

// peer-memory prelude exchange: mailbox life cycle (entries below, kernels in bilin_map.cu)
static void p2p_release(sosie_ctx* ctx) {
  for (int r = 0; r < 256; r++)
    if (ctx->p2p_opened[r]) { cudaIpcCloseMemHandle(ctx->p2p_opened[r]); ctx->p2p_opened[r] = nullptr; }
  if (ctx->p2p_peers_dev) { cudaFree(ctx->p2p_peers_dev); ctx->p2p_peers_dev = nullptr; }
  if (ctx->p2p_done) { cudaFree(ctx->p2p_done); ctx->p2p_done = nullptr; }
  if (ctx->p2p_box) { cudaFree(ctx->p2p_box); ctx->p2p_box = nullptr; }
  ctx->p2p_ready = false; ctx->p2p_nranks = 0; ctx->p2p_rank = -1; ctx->p2p_slot = 0; ctx->p2p_epoch = 0;
}
void p2p_destroy(sosie_ctx* ctx) { p2p_release(ctx); }


extern "C" {

const char* sosie_version(void) { return "sosie_b200 0.1 (sm_100a)"; }

int sosie_create(sosie_ctx** out, int device) {
  if (!out) return SOSIE_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return SOSIE_ERR_CUDA;
  if (device < 0 || device >= ndev) return SOSIE_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return SOSIE_ERR_CUDA;
  sosie_ctx* c = new sosie_ctx();
  c->device = device;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return SOSIE_ERR_CUDA; }
  c->own_stream = true;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) c->num_sms = sms;
  c->always_store_records = getenv("SOSIE_ALWAYS_STORE_RECORDS") != nullptr;
  if (const char* e = getenv("SOSIE_APPLY_TMA")) c->tma_mode = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SOSIE_GRP_SPC")) c->grp_spc = atoi(e);
  *out = c;
  return SOSIE_OK;
}

int sosie_destroy(sosie_ctx* c) {
  if (!c) return SOSIE_OK;
  cudaSetDevice(c->device);
  if (c->q) { queue_flush(c); queue_destroy(c); }
  cudaStreamSynchronize(c->stream);
  p2p_destroy(c);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return SOSIE_OK;
}

const char* sosie_last_error(const sosie_ctx* c) { return c ? c->err.c_str() : "null context"; }

int sosie_set_stream(sosie_ctx* c, void* s) {
  NEED(c);
  cudaStreamSynchronize(ctx->stream);
  if (s) {
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)s;
    ctx->own_stream = false;
  } else if (!ctx->own_stream) {
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  return SOSIE_OK;
}

int sosie_sync(sosie_ctx* c) {
  NEED(c);
  CK(cudaStreamSynchronize(ctx->stream));
  return bilin_map_check(ctx);   // status of a sosie_bilin_map_async that has not been delivered yet
}

int64_t sosie_launch_count(const sosie_ctx* c) { return c ? c->launches : 0; }

int sosie_profile(sosie_ctx* c, int32_t enable) {
  NEED(c);
  ctx->prof = enable != 0;
  return SOSIE_OK;
}
int sosie_get_timing(sosie_ctx* c, int32_t id, float* ms) {
  NEED(c);
  if (id < 0 || id >= SOSIE_T_COUNT || !ms) return SOSIE_ERR_ARG;
  if (!ctx->ev_used[id]) { *ms = -1.f; return SOSIE_OK; }
  CK(cudaEventSynchronize(ctx->ev1[id]));
  CK(cudaEventElapsedTime(ms, ctx->ev0[id], ctx->ev1[id]));
  return SOSIE_OK;
}
int sosie_bilin_set_shard(sosie_ctx* c, int32_t j0, int32_t j1) {
  NEED(c);
  if (j0 < 0 || j1 < j0) { ctx->err = "sosie_bilin_set_shard: bad row range"; return SOSIE_ERR_ARG; }
  if ((j0 == 0) != (j1 == 0)) { ctx->err = "sosie_bilin_set_shard: rows are 1-based (0,0 = all rows)"; return SOSIE_ERR_ARG; }
  if ((ctx->grids_set || ctx->map_ready) && ctx->tg.ny > 0 && j1 > ctx->tg.ny) {
    ctx->err = "sosie_bilin_set_shard: row range exceeds the target grid"; return SOSIE_ERR_ARG;
  }
  ctx->shard_j0 = j0; ctx->shard_j1 = j1;
  ctx->pack_ready = false; ctx->unpacked_applies = 0;
  // the EASY search fills the mapping arrays (and a pipelined first call the target coordinates) only for the rows of the
  // shard active when it ran: a mapping that does not cover the new rows is dropped (the next apply fails with
  // SOSIE_ERR_STATE instead of packing rows that were never computed)
  if (ctx->map_ready && ctx->map_j0 >= 1) {
    const int n0 = j0 >= 1 ? j0 : 1, n1 = j0 >= 1 ? j1 : ctx->tg.ny;
    if (n0 < ctx->map_j0 || n1 > ctx->map_j1) ctx->map_ready = false;
  }
  return SOSIE_OK;
}

int sosie_set_exchange(sosie_ctx* c, int32_t nranks, sosie_allgather_fn fn, void* user) {
  NEED(c);
  if (fn && nranks < 1) { ctx->err = "sosie_set_exchange: nranks must be >= 1"; return SOSIE_ERR_ARG; }
  ctx->xchg_fn = fn; ctx->xchg_user = user; ctx->xchg_nranks = fn ? nranks : 0;
  return SOSIE_OK;
}
int sosie_exchange_used(sosie_ctx* c) { return c ? c->xchg_used : 0; }

// ---------------------------------------------------------------------------------------------------
static int set_grids_common(sosie_ctx* ctx, int k_ew, int nx1, int ny1, int src_1d, int l_reg_src, int nx2, int ny2, int trg_1d,
                            const int32_t* dmask, bool mask_on_device, int i_orca_trg) {
  if (nx1 < 3 || ny1 < 3 || nx2 < 1 || ny2 < 1) { ctx->err = "sosie_bilin_set_grids: bad extents"; return SOSIE_ERR_ARG; }
  ctx->k_ew = k_ew; ctx->l_reg_src = l_reg_src; ctx->i_orca_trg = i_orca_trg;
  ctx->sg = SrcGeom();
  ctx->sg.nx = nx1; ctx->sg.ny = ny1; ctx->sg.separable = src_1d ? 1 : 0;
  const bool keep_map = ctx->map_ready && ctx->map_nt == (size_t)nx2 * ny2 && ctx->tg.nx == nx2 && ctx->tg.ny == ny2 &&
                        ctx->map_info[5] == -1;  // -1: mapping came from sosie_bilin_load_map
  ctx->tg = TrgGeom();
  ctx->tg.nx = nx2; ctx->tg.ny = ny2; ctx->tg.is_1d = trg_1d ? 1 : 0;
  ctx->tg.X = ctx->t_X.p; ctx->tg.Y = ctx->t_Y.p;
  if (int rc = shard_check(ctx, "sosie_bilin_set_grids")) return rc;
  const size_t nt = (size_t)nx2 * ny2;
  ctx->trg_partial = false;
  CK(ctx->t_mask.alloc(nt));
  if (dmask) {
    CK(cudaMemcpyAsync(ctx->t_mask.p, dmask, nt * sizeof(int32_t), mask_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    ctx->mask_is_ones = false;
  } else {
    ctx->mask_is_ones = true;   // filled lazily (TWISTED search, get_map): the EASY path never reads an all-ones mask
  }
  if (int rc = bilin_prepare_src(ctx)) return rc;
  ctx->grids_set = true;
  ctx->pack_ready = false; ctx->unpacked_applies = 0;
  if (!keep_map) ctx->map_ready = false;
  return SOSIE_OK;
}

int sosie_bilin_set_grids(sosie_ctx* c, int32_t k_ew, int32_t nx1, int32_t ny1, const double* X10, const double* Y10, int32_t src_1d,
                          int32_t l_reg_src, int32_t nx2, int32_t ny2, const double* X20, const double* Y20, int32_t trg_1d,
                          const int32_t* mask, int32_t i_orca_trg) {
  NEED(c);
  if (!X10 || !Y10 || !X20 || !Y20) { ctx->err = "sosie_bilin_set_grids: null coordinate pointer"; return SOSIE_ERR_ARG; }
  size_t nsx = src_1d ? (size_t)nx1 : (size_t)nx1 * ny1, nsy = src_1d ? (size_t)ny1 : (size_t)nx1 * ny1;
  size_t ntx = trg_1d ? (size_t)nx2 : (size_t)nx2 * ny2, nty = trg_1d ? (size_t)ny2 : (size_t)nx2 * ny2;
  CK(ctx->s_in_X.alloc(nsx)); CK(ctx->s_in_Y.alloc(nsy)); CK(ctx->t_X.alloc(ntx)); CK(ctx->t_Y.alloc(nty));
  CK(cudaMemcpyAsync(ctx->s_in_X.p, X10, nsx * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->s_in_Y.p, Y10, nsy * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->t_X.p, X20, ntx * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->t_Y.p, Y20, nty * 8, cudaMemcpyHostToDevice, ctx->stream));
  return set_grids_common(ctx, k_ew, nx1, ny1, src_1d, l_reg_src, nx2, ny2, trg_1d, mask, false, i_orca_trg);
}

int sosie_bilin_set_grids_dev(sosie_ctx* c, int32_t k_ew, int32_t nx1, int32_t ny1, const double* dX10, const double* dY10,
                              int32_t src_1d, int32_t l_reg_src, int32_t nx2, int32_t ny2, const double* dX20, const double* dY20,
                              int32_t trg_1d, const int32_t* dmask, int32_t i_orca_trg) {
  NEED(c);
  if (!dX10 || !dY10 || !dX20 || !dY20) { ctx->err = "sosie_bilin_set_grids_dev: null coordinate pointer"; return SOSIE_ERR_ARG; }
  size_t nsx = src_1d ? (size_t)nx1 : (size_t)nx1 * ny1, nsy = src_1d ? (size_t)ny1 : (size_t)nx1 * ny1;
  size_t ntx = trg_1d ? (size_t)nx2 : (size_t)nx2 * ny2, nty = trg_1d ? (size_t)ny2 : (size_t)nx2 * ny2;
  ctx->s_in_X.borrow(dX10, nsx); ctx->s_in_Y.borrow(dY10, nsy); ctx->t_X.borrow(dX20, ntx); ctx->t_Y.borrow(dY20, nty);
  ctx->axes_hint_dev = src_1d != 0;
  const int rc = set_grids_common(ctx, k_ew, nx1, ny1, src_1d, l_reg_src, nx2, ny2, trg_1d, dmask, true, i_orca_trg);
  ctx->axes_hint_dev = false;
  return rc;
}

int sosie_bilin_src_axes_hint(sosie_ctx* c, int32_t nx1, int32_t ny1, const double* X10_host, const double* Y10_host) {
  NEED(c);
  ctx->axes_hint.clear();
  if (!X10_host || !Y10_host) return SOSIE_OK;
  if (nx1 < 3 || ny1 < 3) { ctx->err = "sosie_bilin_src_axes_hint: bad extents"; return SOSIE_ERR_ARG; }
  ctx->axes_hint.resize((size_t)nx1 + ny1);
  memcpy(ctx->axes_hint.data(), X10_host, sizeof(double) * nx1);
  memcpy(ctx->axes_hint.data() + nx1, Y10_host, sizeof(double) * ny1);
  ctx->axes_hint_verify = getenv("SOSIE_VERIFY_HINTS") != nullptr;
  return SOSIE_OK;
}

int sosie_bilin_map(sosie_ctx* c) {
  NEED(c);
  return bilin_map_impl(ctx, false);
}

int sosie_bilin_map_async(sosie_ctx* c) {
  NEED(c);
  return bilin_map_impl(ctx, true);
}

int sosie_set_exchange_dev(sosie_ctx* c, int32_t nranks, sosie_allgather_dev_fn fn, void* user) {
  NEED(c);
  if (fn && (nranks < 1 || nranks > 256)) { ctx->err = "sosie_set_exchange_dev: nranks must be in 1..256"; return SOSIE_ERR_ARG; }
  ctx->xchg_dev_fn = fn; ctx->xchg_dev_user = user; ctx->xchg_dev_nranks = fn ? nranks : 0;
  return SOSIE_OK;
}

// ---- prelude exchange over peer memory (see k_xchg_push / k_xchg_merge in bilin_map.cu) ------------------------------------
int sosie_p2p_init(sosie_ctx* c, int32_t nranks, int32_t rank, size_t slot_bytes, void* ipc_handle_out, void** mailbox_out) {
  NEED(c);
  if (nranks < 2 || nranks > 256 || rank < 0 || rank >= nranks) { ctx->err = "sosie_p2p_init: nranks in 2..256, 0 <= rank < nranks"; return SOSIE_ERR_ARG; }
  if (slot_bytes == 0) slot_bytes = (size_t)2 << 20;   // 2 MB: 64 B + 24 B per target column, 87 000 columns
  slot_bytes = (slot_bytes + 255) / 256 * 256;
  CK(cudaStreamSynchronize(ctx->stream));
  p2p_release(ctx);
  const size_t bytes = 2 * (size_t)nranks * slot_bytes + 2 * (size_t)nranks * sizeof(unsigned) + 256;
  CK(cudaMalloc(&ctx->p2p_box, bytes));
  CK(cudaMemset(ctx->p2p_box, 0, bytes));
  CK(cudaMalloc(&ctx->p2p_done, sizeof(unsigned)));
  CK(cudaMemset(ctx->p2p_done, 0, sizeof(unsigned)));
  CK(cudaMalloc(&ctx->p2p_peers_dev, sizeof(unsigned char*) * (size_t)nranks));
  ctx->p2p_nranks = nranks; ctx->p2p_rank = rank; ctx->p2p_slot = slot_bytes;
  ctx->p2p_spin_cycles = 20000000000LL;
  if (const char* e = getenv("SOSIE_P2P_TIMEOUT_MS")) { const long long ms = atoll(e); if (ms >= 1 && ms <= 600000) ctx->p2p_spin_cycles = ms * 2000000LL; }
  if (mailbox_out) *mailbox_out = ctx->p2p_box;
  if (ipc_handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, ctx->p2p_box) != cudaSuccess) {
      cudaGetLastError();
      ctx->err = "sosie_p2p_init: cudaIpcGetMemHandle failed (CUDA IPC not available here); same-process peers can still connect by address";
      memset(ipc_handle_out, 0, 64);
      return SOSIE_ERR_CUDA;
    }
    memcpy(ipc_handle_out, &h, 64);
  }
  return SOSIE_OK;
}

int sosie_p2p_connect(sosie_ctx* c, const void* ipc_handles, void* const* mailboxes) {
  NEED(c);
  if (!ctx->p2p_box) { ctx->err = "sosie_p2p_connect: sosie_p2p_init first"; return SOSIE_ERR_STATE; }
  if (!ipc_handles && !mailboxes) { ctx->err = "sosie_p2p_connect: IPC handles or mailbox addresses needed"; return SOSIE_ERR_ARG; }
  const int n = ctx->p2p_nranks;
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->p2p_ready = false;
  for (int r = 0; r < 256; r++)      // a second connect replaces the first: its IPC mappings are closed, not leaked
    if (ctx->p2p_opened[r]) { cudaIpcCloseMemHandle(ctx->p2p_opened[r]); ctx->p2p_opened[r] = nullptr; }
  std::vector<unsigned char*> tab((size_t)n, nullptr);
  for (int r = 0; r < n; r++) {
    if (r == ctx->p2p_rank) { tab[r] = ctx->p2p_box; continue; }
    if (mailboxes) { tab[r] = (unsigned char*)mailboxes[r]; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)ipc_handles + 64 * (size_t)r, 64);
    void* q = nullptr;
    if (cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ctx->err = "sosie_p2p_connect: cudaIpcOpenMemHandle failed (no peer access between the ranks' GPUs, or CUDA IPC not available)";
      return SOSIE_ERR_CUDA;
    }
    ctx->p2p_opened[r] = q;
    tab[r] = (unsigned char*)q;
  }
  for (int r = 0; r < n; r++) if (!tab[r]) { ctx->err = "sosie_p2p_connect: null mailbox address"; return SOSIE_ERR_ARG; }
  CK(cudaMemcpy(ctx->p2p_peers_dev, tab.data(), sizeof(unsigned char*) * (size_t)n, cudaMemcpyHostToDevice));
  ctx->p2p_ready = true;
  return SOSIE_OK;
}

// FIND_NEAREST_POINT as a procedure of its own (src/mod_manip.f90:834-963; called directly by ij_from_lon_lat.f90:290)
int sosie_find_nearest_point(sosie_ctx* c, int32_t nx_t, int32_t ny_t, const double* Xtrg, const double* Ytrg, int32_t nx_s,
                             int32_t ny_s, const double* Xsrc, const double* Ysrc, int32_t* JIp, int32_t* JJp,
                             int32_t* mask_domain_trg) {
  NEED(c);
  if (!JIp || !JJp) { ctx->err = "sosie_find_nearest_point: null output pointer"; return SOSIE_ERR_ARG; }
  if (ctx->shard_j0 > 0) { ctx->err = "sosie_find_nearest_point: not available on a sharded context"; return SOSIE_ERR_STATE; }
  // the arrays as they are: no pole extension (l_reg_src = 0 keeps BILIN_2D's prologue out), no periodicity (unused by the search)
  if (int rc = sosie_bilin_set_grids(ctx, -1, nx_s, ny_s, Xsrc, Ysrc, 0, 0, nx_t, ny_t, Xtrg, Ytrg, 0, mask_domain_trg, 0)) return rc;
  ctx->search_only = true;
  int rc = bilin_map_impl(ctx);
  ctx->search_only = false;
  ctx->map_ready = false;
  if (rc) return rc;
  const size_t nt = (size_t)nx_t * ny_t;
  cudaStream_t st = ctx->stream;
  CK(cudaMemcpyAsync(JIp, ctx->d_metrics.p, nt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(JJp, ctx->d_metrics.p + nt, nt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (mask_domain_trg) {   // mask_domain_trg(:,:) = mask_ignore_t(:,:) (:958): TWISTED marks the targets it could not place
    if (ctx->mask_is_ones) { if (int r2 = bilin_materialize_mask(ctx)) return r2; }
    CK(cudaMemcpyAsync(mask_domain_trg, ctx->t_mask.p, nt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  ctx->grids_set = false;   // the next BILIN_2D-style call sets its own grids
  return SOSIE_OK;
}

int sosie_bilin_map_info(sosie_ctx* c, int32_t* info) {
  NEED(c);
  if (!info) return SOSIE_ERR_ARG;
  if (int rc = bilin_map_check(ctx)) return rc;
  for (int k = 0; k < 8; k++) info[k] = ctx->map_info[k];
  return SOSIE_OK;
}

int sosie_bilin_get_map(sosie_ctx* c, int32_t* metrics, double* alphabeta, int16_t* id_problem, float* dist_np, int32_t* mask_out) {
  NEED(c);
  if (int rc = bilin_map_check(ctx)) return rc;
  if (!ctx->map_ready) { ctx->err = "sosie_bilin_get_map: no mapping"; return SOSIE_ERR_STATE; }
  size_t k0, kcnt;
  shard_range(ctx, &k0, &kcnt);  // a sharded context only owns (and only copies) its rows
  const int j0 = (int)(k0 / (size_t)ctx->tg.nx), j1 = j0 + (int)(kcnt / (size_t)ctx->tg.nx) - 1;
  if (int rc = bilin_map_rows_d2h(ctx, ctx->stream, j0, j1, metrics, alphabeta, id_problem, dist_np, mask_out)) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return SOSIE_OK;
}

int sosie_bilin_load_map(sosie_ctx* c, int32_t nx2, int32_t ny2, const int32_t* metrics, const double* alphabeta,
                         const int16_t* id_problem) {
  NEED(c);
  if (!metrics || !alphabeta || nx2 < 1 || ny2 < 1) { ctx->err = "sosie_bilin_load_map: bad arguments"; return SOSIE_ERR_ARG; }
  if (ctx->grids_set && (ctx->tg.nx != nx2 || ctx->tg.ny != ny2)) ctx->grids_set = false;  // a new grid pair follows
  const size_t nt = (size_t)nx2 * ny2;
  cudaStream_t st = ctx->stream;
  CK(ctx->d_metrics.alloc(3 * nt)); CK(ctx->d_ab.alloc(2 * nt)); CK(ctx->d_ipb.alloc(nt)); CK(ctx->d_dnp.alloc(nt));
  ctx->map_nt = nt;
  CK(cudaMemcpyAsync(ctx->d_metrics.p, metrics, 3 * nt * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->d_ab.p, alphabeta, 2 * nt * sizeof(double), cudaMemcpyHostToDevice, st));
  if (id_problem) CK(cudaMemcpyAsync(ctx->d_ipb.p, id_problem, nt * sizeof(int16_t), cudaMemcpyHostToDevice, st));
  else CK(cudaMemsetAsync(ctx->d_ipb.p, 0, nt * sizeof(int16_t), st));
  CK(cudaMemsetAsync(ctx->d_dnp.p, 0, nt * sizeof(float), st));
  CK(cudaStreamSynchronize(st));
  if (!ctx->grids_set) { ctx->tg.nx = nx2; ctx->tg.ny = ny2; }
  ctx->map_ready = true;
  ctx->map_j0 = ctx->map_j1 = 0;   // a loaded mapping holds every row
  ctx->pack_ready = false; ctx->unpacked_applies = 0;
  for (int k = 0; k < 8; k++) ctx->map_info[k] = 0;
  ctx->map_info[5] = -1;
  return SOSIE_OK;
}

// device-resident variants of get_map / load_map: the multi-GPU flow broadcasts the mapping GPU -> GPU (NCCL over NVLink)
int sosie_bilin_get_map_dev(sosie_ctx* c, int32_t* dmetrics, double* dalphabeta, int16_t* did_problem, float* ddist_np) {
  NEED(c);
  if (!ctx->map_ready) { ctx->err = "sosie_bilin_get_map_dev: no mapping"; return SOSIE_ERR_STATE; }
  const size_t nt = (size_t)ctx->tg.nx * ctx->tg.ny;
  cudaStream_t st = ctx->stream;
  if (dmetrics) CK(cudaMemcpyAsync(dmetrics, ctx->d_metrics.p, 3 * nt * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  if (dalphabeta) CK(cudaMemcpyAsync(dalphabeta, ctx->d_ab.p, 2 * nt * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (did_problem) CK(cudaMemcpyAsync(did_problem, ctx->d_ipb.p, nt * sizeof(int16_t), cudaMemcpyDeviceToDevice, st));
  if (ddist_np && ctx->d_dnp.p) CK(cudaMemcpyAsync(ddist_np, ctx->d_dnp.p, nt * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return SOSIE_OK;
}

int sosie_bilin_load_map_dev(sosie_ctx* c, int32_t nx2, int32_t ny2, const int32_t* dmetrics, const double* dalphabeta,
                             const int16_t* did_problem) {
  NEED(c);
  if (!dmetrics || !dalphabeta || nx2 < 1 || ny2 < 1) { ctx->err = "sosie_bilin_load_map_dev: bad arguments"; return SOSIE_ERR_ARG; }
  if (ctx->grids_set && (ctx->tg.nx != nx2 || ctx->tg.ny != ny2)) ctx->grids_set = false;
  const size_t nt = (size_t)nx2 * ny2;
  cudaStream_t st = ctx->stream;
  CK(ctx->d_metrics.alloc(3 * nt)); CK(ctx->d_ab.alloc(2 * nt)); CK(ctx->d_ipb.alloc(nt)); CK(ctx->d_dnp.alloc(nt));
  ctx->map_nt = nt;
  CK(cudaMemcpyAsync(ctx->d_metrics.p, dmetrics, 3 * nt * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(ctx->d_ab.p, dalphabeta, 2 * nt * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (did_problem) CK(cudaMemcpyAsync(ctx->d_ipb.p, did_problem, nt * sizeof(int16_t), cudaMemcpyDeviceToDevice, st));
  else CK(cudaMemsetAsync(ctx->d_ipb.p, 0, nt * sizeof(int16_t), st));
  CK(cudaMemsetAsync(ctx->d_dnp.p, 0, nt * sizeof(float), st));
  if (!ctx->grids_set) { ctx->tg.nx = nx2; ctx->tg.ny = ny2; }
  ctx->map_ready = true;
  ctx->map_j0 = ctx->map_j1 = 0;   // a loaded mapping holds every row
  ctx->pack_ready = false; ctx->unpacked_applies = 0;
  for (int k = 0; k < 8; k++) ctx->map_info[k] = 0;
  ctx->map_info[5] = -1;
  return SOSIE_OK;
}

// host-pointer apply: slab chunks flow through two staging buffer pairs on three streams (H2D | kernels | D2H run
// concurrently: PCIe is full duplex); only the source rows the packed records read are uploaded, only this
// shard's target rows come back.
namespace {
struct ApplyPipe {
  cudaStream_t h2d = nullptr, d2h = nullptr;
  cudaEvent_t e_in[2] = {}, e_ap[2] = {}, e_out[2] = {};
  ~ApplyPipe() {
    for (int b = 0; b < 2; b++) { if (e_in[b]) cudaEventDestroy(e_in[b]); if (e_ap[b]) cudaEventDestroy(e_ap[b]); if (e_out[b]) cudaEventDestroy(e_out[b]); }
    if (h2d) cudaStreamDestroy(h2d);
    if (d2h) cudaStreamDestroy(d2h);
  }
};
}  // namespace

static int apply_host_run(sosie_ctx* ctx, ApplyPipe& P, int nslab, const float* Z1, float* Z2, int first_call) {
  const size_t n1 = (size_t)ctx->sg.nx * ctx->sg.ny, n2 = (size_t)ctx->tg.nx * ctx->tg.ny;
  cudaStream_t sc = ctx->stream;
  if (!ctx->pack_ready) { if (int rc = bilin_pack_impl(ctx)) return rc; }
  size_t k0, kcnt;
  shard_range(ctx, &k0, &kcnt);
  if (ctx->rows_gen != ctx->pack_gen) {
    if (int rc = bilin_src_rowrange(ctx, k0, kcnt, &ctx->src_row_lo, &ctx->src_row_hi)) return rc;
    ctx->rows_gen = ctx->pack_gen;
  }
  const int rlo = ctx->src_row_lo, rhi = ctx->src_row_hi;
  const size_t soff = rhi >= rlo ? (size_t)rlo * ctx->sg.nx : 0, scnt = rhi >= rlo ? (size_t)(rhi - rlo + 1) * ctx->sg.nx : 0;
  const bool whole_src = (scnt == n1), whole_trg = (kcnt == n2);
  const size_t budget = (size_t)1 << 30;  // bytes of staging per buffer pair
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)nslab, budget / ((n1 + n2) * sizeof(float))));
  const int nbuf = (nslab > chunk) ? 2 : 1;
  CK(ctx->d_stage1.alloc(n1 * chunk * nbuf));
  CK(ctx->d_out.alloc(n2 * chunk * nbuf));
  if (nbuf == 1) {
    // one chunk (the per-record / per-level call of the unchanged driver loop): nothing to overlap, so no extra streams or
    // events are created -- upload, kernel and download are queued on the context stream (creating and destroying two
    // streams and six events per call cost more than the 13 MB copy of a D-sized record)
    float* din = ctx->d_stage1.p;
    float* dout = ctx->d_out.p;
    if (whole_src) CK(cudaMemcpyAsync(din, Z1, n1 * nslab * sizeof(float), cudaMemcpyHostToDevice, sc));
    else
      for (int s = 0; s < nslab && scnt; s++)
        CK(cudaMemcpyAsync(din + (size_t)s * n1 + soff, Z1 + (size_t)s * n1 + soff, scnt * sizeof(float), cudaMemcpyHostToDevice, sc));
    if (int rc = bilin_apply_impl(ctx, nslab, din, dout, first_call, 0, 0.f, 0.f, nullptr, 0.f)) return rc;
    if (whole_trg) CK(cudaMemcpyAsync(Z2, dout, n2 * nslab * sizeof(float), cudaMemcpyDeviceToHost, sc));
    else
      for (int s = 0; s < nslab; s++)
        CK(cudaMemcpyAsync(Z2 + (size_t)s * n2 + k0, dout + (size_t)s * n2 + k0, kcnt * sizeof(float), cudaMemcpyDeviceToHost, sc));
    return SOSIE_OK;
  }
  CK(cudaStreamSynchronize(sc));
  CK(cudaStreamCreateWithFlags(&P.h2d, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&P.d2h, cudaStreamNonBlocking));
  for (int b = 0; b < 2; b++) {
    CK(cudaEventCreateWithFlags(&P.e_in[b], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&P.e_ap[b], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&P.e_out[b], cudaEventDisableTiming));
  }
  int it = 0;
  for (int s0 = 0; s0 < nslab; s0 += chunk, it++) {
    const int ns = std::min(chunk, nslab - s0), b = it % nbuf;
    float* din = ctx->d_stage1.p + (size_t)b * chunk * n1;
    float* dout = ctx->d_out.p + (size_t)b * chunk * n2;
    if (it >= nbuf) CK(cudaStreamWaitEvent(P.h2d, P.e_ap[b], 0));  // the kernels of chunk it-2 are done with this input buffer
    if (whole_src) CK(cudaMemcpyAsync(din, Z1 + (size_t)s0 * n1, n1 * ns * sizeof(float), cudaMemcpyHostToDevice, P.h2d));
    else
      for (int s = 0; s < ns && scnt; s++)
        CK(cudaMemcpyAsync(din + (size_t)s * n1 + soff, Z1 + (size_t)(s0 + s) * n1 + soff, scnt * sizeof(float), cudaMemcpyHostToDevice, P.h2d));
    CK(cudaEventRecord(P.e_in[b], P.h2d));
    CK(cudaStreamWaitEvent(sc, P.e_in[b], 0));
    if (it >= nbuf) CK(cudaStreamWaitEvent(sc, P.e_out[b], 0));    // chunk it-2's results have left this output buffer
    if (int rc = bilin_apply_impl(ctx, ns, din, dout, first_call && s0 == 0, 0, 0.f, 0.f, nullptr, 0.f)) return rc;
    CK(cudaEventRecord(P.e_ap[b], sc));
    CK(cudaStreamWaitEvent(P.d2h, P.e_ap[b], 0));
    if (whole_trg) CK(cudaMemcpyAsync(Z2 + (size_t)s0 * n2, dout, n2 * ns * sizeof(float), cudaMemcpyDeviceToHost, P.d2h));
    else
      for (int s = 0; s < ns; s++)
        CK(cudaMemcpyAsync(Z2 + (size_t)(s0 + s) * n2 + k0, dout + (size_t)s * n2 + k0, kcnt * sizeof(float), cudaMemcpyDeviceToHost, P.d2h));
    CK(cudaEventRecord(P.e_out[b], P.d2h));
  }
  return SOSIE_OK;
}

static int apply_host(sosie_ctx* ctx, int nslab, const float* Z1, float* Z2, int first_call) {
  if (!ctx->grids_set) { ctx->err = "sosie_bilin_apply: grids not set"; return SOSIE_ERR_STATE; }
  if (int rc = bilin_map_check(ctx)) return rc;
  if (nslab <= 0) return SOSIE_OK;
  ApplyPipe P;
  int rc = apply_host_run(ctx, P, nslab, Z1, Z2, first_call);
  if (P.h2d) cudaStreamSynchronize(P.h2d);
  if (P.d2h) cudaStreamSynchronize(P.d2h);
  cudaStreamSynchronize(ctx->stream);
  return rc;
}

int sosie_bilin_apply(sosie_ctx* c, int32_t nslab, const float* Z1, float* Z2, int32_t first_call) {
  NEED(c);
  if (!Z1 || !Z2 || nslab < 0) { ctx->err = "sosie_bilin_apply: bad arguments"; return SOSIE_ERR_ARG; }
  return apply_host(ctx, nslab, Z1, Z2, first_call);
}

int sosie_bilin_apply_dev(sosie_ctx* c, int32_t nslab, const float* dZ1, float* dZ2, int32_t first_call) {
  NEED(c);
  if (!ctx->grids_set) { ctx->err = "sosie_bilin_apply_dev: grids not set"; return SOSIE_ERR_STATE; }
  return bilin_apply_impl(ctx, nslab, dZ1, dZ2, first_call, 0, 0.f, 0.f, nullptr, 0.f);
}

int sosie_bilin_apply_ex_dev(sosie_ctx* c, int32_t nslab, const float* dZ1, float* dZ2, int32_t first_call, int32_t use_clamp,
                             float vmin, float vmax, const int32_t* dmask_trg, float rmiss_val) {
  NEED(c);
  if (!ctx->grids_set) { ctx->err = "sosie_bilin_apply_ex_dev: grids not set"; return SOSIE_ERR_STATE; }
  return bilin_apply_impl(ctx, nslab, dZ1, dZ2, first_call, use_clamp, vmin, vmax, dmask_trg, rmiss_val);
}

int sosie_bilin_apply_uv_dev(sosie_ctx* c, int32_t nslab, const float* dU1, const float* dV1, float* dUo, float* dVo,
                             const double* dxcos, const double* dxsin) {
  NEED(c);
  if (!ctx->grids_set) { ctx->err = "sosie_bilin_apply_uv_dev: grids not set"; return SOSIE_ERR_STATE; }
  return bilin_apply_uv_impl(ctx, nslab, dU1, dV1, dUo, dVo, dxcos, dxsin);
}

static int bilin_2d_common(sosie_ctx* ctx, int32_t k_ew, int32_t nx1, int32_t ny1, const double* X10, const double* Y10, int32_t src_1d,
                           const float* Z1, int32_t nx2, int32_t ny2, const double* X20, const double* Y20, int32_t trg_1d, float* Z2,
                           const int32_t* mask, int32_t l_reg_src, int32_t i_orca_trg, int32_t* l_first_call, int32_t have_mapping,
                           int32_t nslab, int32_t* metrics, double* alphabeta, int16_t* id_problem, float* dist_np) {
  if (!l_first_call) return SOSIE_ERR_ARG;
  if (!X10 || !Y10 || !X20 || !Y20 || !Z1 || !Z2 || nslab < 0) { ctx->err = "sosie_bilin_2d: null pointer / bad nslab"; return SOSIE_ERR_ARG; }
  int first = *l_first_call;
  if (first) {
    bool handled = false;
    if (!have_mapping && src_1d && !trg_1d) {
      if (int rc = bilin_first_pipelined(ctx, k_ew, nx1, ny1, X10, Y10, Z1, nx2, ny2, X20, Y20, Z2, mask, l_reg_src, i_orca_trg, nslab,
                                         metrics, alphabeta, id_problem, dist_np, &handled)) return rc;
    }
    if (handled) { *l_first_call = 0; return SOSIE_OK; }
    if (int rc = sosie_bilin_set_grids(ctx, k_ew, nx1, ny1, X10, Y10, src_1d, l_reg_src, nx2, ny2, X20, Y20, trg_1d, mask, i_orca_trg)) return rc;
    if (!have_mapping) { if (int rc = bilin_map_impl(ctx)) return rc; }
    else if (!ctx->map_ready) { ctx->err = "sosie_bilin_2d: have_mapping set but no mapping loaded"; return SOSIE_ERR_STATE; }
    if (metrics || alphabeta || id_problem || dist_np) {
      if (int rc = sosie_bilin_get_map(ctx, metrics, alphabeta, id_problem, dist_np, nullptr)) return rc;
    }
  } else {
    if (!ctx->grids_set || ctx->sg.nx != nx1 || ctx->sg.ny != ny1 || ctx->tg.nx != nx2 || ctx->tg.ny != ny2) {
      ctx->err = "sosie_bilin_2d: extents differ from the first call"; return SOSIE_ERR_ARG;
    }
  }
  if (int rc = apply_host(ctx, nslab, Z1, Z2, first)) return rc;
  *l_first_call = 0;  // l_first_call_interp_routine = .FALSE. (:267)
  return SOSIE_OK;
}

int sosie_bilin_2d(sosie_ctx* c, int32_t k_ew, int32_t nx1, int32_t ny1, const double* X10, const double* Y10, int32_t src_1d,
                   const float* Z1, int32_t nx2, int32_t ny2, const double* X20, const double* Y20, int32_t trg_1d, float* Z2,
                   const int32_t* mask, int32_t l_reg_src, int32_t i_orca_trg, int32_t* l_first_call, int32_t have_mapping,
                   int32_t nslab) {
  NEED(c);
  return bilin_2d_common(ctx, k_ew, nx1, ny1, X10, Y10, src_1d, Z1, nx2, ny2, X20, Y20, trg_1d, Z2, mask, l_reg_src, i_orca_trg,
                         l_first_call, have_mapping, nslab, nullptr, nullptr, nullptr, nullptr);
}

int sosie_bilin_2d_first(sosie_ctx* c, int32_t k_ew, int32_t nx1, int32_t ny1, const double* X10, const double* Y10, int32_t src_1d,
                         const float* Z1, int32_t nx2, int32_t ny2, const double* X20, const double* Y20, int32_t trg_1d, float* Z2,
                         const int32_t* mask, int32_t l_reg_src, int32_t i_orca_trg, int32_t nslab, int32_t* metrics,
                         double* alphabeta, int16_t* id_problem, float* dist_np) {
  NEED(c);
  int32_t first = 1;
  return bilin_2d_common(ctx, k_ew, nx1, ny1, X10, Y10, src_1d, Z1, nx2, ny2, X20, Y20, trg_1d, Z2, mask, l_reg_src, i_orca_trg,
                         &first, 0, nslab, metrics, alphabeta, id_problem, dist_np);
}

int sosie_apply_tma(sosie_ctx* c, int32_t on) {
  NEED(c);
  ctx->tma_mode = (on & 1) ? 1 : 0;
  ctx->grp_mode = (on & 4) ? 0 : 1;     // bit 2: also switch the group-staged kernel off (4-byte cp.async kernel only)
  return SOSIE_OK;
}
int sosie_apply_tile_classes(sosie_ctx* c, int32_t* counts) {
  NEED(c);
  if (!counts) return SOSIE_ERR_ARG;
  counts[0] = ctx->boxes_ready ? ctx->tiles_tma : -1; counts[1] = ctx->boxes_ready ? ctx->tiles_grp : -1;
  counts[2] = ctx->boxes_ready ? ctx->tiles_other : -1;
  return SOSIE_OK;
}
int sosie_selftest_div(sosie_ctx* c, uint64_t n, uint64_t seed, uint64_t* mismatches) {
  NEED(c);
  if (!mismatches) return SOSIE_ERR_ARG;
  unsigned long long m = 0;
  if (int rc = selftest_div_impl(ctx, n, seed, &m)) return rc;
  *mismatches = m;
  return SOSIE_OK;
}

int sosie_bilin_src_rows(sosie_ctx* c, int32_t* rows) {
  NEED(c);
  if (!rows) return SOSIE_ERR_ARG;
  if (!ctx->map_ready) { ctx->err = "sosie_bilin_src_rows: no mapping"; return SOSIE_ERR_STATE; }
  if (!ctx->pack_ready) { if (int rc = bilin_pack_impl(ctx)) return rc; }
  if (ctx->rows_gen != ctx->pack_gen) {
    size_t k0, kcnt;
    shard_range(ctx, &k0, &kcnt);
    if (int rc = bilin_src_rowrange(ctx, k0, kcnt, &ctx->src_row_lo, &ctx->src_row_hi)) return rc;
    ctx->rows_gen = ctx->pack_gen;
  }
  rows[0] = ctx->src_row_lo; rows[1] = ctx->src_row_hi;
  return SOSIE_OK;
}

int sosie_set_pipeline_min_targets(sosie_ctx* c, int64_t n) {
  NEED(c);
  ctx->pipeline_min_targets = n < 0 ? 0 : (size_t)n;
  return SOSIE_OK;
}

int sosie_host_register(void* p, size_t bytes) {
  if (!p || !bytes) return SOSIE_ERR_ARG;
  if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return SOSIE_ERR_CUDA; }
  pin_registry_add(p, bytes);
  return SOSIE_OK;
}
int sosie_host_unregister(void* p) {
  if (!p) return SOSIE_ERR_ARG;
  pin_registry_remove(p);
  return cudaHostUnregister(p) == cudaSuccess ? SOSIE_OK : SOSIE_ERR_CUDA;
}

int sosie_interp_bl(sosie_ctx* c, int32_t k_ew, int32_t nxi, int32_t nyi, const float* Z_in, int32_t n, const int32_t* jiP,
                    const int32_t* jjP, const int32_t* iqd, const double* xa, const double* xb, float* out) {
  NEED(c);
  if (n <= 0) return SOSIE_OK;
  cudaStream_t st = ctx->stream;
  DBuf<float> dZ, dout; DBuf<int32_t> di, dj, dq; DBuf<double> da, db;
  size_t nz = (size_t)nxi * nyi;
  CK(dZ.alloc(nz)); CK(dout.alloc(n)); CK(di.alloc(n)); CK(dj.alloc(n)); CK(dq.alloc(n)); CK(da.alloc(n)); CK(db.alloc(n));
  CK(cudaMemcpyAsync(dZ.p, Z_in, nz * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(di.p, jiP, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dj.p, jjP, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dq.p, iqd, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(da.p, xa, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(db.p, xb, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  if (int rc = interp_bl_impl(ctx, k_ew, nxi, nyi, dZ.p, n, di.p, dj.p, dq.p, da.p, db.p, dout.p)) return rc;
  CK(cudaMemcpyAsync(out, dout.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SOSIE_OK;
}

int sosie_track_interp_dev(sosie_ctx* c, int32_t ni, int32_t nj, int32_t nk, const float* dF1, const float* dF2, double vt1, double vt2,
                           const int8_t* dimask, int64_t n, const double* drt, const int32_t* djiP, const int32_t* djjP, const int32_t* diqd,
                           const double* dxa, const double* dxb, const double* dr_obs, double obs_lo, double obs_hi, int32_t clip_missing,
                           double rmissval, double* df_bl, double* df_np, int8_t* dfmask) {
  NEED(c);
  return track_interp_impl(ctx, ni, nj, nk, dF1, dF2, vt1, vt2, dimask, n, drt, djiP, djjP, diqd, dxa, dxb, dr_obs, obs_lo, obs_hi, clip_missing,
                           rmissval, df_bl, df_np, dfmask);
}

int sosie_track_interp(sosie_ctx* c, int32_t ni, int32_t nj, int32_t nk, const float* F1, const float* F2, double vt1, double vt2,
                       const int8_t* imask, int64_t n, const double* rt, const int32_t* jiP, const int32_t* jjP, const int32_t* iqd,
                       const double* xa, const double* xb, const double* r_obs, double obs_lo, double obs_hi, int32_t clip_missing,
                       double rmissval, double* f_bl, double* f_np, int8_t* fmask) {
  NEED(c);
  if (n <= 0) return SOSIE_OK;
  if (ni < 1 || nj < 1 || nk < 1 || !F1 || !imask || !jiP || !jjP || !iqd || !xa || !xb || !r_obs || !f_bl || !f_np || !fmask || (F2 && !rt)) {
    ctx->err = "sosie_track_interp: bad arguments"; return SOSIE_ERR_ARG;
  }
  cudaStream_t st = ctx->stream;
  const size_t nf = (size_t)ni * nj * nk, np_ = (size_t)n, no = np_ * nk;
  DBuf<float> dF1, dF2; DBuf<int8_t> dM, dfm; DBuf<int32_t> di, dj, dq; DBuf<double> drt, da, db, dro, dbl, dnp;
  CK(dF1.alloc(nf)); CK(dM.alloc(nf)); CK(di.alloc(np_)); CK(dj.alloc(np_)); CK(dq.alloc(np_)); CK(da.alloc(np_)); CK(db.alloc(np_));
  CK(dro.alloc(np_)); CK(dbl.alloc(no)); CK(dnp.alloc(no)); CK(dfm.alloc(no));
  CK(cudaMemcpyAsync(dF1.p, F1, nf * 4, cudaMemcpyHostToDevice, st));
  if (F2) {
    CK(dF2.alloc(nf)); CK(drt.alloc(np_));
    CK(cudaMemcpyAsync(dF2.p, F2, nf * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(drt.p, rt, np_ * 8, cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemcpyAsync(dM.p, imask, nf, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(di.p, jiP, np_ * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dj.p, jjP, np_ * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dq.p, iqd, np_ * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(da.p, xa, np_ * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(db.p, xb, np_ * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dro.p, r_obs, np_ * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dbl.p, f_bl, no * 8, cudaMemcpyHostToDevice, st));   // pre-filled by the caller, kept for invalid points
  CK(cudaMemcpyAsync(dnp.p, f_np, no * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dfm.p, fmask, no, cudaMemcpyHostToDevice, st));
  if (int rc = track_interp_impl(ctx, ni, nj, nk, dF1.p, F2 ? dF2.p : nullptr, vt1, vt2, dM.p, n, F2 ? drt.p : nullptr, di.p, dj.p, dq.p, da.p,
                                 db.p, dro.p, obs_lo, obs_hi, clip_missing, rmissval, dbl.p, dnp.p, dfm.p)) return rc;
  CK(cudaMemcpyAsync(f_bl, dbl.p, no * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(f_np, dnp.p, no * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(fmask, dfm.p, no, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SOSIE_OK;
}

// ---------------------------------------------------------------------------------------------------
int sosie_akima_set_grids(sosie_ctx* c, int32_t k_ew, int32_t nx1, int32_t ny1, const double* X10, const double* Y10, int32_t src_1d,
                          int32_t i_orca_src, int32_t nx2, int32_t ny2, const double* X20, const double* Y20, int32_t trg_1d) {
  NEED(c);
  if (nx1 < 3 || ny1 < 3 || nx2 < 1 || ny2 < 1 || !X10 || !Y10 || !X20 || !Y20) { ctx->err = "sosie_akima_set_grids: bad arguments"; return SOSIE_ERR_ARG; }
  cudaStream_t st = ctx->stream;
  ctx->ak_set = false;
  ctx->ak_kew = k_ew; ctx->ak_nx1 = nx1; ctx->ak_ny1 = ny1; ctx->ak_nx2 = nx2; ctx->ak_ny2 = ny2; ctx->ak_iorca = i_orca_src;
  const size_t n1 = (size_t)nx1 * ny1, nt = (size_t)nx2 * ny2;
  // stage the 2-D source coordinates in ak_ZX8/ak_ZY8 (akima_prepare_impl consumes them)
  ctx->ak_ZX8.release(); ctx->ak_ZY8.release();
  CK(ctx->ak_ZX8.alloc(n1)); CK(ctx->ak_ZY8.alloc(n1));
  if (src_1d) {
    DBuf<double> ax, ay; CK(ax.alloc(nx1)); CK(ay.alloc(ny1));
    CK(cudaMemcpyAsync(ax.p, X10, (size_t)nx1 * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ay.p, Y10, (size_t)ny1 * 8, cudaMemcpyHostToDevice, st));
    if (int rc = akima_expand_axes(ctx, ax.p, ay.p, nx1, ny1, ctx->ak_ZX8.p, ctx->ak_ZY8.p)) return rc;
    CK(cudaStreamSynchronize(st));
  } else {
    CK(cudaMemcpyAsync(ctx->ak_ZX8.p, X10, n1 * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->ak_ZY8.p, Y10, n1 * 8, cudaMemcpyHostToDevice, st));
  }
  CK(ctx->ak_X2.alloc(nt)); CK(ctx->ak_Y2.alloc(nt));
  if (trg_1d) {
    DBuf<double> ax, ay; CK(ax.alloc(nx2)); CK(ay.alloc(ny2));
    CK(cudaMemcpyAsync(ax.p, X20, (size_t)nx2 * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ay.p, Y20, (size_t)ny2 * 8, cudaMemcpyHostToDevice, st));
    if (int rc = akima_expand_axes(ctx, ax.p, ay.p, nx2, ny2, ctx->ak_X2.p, ctx->ak_Y2.p)) return rc;
    CK(cudaStreamSynchronize(st));
  } else {
    CK(cudaMemcpyAsync(ctx->ak_X2.p, X20, nt * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->ak_Y2.p, Y20, nt * 8, cudaMemcpyHostToDevice, st));
  }
  return akima_prepare_impl(ctx);
}

int sosie_akima_apply_dev(sosie_ctx* c, int32_t nslab, const float* dZ1, float* dZ2) {
  NEED(c);
  return akima_apply_impl(ctx, nslab, dZ1, dZ2);
}

int sosie_akima_apply(sosie_ctx* c, int32_t nslab, const float* Z1, float* Z2) {
  NEED(c);
  if (!ctx->ak_set) { ctx->err = "sosie_akima_apply: grids not set"; return SOSIE_ERR_STATE; }
  if (!Z1 || !Z2 || nslab < 0) return SOSIE_ERR_ARG;
  const size_t n1 = (size_t)ctx->ak_nx1 * ctx->ak_ny1, n2 = (size_t)ctx->ak_nx2 * ctx->ak_ny2;
  const size_t budget = (size_t)2 << 30;
  int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)nslab, budget / ((n1 + n2) * sizeof(float))));
  CK(ctx->ak_stage1.alloc(n1 * chunk)); CK(ctx->ak_stage2.alloc(n2 * chunk));
  cudaStream_t st = ctx->stream;
  for (int s0 = 0; s0 < nslab; s0 += chunk) {
    int ns = std::min(chunk, nslab - s0);
    CK(cudaMemcpyAsync(ctx->ak_stage1.p, Z1 + (size_t)s0 * n1, n1 * ns * sizeof(float), cudaMemcpyHostToDevice, st));
    if (int rc = akima_apply_impl(ctx, ns, ctx->ak_stage1.p, ctx->ak_stage2.p)) return rc;
    CK(cudaMemcpyAsync(Z2 + (size_t)s0 * n2, ctx->ak_stage2.p, n2 * ns * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  return SOSIE_OK;
}

int sosie_akima_untiled_slopes(sosie_ctx* c, int32_t on) {
  NEED(c);
  ctx->ak_untiled_slopes = on == 1;     // 0: tiled slopes kernel + evaluation (default), 1: per-point slopes kernel
  ctx->ak_fused = on == 3;              // 3: regular sources through the fused slopes + evaluation kernel (no slope arrays)
  return SOSIE_OK;
}

int sosie_akima_get_ixy(sosie_ctx* c, int32_t* ixy) {
  NEED(c);
  if (!ctx->ak_set || !ixy) return SOSIE_ERR_STATE;
  const size_t nt = (size_t)ctx->ak_nx2 * ctx->ak_ny2;
  CK(cudaMemcpyAsync(ixy, ctx->ak_ixy.p, 2 * nt * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return SOSIE_OK;
}

// ---------------------------------------------------------------------------------------------------
int sosie_bdrown_dev(sosie_ctx* c, int32_t k_ew, int32_t ni, int32_t nj, int32_t nslab, float* dX, const int32_t* dmask,
                     int32_t mask_per_slab, int32_t nb_inc, int32_t nb_smooth, float pmin, float pmax, float pval_land,
                     int32_t* nsweeps_host) {
  NEED(c);
  if (ni < 3 || nj < 3 || nslab < 1 || !dX || !dmask) { ctx->err = "sosie_bdrown: bad arguments"; return SOSIE_ERR_ARG; }
  return bdrown_impl(ctx, k_ew, ni, nj, nslab, dX, dmask, mask_per_slab, nb_inc, nb_smooth, pmin, pmax, pval_land, nsweeps_host);
}

int sosie_bdrown(sosie_ctx* c, int32_t k_ew, int32_t ni, int32_t nj, int32_t nslab, float* X, const int32_t* mask,
                 int32_t mask_per_slab, int32_t nb_inc, int32_t nb_smooth, float pmin, float pmax, float pval_land, int32_t* nsweeps) {
  NEED(c);
  if (ni < 3 || nj < 3 || nslab < 1 || !X || !mask) { ctx->err = "sosie_bdrown: bad arguments"; return SOSIE_ERR_ARG; }
  const size_t n = (size_t)ni * nj;
  cudaStream_t st = ctx->stream;
  DBuf<float>& dX = ctx->bd_stage_x; DBuf<int32_t>& dM = ctx->bd_stage_m;   // kept between calls: the level loop calls this once per slab
  CK(dX.alloc(n * nslab)); CK(dM.alloc(n * (mask_per_slab ? nslab : 1)));
  CK(cudaMemcpyAsync(dX.p, X, n * nslab * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dM.p, mask, n * (mask_per_slab ? nslab : 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  if (int rc = bdrown_impl(ctx, k_ew, ni, nj, nslab, dX.p, dM.p, mask_per_slab, nb_inc, nb_smooth, pmin, pmax, pval_land, nsweeps)) return rc;
  CK(cudaMemcpyAsync(X, dX.p, n * nslab * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SOSIE_OK;
}

int sosie_bdrown_force_general(sosie_ctx* c, int32_t on) {
  NEED(c);
  ctx->force_general_bdrown = on;
  return SOSIE_OK;
}

int sosie_smoother(sosie_ctx* c, int32_t k_ew, int32_t ni, int32_t nj, int32_t nslab, float* X, int32_t nb_smooth, const int32_t* msk,
                   int32_t l_emp) {
  NEED(c);
  if (ni < 3 || nj < 3 || nslab < 1 || !X) { ctx->err = "sosie_smoother: bad arguments"; return SOSIE_ERR_ARG; }
  const size_t n = (size_t)ni * nj;
  cudaStream_t st = ctx->stream;
  DBuf<float>& dX = ctx->bd_stage_x; DBuf<int32_t>& dM = ctx->bd_stage_m;
  CK(dX.alloc(n * nslab));
  CK(cudaMemcpyAsync(dX.p, X, n * nslab * sizeof(float), cudaMemcpyHostToDevice, st));
  if (msk) { CK(dM.alloc(n)); CK(cudaMemcpyAsync(dM.p, msk, n * sizeof(int32_t), cudaMemcpyHostToDevice, st)); }
  if (int rc = smoother_impl(ctx, k_ew, ni, nj, nslab, dX.p, nb_smooth, msk ? dM.p : nullptr, l_emp)) return rc;
  CK(cudaMemcpyAsync(X, dX.p, n * nslab * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SOSIE_OK;
}

int sosie_drown(sosie_ctx* c, int32_t k_ew, int32_t ni, int32_t nj, int32_t nslab, float* X, const int8_t* mask, int32_t nb_inc,
                int32_t* nsweeps) {
  NEED(c);
  if (ni < 3 || nj < 3 || nslab < 1 || !X || !mask) { ctx->err = "sosie_drown: bad arguments"; return SOSIE_ERR_ARG; }
  // the periodic index folding of src/mod_drown.f90:110-117 leaves 1..ni only when ni >= 2*k_ew + 1; below that the
  // reference reads maskv out of bounds (undefined): refused
  if (k_ew >= 0 && ni < 2 * k_ew + 1) { ctx->err = "sosie_drown: ni < 2*k_ew + 1 (the reference indexes out of bounds)"; return SOSIE_ERR_ARG; }
  const size_t n = (size_t)ni * nj;
  cudaStream_t st = ctx->stream;
  DBuf<float> dX; DBuf<int8_t> dM;
  CK(dX.alloc(n * nslab)); CK(dM.alloc(n));
  CK(cudaMemcpyAsync(dX.p, X, n * nslab * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dM.p, mask, n, cudaMemcpyHostToDevice, st));
  if (int rc = drown_impl(ctx, k_ew, ni, nj, nslab, dX.p, dM.p, nb_inc, nsweeps)) return rc;
  CK(cudaMemcpyAsync(X, dX.p, n * nslab * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SOSIE_OK;
}

// ---------------------------------------------------------------------------------------------------
int sosie_akima_1d_3d_dev(sosie_ctx* c, int32_t nx, int32_t ny, int32_t nk1, const float* vz1, const float* dxf1, int32_t nk2,
                          const float* vz2, float* dxf2, float rfill_val) {
  NEED(c);
  return akima_1d_3d_impl(ctx, nx, ny, nk1, vz1, dxf1, nk2, vz2, dxf2, rfill_val);
}

int sosie_akima_1d_3d(sosie_ctx* c, int32_t nx, int32_t ny, int32_t nk1, const float* vz1, const float* xf1, int32_t nk2, const float* vz2,
                      float* xf2, float rfill_val) {
  NEED(c);
  if (nx < 1 || ny < 1 || nk1 < 3 || nk2 < 1 || !xf1 || !xf2) { ctx->err = "sosie_akima_1d_3d: bad arguments"; return SOSIE_ERR_ARG; }
  const size_t n = (size_t)nx * ny;
  cudaStream_t st = ctx->stream;
  CK(ctx->v_in.alloc(n * nk1)); CK(ctx->v_out.alloc(n * nk2));
  CK(cudaMemcpyAsync(ctx->v_in.p, xf1, n * nk1 * sizeof(float), cudaMemcpyHostToDevice, st));
  if (int rc = akima_1d_3d_impl(ctx, nx, ny, nk1, vz1, ctx->v_in.p, nk2, vz2, ctx->v_out.p, rfill_val)) return rc;
  CK(cudaMemcpyAsync(xf2, ctx->v_out.p, n * nk2 * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SOSIE_OK;
}

int sosie_akima_1d_1d_dev(sosie_ctx* c, int64_t ncol, int32_t nk1, const double* dvz1, int64_t vz1_cs, int64_t vz1_ks, const double* dvf1,
                          int64_t vf1_cs, int64_t vf1_ks, int32_t nk2, const double* dvz2, int64_t vz2_cs, int64_t vz2_ks, float* dvf2,
                          int64_t vf2_cs, int64_t vf2_ks, int32_t has_rmask, double rmask_val, int32_t l_surf_persist, int32_t l_botm_persist,
                          int32_t* dstatus) {
  NEED(c);
  if (vz1_cs < 0 || vz1_ks < 0 || vf1_cs < 0 || vf1_ks < 0 || vz2_cs < 0 || vz2_ks < 0 || vf2_cs < 0 || vf2_ks < 0) {
    ctx->err = "sosie_akima_1d_1d: negative stride"; return SOSIE_ERR_ARG;
  }
  return akima_1d_1d_impl(ctx, ncol, nk1, dvz1, vz1_cs, vz1_ks, dvf1, vf1_cs, vf1_ks, nk2, dvz2, vz2_cs, vz2_ks, dvf2, vf2_cs, vf2_ks, has_rmask,
                          rmask_val, l_surf_persist, l_botm_persist, dstatus);
}

int sosie_akima_1d_1d(sosie_ctx* c, int64_t ncol, int32_t nk1, const double* vz1, int64_t vz1_cs, int64_t vz1_ks, const double* vf1,
                      int64_t vf1_cs, int64_t vf1_ks, int32_t nk2, const double* vz2, int64_t vz2_cs, int64_t vz2_ks, float* vf2, int64_t vf2_cs,
                      int64_t vf2_ks, int32_t has_rmask, double rmask_val, int32_t l_surf_persist, int32_t l_botm_persist, int32_t* status) {
  NEED(c);
  if (ncol < 1 || nk1 < 1 || nk2 < 1 || !vz1 || !vf1 || !vz2 || !vf2 || vz1_cs < 0 || vz1_ks < 0 || vf1_cs < 0 || vf1_ks < 0 || vz2_cs < 0 ||
      vz2_ks < 0 || vf2_cs < 0 || vf2_ks < 0) { ctx->err = "sosie_akima_1d_1d: bad arguments"; return SOSIE_ERR_ARG; }
  auto extent = [&](int64_t cs, int64_t ks, int nk) { return (size_t)((ncol - 1) * cs + (int64_t)(nk - 1) * ks + 1); };
  const size_t e1 = extent(vz1_cs, vz1_ks, nk1), e2 = extent(vf1_cs, vf1_ks, nk1), e3 = extent(vz2_cs, vz2_ks, nk2), e4 = extent(vf2_cs, vf2_ks, nk2);
  cudaStream_t st = ctx->stream;
  CK(ctx->v_d1.alloc(e1)); CK(ctx->v_d2.alloc(e2)); CK(ctx->v_d3.alloc(e3)); CK(ctx->v_out.alloc(e4));
  if (status) CK(ctx->v_mask.alloc((size_t)ncol));
  CK(cudaMemcpyAsync(ctx->v_d1.p, vz1, e1 * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->v_d2.p, vf1, e2 * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->v_d3.p, vz2, e3 * sizeof(double), cudaMemcpyHostToDevice, st));
  // strided outputs: elements of vf2 outside the columns keep the caller's values
  if (!(vf2_ks == 1 && vf2_cs == nk2) && !(vf2_cs == 1 && vf2_ks == ncol)) CK(cudaMemcpyAsync(ctx->v_out.p, vf2, e4 * sizeof(float), cudaMemcpyHostToDevice, st));
  int rc = akima_1d_1d_impl(ctx, ncol, nk1, ctx->v_d1.p, vz1_cs, vz1_ks, ctx->v_d2.p, vf1_cs, vf1_ks, nk2, ctx->v_d3.p, vz2_cs, vz2_ks, ctx->v_out.p,
                            vf2_cs, vf2_ks, has_rmask, rmask_val, l_surf_persist, l_botm_persist, status ? ctx->v_mask.p : nullptr);
  if (rc != SOSIE_OK && rc != SOSIE_ERR_REF_STOP) return rc;
  CK(cudaMemcpyAsync(vf2, ctx->v_out.p, e4 * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (status) CK(cudaMemcpyAsync(status, ctx->v_mask.p, (size_t)ncol * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return rc;
}

int sosie_interp3d_sigma_dev(sosie_ctx* c, int32_t ni, int32_t nj, int32_t nk_src, const float* ddepth_src, const float* ddata3d_tmp,
                             int32_t nk_trg, const float* ddepth_trg, const int32_t* dmask_trg, int32_t lmout, int32_t src_is_sigma,
                             int32_t trg_is_sigma, float* ddata3d_trg) {
  NEED(c);
  return interp3d_sigma_impl(ctx, ni, nj, nk_src, ddepth_src, ddata3d_tmp, nk_trg, ddepth_trg, dmask_trg, lmout, src_is_sigma, trg_is_sigma,
                             ddata3d_trg);
}

int sosie_interp3d_sigma(sosie_ctx* c, int32_t ni, int32_t nj, int32_t nk_src, const float* depth_src, const float* data3d_tmp, int32_t nk_trg,
                         const float* depth_trg, const int32_t* mask_trg, int32_t lmout, int32_t src_is_sigma, int32_t trg_is_sigma,
                         float* data3d_trg) {
  NEED(c);
  if (ni < 1 || nj < 1 || nk_src < 1 || nk_trg < 1 || !depth_src || !data3d_tmp || !depth_trg || !data3d_trg || (lmout && !mask_trg)) {
    ctx->err = "sosie_interp3d_sigma: bad arguments"; return SOSIE_ERR_ARG;
  }
  const size_t n = (size_t)ni * nj;
  cudaStream_t st = ctx->stream;
  CK(ctx->v_f1.alloc(n * nk_src)); CK(ctx->v_f2.alloc(n * nk_src)); CK(ctx->v_f3.alloc(n * nk_trg)); CK(ctx->v_out.alloc(n * nk_trg));
  CK(cudaMemcpyAsync(ctx->v_f1.p, depth_src, n * nk_src * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->v_f2.p, data3d_tmp, n * nk_src * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->v_f3.p, depth_trg, n * nk_trg * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->v_out.p, data3d_trg, n * nk_trg * sizeof(float), cudaMemcpyHostToDevice, st));
  const int32_t* dm = nullptr;
  if (mask_trg) {
    CK(ctx->v_mask.alloc(n * nk_trg));
    CK(cudaMemcpyAsync(ctx->v_mask.p, mask_trg, n * nk_trg * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    dm = ctx->v_mask.p;
  }
  if (int rc = interp3d_sigma_impl(ctx, ni, nj, nk_src, ctx->v_f1.p, ctx->v_f2.p, nk_trg, ctx->v_f3.p, dm, lmout, src_is_sigma, trg_is_sigma,
                                   ctx->v_out.p)) return rc;
  CK(cudaMemcpyAsync(data3d_trg, ctx->v_out.p, n * nk_trg * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SOSIE_OK;
}

int sosie_lbc_lnk_dev(sosie_ctx* c, int32_t nperio, int32_t jpi, int32_t jpj, int32_t nslab, float* dX, int32_t cd_type, double psgn,
                      int32_t has_pval, double pval) {
  NEED(c);
  return lbc_lnk_impl(ctx, nperio, jpi, jpj, nslab, dX, cd_type, psgn, has_pval, pval);
}

int sosie_lbc_lnk(sosie_ctx* c, int32_t nperio, int32_t jpi, int32_t jpj, int32_t nslab, float* X, int32_t cd_type, double psgn,
                  int32_t has_pval, double pval) {
  NEED(c);
  if (jpi < 4 || jpj < 4 || nslab < 1 || !X) { ctx->err = "sosie_lbc_lnk: bad arguments"; return SOSIE_ERR_ARG; }
  const size_t tot = (size_t)jpi * jpj * nslab;
  cudaStream_t st = ctx->stream;
  CK(ctx->v_in.alloc(tot));
  CK(cudaMemcpyAsync(ctx->v_in.p, X, tot * sizeof(float), cudaMemcpyHostToDevice, st));
  if (int rc = lbc_lnk_impl(ctx, nperio, jpi, jpj, nslab, ctx->v_in.p, cd_type, psgn, has_pval, pval)) return rc;
  CK(cudaMemcpyAsync(X, ctx->v_in.p, tot * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SOSIE_OK;
}

int sosie_interp3d_post_dev(sosie_ctx* c, int32_t ni, int32_t nj, int32_t nk, float* dX, const int32_t* dmask_trg, int32_t ixtrpl_bot,
                            float vmin, float vmax, float rmiss_val, int32_t i_orca_trg, int32_t c_orca_trg, int32_t lmout) {
  NEED(c);
  return interp3d_post_impl(ctx, ni, nj, nk, dX, dmask_trg, ixtrpl_bot, vmin, vmax, rmiss_val, i_orca_trg, c_orca_trg, lmout);
}

int sosie_interp3d_post(sosie_ctx* c, int32_t ni, int32_t nj, int32_t nk, float* X, const int32_t* mask_trg, int32_t ixtrpl_bot, float vmin,
                        float vmax, float rmiss_val, int32_t i_orca_trg, int32_t c_orca_trg, int32_t lmout) {
  NEED(c);
  if (ni < 1 || nj < 1 || nk < 1 || !X) { ctx->err = "sosie_interp3d_post: bad arguments"; return SOSIE_ERR_ARG; }
  const size_t tot = (size_t)ni * nj * nk;
  cudaStream_t st = ctx->stream;
  CK(ctx->v_in.alloc(tot));
  CK(cudaMemcpyAsync(ctx->v_in.p, X, tot * sizeof(float), cudaMemcpyHostToDevice, st));
  const int32_t* dm = nullptr;
  if (mask_trg) {
    CK(ctx->v_mask.alloc(tot));
    CK(cudaMemcpyAsync(ctx->v_mask.p, mask_trg, tot * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    dm = ctx->v_mask.p;
  }
  if (int rc = interp3d_post_impl(ctx, ni, nj, nk, ctx->v_in.p, dm, ixtrpl_bot, vmin, vmax, rmiss_val, i_orca_trg, c_orca_trg, lmout)) return rc;
  CK(cudaMemcpyAsync(X, ctx->v_in.p, tot * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SOSIE_OK;
}

int sosie_angle2(sosie_ctx* c, int32_t nperio, int32_t nx, int32_t ny, const double* plamt, const double* pphit, const double* plamu,
                 const double* pphiu, const double* plamv, const double* pphiv, const double* plamf, const double* pphif, double* gcost,
                 double* gsint, double* gcosu, double* gsinu, double* gcosv, double* gsinv, double* gcosf, double* gsinf) {
  NEED(c);
  const double* in[8] = {plamt, pphit, plamu, pphiu, plamv, pphiv, plamf, pphif};
  double* out[8] = {gcost, gsint, gcosu, gsinu, gcosv, gsinv, gcosf, gsinf};
  for (int q = 0; q < 8; q++) if (!in[q] || !out[q]) { ctx->err = "sosie_angle2: null pointer"; return SOSIE_ERR_ARG; }
  if (nx < 4 || ny < 5) { ctx->err = "sosie_angle2: bad extents"; return SOSIE_ERR_ARG; }
  const size_t n = (size_t)nx * ny;
  cudaStream_t st = ctx->stream;
  CK(ctx->v_ang.alloc(16 * n));
  for (int q = 0; q < 8; q++) CK(cudaMemcpyAsync(ctx->v_ang.p + (size_t)q * n, in[q], n * sizeof(double), cudaMemcpyHostToDevice, st));
  if (int rc = angle2_impl(ctx, nperio, nx, ny, ctx->v_ang.p, ctx->v_ang.p + 8 * n)) return rc;
  for (int q = 0; q < 8; q++) CK(cudaMemcpyAsync(out[q], ctx->v_ang.p + (size_t)(8 + q) * n, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SOSIE_OK;
}

int sosie_angle2_dev(sosie_ctx* c, int32_t nperio, int32_t nx, int32_t ny, const double* dcoords, double* dangles) {
  NEED(c);
  return angle2_impl(ctx, nperio, nx, ny, dcoords, dangles);
}

int sosie_extrp_hl_dev(sosie_ctx* c, int32_t ni, int32_t nj, int32_t nslab, float* dX, int32_t jj_ex_top, int32_t jj_ex_btm, int32_t nlat_icr_trg) {
  NEED(c);
  return extrp_hl_impl(ctx, ni, nj, nslab, dX, jj_ex_top, jj_ex_btm, nlat_icr_trg);
}

int sosie_extrp_hl(sosie_ctx* c, int32_t ni, int32_t nj, int32_t nslab, float* X, int32_t jj_ex_top, int32_t jj_ex_btm, int32_t nlat_icr_trg) {
  NEED(c);
  if (ni < 1 || nj < 2 || nslab < 1 || !X) { ctx->err = "sosie_extrp_hl: bad arguments"; return SOSIE_ERR_ARG; }
  const size_t tot = (size_t)ni * nj * nslab;
  cudaStream_t st = ctx->stream;
  CK(ctx->v_in.alloc(tot));
  CK(cudaMemcpyAsync(ctx->v_in.p, X, tot * sizeof(float), cudaMemcpyHostToDevice, st));
  if (int rc = extrp_hl_impl(ctx, ni, nj, nslab, ctx->v_in.p, jj_ex_top, jj_ex_btm, nlat_icr_trg)) return rc;
  CK(cudaMemcpyAsync(X, ctx->v_in.p, tot * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SOSIE_OK;
}

int sosie_rotate_uv_dev(sosie_ctx* c, int64_t n, int32_t nslab, const float* dU, const float* dV, const double* dxcos,
                        const double* dxsin, float* dUo, float* dVo, int32_t inverse) {
  NEED(c);
  return rotate_impl(ctx, n, nslab, dU, dV, dxcos, dxsin, dUo, dVo, inverse);
}

int sosie_rotate_uv(sosie_ctx* c, int64_t n, int32_t nslab, const float* U, const float* V, const double* xcos, const double* xsin,
                    float* Uo, float* Vo, int32_t inverse) {
  NEED(c);
  if (n <= 0 || nslab <= 0) return SOSIE_OK;
  cudaStream_t st = ctx->stream;
  size_t tot = (size_t)n * nslab;
  DBuf<float> dU, dV, dUo, dVo; DBuf<double> dc, ds;
  CK(dU.alloc(tot)); CK(dV.alloc(tot)); CK(dUo.alloc(tot)); CK(dVo.alloc(tot)); CK(dc.alloc(n)); CK(ds.alloc(n));
  CK(cudaMemcpyAsync(dU.p, U, tot * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dV.p, V, tot * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dc.p, xcos, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ds.p, xsin, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  if (int rc = rotate_impl(ctx, n, nslab, dU.p, dV.p, dc.p, ds.p, dUo.p, dVo.p, inverse)) return rc;
  CK(cudaMemcpyAsync(Uo, dUo.p, tot * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(Vo, dVo.p, tot * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SOSIE_OK;
}

}  // extern "C"
