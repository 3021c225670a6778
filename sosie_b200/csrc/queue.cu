// queue.cu -- asynchronous, coalescing front end of the per-record / per-level calls ("..._enqueue / _flush").
//
// The reference driver calls INTERP_2D once per time record (src/sosie.f90:133-165) and, in 3-D jobs, BDROWN and then
// bilin_2d / akima_2d once per level (src/mod_interp.f90:203-265, 288-337): one slab per call.  A synchronous one-slab call
// pays a host round trip per slab and launches kernels that fill one SM (BDROWN on an ORCA slab) or a fraction of the GPU.
// Here a call ENQUEUES its slab and returns: the slab goes up on a copy stream at once (straight from the caller's array when
// that is page-locked, else through a pinned staging ring), up to `depth` queued slabs of the same operation are
// coalesced into ONE batched launch (the kernels the batched entries use: k_apply_staged, k_bdrown_smem over `depth` CTAs,
// the batched Akima), the results come back on a third stream, and the caller's output arrays are complete after
// sosie_flush (or after any synchronous entry of the same context, which flushes first).  Two batches are in flight: the
// uploads of batch n+1 and the downloads of batch n-1 overlap the kernels of batch n.
//
// Contract: a page-locked input (sosie_host_alloc / sosie_host_register) is BORROWED until the flush; a pageable input has
// been copied when the call returns and may be reused at once.  Output arrays must stay valid until the flush.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace {

enum { Q_NONE = 0, Q_BILIN = 1, Q_AKIMA = 2, Q_BDROWN = 3 };

// one host -> device copy of a batch; small slabs (an ORCA level is ~100 KB) are not copied one by one at enqueue time but
// merged at submit time: consecutive levels of a cube are adjacent in host memory and in the batch buffer, so a whole
// level loop goes up (and comes back) in one or two cudaMemcpyAsync calls instead of 62
struct QCopy { void* d; const void* h; size_t bytes; };
#define Q_DEFER_BYTES ((size_t)1 << 20)

struct QItem {
  float* dst = nullptr;        // caller's output slab
  float* stage_out = nullptr;  // pinned staging the result lands in when dst is pageable (else nullptr: DMA straight to dst)
  int32_t* nsw_dst = nullptr;  // BDROWN: sweeps executed (optional)
};

struct BdParams {
  int k_ew = 0, ni = 0, nj = 0, nb_inc = 0, nb_smooth = 0;
  float pmin = 0.f, pmax = 0.f, pval = 0.f;
  bool operator==(const BdParams& o) const {
    return k_ew == o.k_ew && ni == o.ni && nj == o.nj && nb_inc == o.nb_inc && nb_smooth == o.nb_smooth && pmin == o.pmin && pmax == o.pmax &&
           pval == o.pval;
  }
};

struct QBatch {
  int op = Q_NONE;
  BdParams bd;
  size_t n_in = 0, n_out = 0;   // elements per slab
  std::vector<QItem> items;
  std::vector<QCopy> up;        // deferred uploads (small slabs), issued merged at submit
  bool open = false, in_flight = false;
  cudaEvent_t e_in = nullptr, e_done = nullptr, e_out = nullptr;
  DBuf<float> d_in, d_out;
  DBuf<int32_t> d_mask;
  float* h_in = nullptr;  size_t h_in_n = 0;     // pinned staging rings (allocated on first need)
  float* h_out = nullptr; size_t h_out_n = 0;
  int32_t* h_mask = nullptr; size_t h_mask_n = 0;
  std::vector<int32_t> nsw;
};

}  // namespace

struct SosieQueue {
  int depth = 32;
  int cap = 0;                  // slabs per batch of the open configuration (depth clamped by the memory budget)
  cudaStream_t h2d = nullptr, d2h = nullptr;
  QBatch b[2];
  int cur = 0;                  // batch being filled (or to be filled next)
  int64_t n_batches = 0, n_slabs = 0, n_staged_in = 0, n_staged_out = 0;
};

static int pinned_grow(sosie_ctx* ctx, void** p, size_t* have, size_t want_bytes) {
  if (*p && *have >= want_bytes) return SOSIE_OK;
  if (*p) cudaFreeHost(*p);
  *p = nullptr; *have = 0;
  CK(cudaHostAlloc(p, want_bytes, cudaHostAllocDefault));
  *have = want_bytes;
  return SOSIE_OK;
}

// page-locked ranges this library created or registered itself (sosie_host_alloc / sosie_host_register): looked up without
// a driver call; anything else is asked of the driver (cudaPointerGetAttributes, ~1 us)
namespace {
struct PinRange { const char* lo; const char* hi; };
std::mutex g_pin_mu;
std::vector<PinRange> g_pins;
}  // namespace
void pin_registry_add(const void* p, size_t bytes) {
  std::lock_guard<std::mutex> g(g_pin_mu);
  g_pins.push_back({(const char*)p, (const char*)p + bytes});
}
void pin_registry_remove(const void* p) {
  std::lock_guard<std::mutex> g(g_pin_mu);
  for (size_t i = 0; i < g_pins.size(); i++)
    if (g_pins[i].lo == (const char*)p) { g_pins.erase(g_pins.begin() + i); return; }
}
static bool is_pinned(const void* p) {
  {
    std::lock_guard<std::mutex> g(g_pin_mu);
    for (const PinRange& r : g_pins)
      if ((const char*)p >= r.lo && (const char*)p < r.hi) return true;
  }
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

static int queue_get(sosie_ctx* ctx, SosieQueue** out) {
  if (!ctx->q) {
    SosieQueue* q = new SosieQueue();
    if (const char* e = getenv("SOSIE_QUEUE_DEPTH")) { int d = atoi(e); if (d >= 1 && d <= 4096) q->depth = d; }
    ctx->q = q;
    CK(cudaStreamCreateWithFlags(&q->h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&q->d2h, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
      CK(cudaEventCreateWithFlags(&q->b[k].e_in, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&q->b[k].e_done, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&q->b[k].e_out, cudaEventDisableTiming));
    }
  }
  *out = ctx->q;
  return SOSIE_OK;
}

void queue_destroy(sosie_ctx* ctx) {
  SosieQueue* q = ctx->q;
  if (!q) return;
  for (int k = 0; k < 2; k++) {
    QBatch& B = q->b[k];
    if (B.e_in) cudaEventDestroy(B.e_in);
    if (B.e_done) cudaEventDestroy(B.e_done);
    if (B.e_out) cudaEventDestroy(B.e_out);
    if (B.h_in) cudaFreeHost(B.h_in);
    if (B.h_out) cudaFreeHost(B.h_out);
    if (B.h_mask) cudaFreeHost(B.h_mask);
  }
  if (q->h2d) cudaStreamDestroy(q->h2d);
  if (q->d2h) cudaStreamDestroy(q->d2h);
  delete q;
  ctx->q = nullptr;
}

// results of batch B are on the host: hand the staged ones over to the caller's arrays
static int queue_retire(sosie_ctx* ctx, QBatch& B) {
  if (!B.in_flight) return SOSIE_OK;
  CK(cudaEventSynchronize(B.e_out));
  size_t k0 = 0, kcnt = B.n_out;
  if (B.op == Q_BILIN) shard_range(ctx, &k0, &kcnt);
  for (size_t i = 0; i < B.items.size(); i++) {
    QItem& it = B.items[i];
    if (it.stage_out) memcpy(it.dst + k0, it.stage_out + k0, kcnt * sizeof(float));
    if (it.nsw_dst) *it.nsw_dst = B.nsw[i];
  }
  B.items.clear();
  B.in_flight = false;
  B.op = Q_NONE;
  return SOSIE_OK;
}

// launch the batched kernels of the open batch and queue its downloads
static int queue_submit(sosie_ctx* ctx, SosieQueue* q, QBatch& B) {
  if (!B.open) return SOSIE_OK;
  B.open = false;
  const int cnt = (int)B.items.size();
  if (cnt == 0) { B.op = Q_NONE; return SOSIE_OK; }
  cudaStream_t sc = ctx->stream;
  for (const QCopy& c : B.up) CK(cudaMemcpyAsync(c.d, c.h, c.bytes, cudaMemcpyHostToDevice, q->h2d));
  B.up.clear();
  CK(cudaEventRecord(B.e_in, q->h2d));
  CK(cudaStreamWaitEvent(sc, B.e_in, 0));
  float* res = B.d_out.p;
  bool want_nsw = false;
  if (B.op == Q_BILIN) {
    if (int rc = bilin_apply_impl(ctx, cnt, B.d_in.p, B.d_out.p, 0, 0, 0.f, 0.f, nullptr, 0.f)) return rc;
  } else if (B.op == Q_AKIMA) {
    if (int rc = akima_apply_impl(ctx, cnt, B.d_in.p, B.d_out.p)) return rc;
  } else {
    for (auto& it : B.items) want_nsw |= it.nsw_dst != nullptr;
    B.nsw.assign(cnt, 0);
    if (int rc = bdrown_impl(ctx, B.bd.k_ew, B.bd.ni, B.bd.nj, cnt, B.d_in.p, B.d_mask.p, 1, B.bd.nb_inc, B.bd.nb_smooth, B.bd.pmin, B.bd.pmax,
                             B.bd.pval, want_nsw ? B.nsw.data() : nullptr)) return rc;
    res = B.d_in.p;   // in place
  }
  CK(cudaEventRecord(B.e_done, sc));
  CK(cudaStreamWaitEvent(q->d2h, B.e_done, 0));
  size_t k0 = 0, kcnt = B.n_out;
  if (B.op == Q_BILIN) shard_range(ctx, &k0, &kcnt);   // a sharded context owns (and returns) only its target rows
  for (int i = 0; i < cnt && kcnt; ) {
    // results that are adjacent on both sides (consecutive levels of a cube, whole slabs) come back in one copy
    QItem& it = B.items[i];
    float* hd = (it.stage_out ? it.stage_out : it.dst) + k0;
    size_t run = kcnt;
    int j = i + 1;
    if (kcnt == B.n_out)
      for (; j < cnt; j++) {
        QItem& nx = B.items[j];
        if ((nx.stage_out ? nx.stage_out : nx.dst) != hd + run) break;
        run += kcnt;
      }
    CK(cudaMemcpyAsync(hd, res + (size_t)i * B.n_out + k0, run * sizeof(float), cudaMemcpyDeviceToHost, q->d2h));
    i = j;
  }
  CK(cudaEventRecord(B.e_out, q->d2h));
  B.in_flight = true;
  q->n_batches++;
  q->cur ^= 1;
  return SOSIE_OK;
}

int queue_flush(sosie_ctx* ctx) {
  SosieQueue* q = ctx->q;
  if (!q) return SOSIE_OK;
  // retire in submission order (two enqueues may name the same output array): with a batch open, the other one is the
  // older; with none open, the one `cur` points at was submitted before the other
  const int c0 = q->cur;
  const bool was_open = q->b[c0].open;
  int rc = SOSIE_OK;
  if (was_open) { if ((rc = queue_submit(ctx, q, q->b[c0]))) return rc; }
  QBatch& first = was_open ? q->b[c0 ^ 1] : q->b[c0];
  QBatch& second = was_open ? q->b[c0] : q->b[c0 ^ 1];
  if ((rc = queue_retire(ctx, first))) return rc;
  if ((rc = queue_retire(ctx, second))) return rc;
  return SOSIE_OK;
}

bool queue_pending(const sosie_ctx* ctx) {
  const SosieQueue* q = ctx->q;
  return q && (q->b[0].open || q->b[0].in_flight || q->b[1].open || q->b[1].in_flight);
}

// open a batch for (op, params) -- or return the open one when it matches and has room
static int queue_open(sosie_ctx* ctx, SosieQueue* q, int op, const BdParams* bd, size_t n_in, size_t n_out, QBatch** out) {
  QBatch* B = &q->b[q->cur];
  if (B->open) {
    const bool same = B->op == op && B->n_in == n_in && B->n_out == n_out && (op != Q_BDROWN || B->bd == *bd);
    if (same && (int)B->items.size() < q->cap) { *out = B; return SOSIE_OK; }
    if (int rc = queue_submit(ctx, q, *B)) return rc;   // flips q->cur
    B = &q->b[q->cur];
  }
  if (int rc = queue_retire(ctx, *B)) return rc;       // its buffers are about to be overwritten
  const size_t budget = (size_t)1 << 31;               // bytes of device staging per batch
  const size_t per = (n_in + (op == Q_BDROWN ? n_in : n_out)) * sizeof(float);
  q->cap = (int)std::max<size_t>(1, std::min<size_t>((size_t)q->depth, budget / std::max<size_t>(per, 1)));
  B->op = op; B->n_in = n_in; B->n_out = n_out;
  if (bd) B->bd = *bd;
  CK(B->d_in.alloc(n_in * q->cap));
  if (op == Q_BDROWN) CK(B->d_mask.alloc(n_in * q->cap));
  else CK(B->d_out.alloc(n_out * q->cap));
  B->items.clear();
  B->up.clear();
  B->open = true;
  *out = B;
  return SOSIE_OK;
}

// upload `cnt` elements starting at element `off` of the caller's slab into slot i of the batch
template <class T>
static int queue_upload(sosie_ctx* ctx, SosieQueue* q, T* dslot, const T* src, size_t off, size_t cnt, T** hring, size_t* hring_n,
                        size_t slab_elems, int slot) {
  if (cnt == 0) return SOSIE_OK;
  const T* hsrc = src;
  if (!is_pinned(src)) {
    if (int rc = pinned_grow(ctx, (void**)hring, hring_n, slab_elems * (size_t)q->cap * sizeof(T))) return rc;
    T* hs = *hring + (size_t)slot * slab_elems;
    memcpy(hs + off, src + off, cnt * sizeof(T));
    q->n_staged_in++;
    hsrc = hs;
  }
  if (cnt * sizeof(T) < Q_DEFER_BYTES) {
    std::vector<QCopy>& up = q->b[q->cur].up;
    if (!up.empty() && (const char*)up.back().h + up.back().bytes == (const char*)(hsrc + off) &&
        (char*)up.back().d + up.back().bytes == (char*)(dslot + off)) up.back().bytes += cnt * sizeof(T);
    else up.push_back({dslot + off, hsrc + off, cnt * sizeof(T)});
    return SOSIE_OK;
  }
  CK(cudaMemcpyAsync(dslot + off, hsrc + off, cnt * sizeof(T), cudaMemcpyHostToDevice, q->h2d));
  return SOSIE_OK;
}

static int queue_item_out(sosie_ctx* ctx, SosieQueue* q, QBatch* B, float* dst, int32_t* nsw_dst) {
  QItem it;
  it.dst = dst; it.nsw_dst = nsw_dst;
  if (!is_pinned(dst)) {
    if (int rc = pinned_grow(ctx, (void**)&B->h_out, &B->h_out_n, B->n_out * (size_t)q->cap * sizeof(float))) return rc;
    it.stage_out = B->h_out + B->items.size() * B->n_out;
    q->n_staged_out++;
  }
  B->items.push_back(it);
  q->n_slabs++;
  if ((int)B->items.size() >= q->cap) return queue_submit(ctx, q, *B);
  return SOSIE_OK;
}

#define NEEDQ(ctx_)                                  \
  if (!(ctx_)) return SOSIE_ERR_ARG;                 \
  sosie_ctx* ctx = (ctx_);                           \
  if (cudaSetDevice(ctx->device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return SOSIE_ERR_CUDA; } \
  SosieQueue* q = nullptr;                           \
  if (int rcq__ = queue_get(ctx, &q)) return rcq__;

extern "C" {

int sosie_set_queue_depth(sosie_ctx* c, int32_t depth) {
  NEEDQ(c);
  if (depth < 1 || depth > 4096) { ctx->err = "sosie_set_queue_depth: depth must be in 1..4096"; return SOSIE_ERR_ARG; }
  if (int rc = queue_flush(ctx)) return rc;
  q->depth = depth;
  return SOSIE_OK;
}

int sosie_flush(sosie_ctx* c) {
  if (!c) return SOSIE_ERR_ARG;
  sosie_ctx* ctx = c;
  if (cudaSetDevice(ctx->device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return SOSIE_ERR_CUDA; }
  return queue_flush(ctx);
}

int sosie_queue_stats(sosie_ctx* c, int64_t* stats) {
  if (!c || !stats) return SOSIE_ERR_ARG;
  SosieQueue* q = c->q;
  stats[0] = q ? q->n_batches : 0; stats[1] = q ? q->n_slabs : 0; stats[2] = q ? q->n_staged_in : 0; stats[3] = q ? q->n_staged_out : 0;
  return SOSIE_OK;
}

int sosie_bilin_apply_enqueue(sosie_ctx* c, const float* Z1, float* Z2) {
  NEEDQ(c);
  if (!Z1 || !Z2) { ctx->err = "sosie_bilin_apply_enqueue: null pointer"; return SOSIE_ERR_ARG; }
  if (!ctx->grids_set || !ctx->map_ready) { ctx->err = "sosie_bilin_apply_enqueue: no grids / mapping (the first call of a job is synchronous)"; return SOSIE_ERR_STATE; }
  if (!ctx->pack_ready) { if (int rc = bilin_pack_impl(ctx)) return rc; }
  size_t k0, kcnt;
  shard_range(ctx, &k0, &kcnt);
  if (ctx->rows_gen != ctx->pack_gen) {
    if (int rc = bilin_src_rowrange(ctx, k0, kcnt, &ctx->src_row_lo, &ctx->src_row_hi)) return rc;
    ctx->rows_gen = ctx->pack_gen;
  }
  const size_t n1 = (size_t)ctx->sg.nx * ctx->sg.ny, n2 = (size_t)ctx->tg.nx * ctx->tg.ny;
  QBatch* B = nullptr;
  if (int rc = queue_open(ctx, q, Q_BILIN, nullptr, n1, n2, &B)) return rc;
  const int rlo = ctx->src_row_lo, rhi = ctx->src_row_hi, slot = (int)B->items.size();
  const size_t soff = rhi >= rlo ? (size_t)rlo * ctx->sg.nx : 0, scnt = rhi >= rlo ? (size_t)(rhi - rlo + 1) * ctx->sg.nx : 0;
  if (int rc = queue_upload<float>(ctx, q, B->d_in.p + (size_t)slot * n1, Z1, soff, scnt, &B->h_in, &B->h_in_n, n1, slot)) return rc;
  return queue_item_out(ctx, q, B, Z2, nullptr);
}

int sosie_akima_apply_enqueue(sosie_ctx* c, const float* Z1, float* Z2) {
  NEEDQ(c);
  if (!Z1 || !Z2) { ctx->err = "sosie_akima_apply_enqueue: null pointer"; return SOSIE_ERR_ARG; }
  if (!ctx->ak_set) { ctx->err = "sosie_akima_apply_enqueue: grids not set"; return SOSIE_ERR_STATE; }
  const size_t n1 = (size_t)ctx->ak_nx1 * ctx->ak_ny1, n2 = (size_t)ctx->ak_nx2 * ctx->ak_ny2;
  QBatch* B = nullptr;
  if (int rc = queue_open(ctx, q, Q_AKIMA, nullptr, n1, n2, &B)) return rc;
  const int slot = (int)B->items.size();
  if (int rc = queue_upload<float>(ctx, q, B->d_in.p + (size_t)slot * n1, Z1, 0, n1, &B->h_in, &B->h_in_n, n1, slot)) return rc;
  return queue_item_out(ctx, q, B, Z2, nullptr);
}

int sosie_bdrown_enqueue(sosie_ctx* c, int32_t k_ew, int32_t ni, int32_t nj, float* X, const int32_t* mask, int32_t nb_inc,
                         int32_t nb_smooth, float pmin, float pmax, float pval_land, int32_t* nsweeps) {
  NEEDQ(c);
  if (ni < 3 || nj < 3 || !X || !mask) { ctx->err = "sosie_bdrown_enqueue: bad arguments"; return SOSIE_ERR_ARG; }
  BdParams bd;
  bd.k_ew = k_ew; bd.ni = ni; bd.nj = nj; bd.nb_inc = nb_inc; bd.nb_smooth = nb_smooth; bd.pmin = pmin; bd.pmax = pmax; bd.pval = pval_land;
  const size_t n = (size_t)ni * nj;
  QBatch* B = nullptr;
  if (int rc = queue_open(ctx, q, Q_BDROWN, &bd, n, n, &B)) return rc;
  const int slot = (int)B->items.size();
  if (int rc = queue_upload<float>(ctx, q, B->d_in.p + (size_t)slot * n, X, 0, n, &B->h_in, &B->h_in_n, n, slot)) return rc;
  if (int rc = queue_upload<int32_t>(ctx, q, B->d_mask.p + (size_t)slot * n, mask, 0, n, &B->h_mask, &B->h_mask_n, n, slot)) return rc;
  return queue_item_out(ctx, q, B, X, nsweeps);
}

void* sosie_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  pin_registry_add(p, bytes);
  return p;
}
int sosie_host_free(void* p) {
  if (!p) return SOSIE_OK;
  pin_registry_remove(p);
  return cudaFreeHost(p) == cudaSuccess ? SOSIE_OK : SOSIE_ERR_CUDA;
}

}  // extern "C"
