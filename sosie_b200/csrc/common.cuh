// common.cuh -- context, device buffers and launch helpers shared by all translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/sosie_b200.h"
#include "pmath.h"

#define SOSIE_NUM_SMS_B200 148
#define SOSIE_T_COUNT 8
#define SOSIE_ISCRATCH 256   // ints: [0] error flags, [1] sortedness probe, [2..3] source row range, [8..] per-chunk error flags

template <class T>
struct DBuf {  // owning device buffer (or borrowed pointer when own == false)
  T* p = nullptr;
  size_t n = 0;
  bool own = true;
  cudaError_t alloc(size_t count) {
    if (own && p && n >= count && count > 0) return cudaSuccess;
    release();
    own = true;
    n = count;
    if (count == 0) return cudaSuccess;
    return cudaMalloc((void**)&p, count * sizeof(T));
  }
  void borrow(const T* q, size_t count) {
    release();
    p = const_cast<T*>(q);
    n = count;
    own = false;
  }
  void release() {
    if (own && p) cudaFree(p);
    p = nullptr;
    n = 0;
    own = true;
  }
  ~DBuf() { release(); }
};

// Source grid geometry as the kernels see it (1-based accessors below).
struct SrcGeom {
  int nx = 0, ny = 0;      // original extents
  int nyw = 0;             // working extent in j (ny + 2 when the pole rows are added)
  int separable = 0;       // 1: 1-D axes + trig tables; 0: full 2-D arrays (nx,nyw)
  const double* lon1 = nullptr;  // [nx]      (separable)
  const double* lat1 = nullptr;  // [nyw]
  const double* clon = nullptr;  // cos/sin of lon*zconv  [nx]
  const double* slon = nullptr;
  const double* clat = nullptr;  // cos/sin of lat*zconv  [nyw]
  const double* slat = nullptr;
  const double* X = nullptr;     // [nx*nyw]  (non separable)
  const double* Y = nullptr;
  const double* vx = nullptr;    // unit vectors [nx*nyw]
  const double* vy = nullptr;
  const double* vz = nullptr;
  const double* merc = nullptr;  // Mercator ordinate of heading() per source latitude: [nyw] (separable) or [nx*nyw]
  // separable only: the four neighbour headings of MAPPING_BL (:512-515) depend on the column (E, W) or the row
  // (N, S) of the nearest point alone; tabulated once per grid with the same device function (identical bits)
  const double* hE = nullptr;    // [nx]
  const double* hW = nullptr;    // [nx]
  const double* hN = nullptr;    // [nyw]
  const double* hS = nullptr;    // [nyw]
  const double* lonm = nullptr;  // [nx] MOD(lon1, 360.) as the body forms it (:495-499)
  int lonm_same = 0;             // 1: every lon1 already satisfies |lon| < 360, MOD() changed nothing
  // uniform-axis hints of the EASY bracket search (0 = none): index guess = (r - v0) * inv_d, verified exactly
  double lon0 = 0.0, inv_dlon = 0.0, lat0 = 0.0, inv_dlat = 0.0;
};

struct TrgGeom {
  int nx = 0, ny = 0;
  int is_1d = 0;
  const double* X = nullptr;  // [nx] or [nx*ny]
  const double* Y = nullptr;  // [ny] or [nx*ny]
};

struct SosieQueue;   // queue.cu: asynchronous enqueue / flush front end

struct sosie_ctx {
  int device = 0;
  SosieQueue* q = nullptr;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  std::string err;
  int64_t launches = 0;
  int num_sms = SOSIE_NUM_SMS_B200;
  bool attr_tma = false, attr_bdrown = false;   // opt-in shared-memory sizes set on this context's device

  // ---- bilinear state
  int k_ew = -1, l_reg_src = 0, i_orca_trg = 0, l_add_extra_j = 0;
  bool grids_set = false, map_ready = false, pack_ready = false;
  bool always_store_records = false;   // test hook (SOSIE_ALWAYS_STORE_RECORDS): the first apply stores the records too
  int unpacked_applies = 0;      // single-slab applies served straight from the mapping arrays since it was built (see bilin_apply_impl)
  bool search_only = false;     // sosie_find_nearest_point: the search stage alone (JIp, JJp into the first two planes of d_metrics)
  size_t map_nt = 0;             // target count the cached mapping was built/loaded for
  int shard_j0 = 0, shard_j1 = 0; // 1-based inclusive target rows handled by this context (0,0 = all)
  int map_j0 = 0, map_j1 = 0;     // 1-based inclusive target rows the cached mapping holds (0,0 = every row: unsharded build, load_map)
  // prelude exchange between the ranks of a row-sharded job (sosie_set_exchange); NULL: every rank reduces the whole grid
  int (*xchg_fn)(void*, const void*, void*, size_t) = nullptr;
  void* xchg_user = nullptr;
  int xchg_nranks = 0;
  int xchg_used = 0;            // 1: the last pipelined first call / device-planned mapping took the exchange path
  // prelude exchange over PEER MEMORY (sosie_p2p_init / sosie_p2p_connect): every rank owns a mailbox its peers write their
  // records into with plain stores over NVLink -- no collective launch; 2 = the last mapping took this path
  unsigned char* p2p_box = nullptr;        // this rank's mailbox: [2 parities][nranks][p2p_slot] bytes, then [2][nranks] uint32 flags
  size_t p2p_slot = 0;
  int p2p_nranks = 0, p2p_rank = -1;
  bool p2p_ready = false;
  unsigned p2p_epoch = 0;
  long long p2p_spin_cycles = 20000000000LL;   // bound of the consumer's wait (~10 s; SOSIE_P2P_TIMEOUT_MS at sosie_p2p_init)
  void* p2p_opened[256] = {nullptr};       // peers' mailboxes opened through CUDA IPC (closed in sosie_destroy)
  unsigned char** p2p_peers_dev = nullptr; // device table of the nranks mailbox addresses (own included)
  unsigned* p2p_done = nullptr;            // block counter of the push kernel
  int (*xchg_dev_fn)(void*, const void*, void*, size_t, void*) = nullptr;   // device-side all-gather (sosie_set_exchange_dev)
  void* xchg_dev_user = nullptr;
  int xchg_dev_nranks = 0;
  bool map_check_pending = false;   // a device-planned mapping whose status (errors, map_info) has not been read back yet
  bool trg_partial = false;     // the pipelined first call of a sharded context left only part of the target coordinates in HBM
  // ---- per-kernel CUDA-event timing (sosie_profile / sosie_get_timing)
  bool prof = false;
  cudaEvent_t ev0[SOSIE_T_COUNT] = {}, ev1[SOSIE_T_COUNT] = {};
  bool ev_used[SOSIE_T_COUNT] = {};
  SrcGeom sg;
  TrgGeom tg;
  DBuf<double> s_lon1, s_lat1, s_clon, s_slon, s_clat, s_slat, s_X, s_Y, s_vx, s_vy, s_vz, s_merc, s_head;
  DBuf<double> s_in_X, s_in_Y;   // uploaded (or borrowed) raw source coordinates
  DBuf<double> t_X, t_Y;
  DBuf<int32_t> t_mask;          // mask_domain_trg as modified by the search
  bool mask_is_ones = false;     // no mask was given: t_mask is NOT filled (kernels take a null mask = all 1) until somebody writes it
  DBuf<int32_t> d_metrics;       // (nx2,ny2,3)
  DBuf<double> d_ab;             // (nx2,ny2,2)
  DBuf<int16_t> d_ipb;
  DBuf<float> d_dnp;
  DBuf<int32_t> d_im;            // ji_m table of EXT_NORTH_TO_90_REGG [nx1]
  DBuf<int4> d_pidx;             // packed apply records
  DBuf<float4> d_pw;
  DBuf<int4> d_boxes;            // per-tile source bounding boxes of the staged apply
  DBuf<int32_t> d_tilelist;      // four compact lists of ntiles entries: tiles of the group-staged kernel with 8 or 4 / 2 / 1 slabs per group, the rest
  DBuf<int32_t> d_tilecnt;
  int tiles_cls[4] = {0, 0, 0, 0};   // their lengths
  bool boxes_ready = false;
  int tiles_tma = 0, tiles_grp = 0, tiles_other = 0;   // tiles that fit the TMA box / the group-staged kernel / neither (counted with the boxes)
  int tma_mode = 0;                     // 1: batched applies stage TMA-class tiles with cp.async.bulk.tensor (SOSIE_APPLY_TMA=1, sosie_apply_tma)
  int grp_spc = 0;                      // > 0: slabs per block of the group-staged kernel (tuning hook SOSIE_GRP_SPC; 0 = automatic)
  int grp_mode = 1;                     // 0: no group-staged kernel either (test hook: every variant must agree bit for bit)
  DBuf<float> d_stage1, d_stage2;  // staging for host-pointer apply
  DBuf<float> d_out;               // result staging of the host-pointer paths
  size_t pipeline_min_targets = (size_t)1 << 20;  // first calls smaller than this take the sequential path
  uint64_t pack_gen = 0, rows_gen = ~0ULL;        // generation of the packed records / of the cached source row range
  int src_row_lo = 0, src_row_hi = -1;            // physical source rows (0-based) the packed records of this shard read
  DBuf<double> d_scratch;        // Prelude + misc scalars of the search
  DBuf<int32_t> d_iscratch;      // error flag
  DBuf<double> d_vax;            // vlon/vlat axes of the EASY search
  // persistent scratch slots handed out in call order (cudaMalloc/cudaFree cost ~0.1-1 ms each and cudaFree
  // synchronises: the mapping would otherwise spend more time allocating than computing on small grids)
  std::vector<DBuf<unsigned char>*> ws_slots;
  size_t ws_cursor = 0;
  // page-locked, device-mapped "mailbox": small results (scalars, flags, axis tables) are written to it BY A KERNEL
  // and read by the host after a stream sync.  A cudaMemcpy D2H would queue behind the bulk downloads that the
  // pipelined paths keep on the copy engine (measured: a 72-byte read-back waited 16 ms for 1.2 GB of mapping).
  unsigned char* mailbox = nullptr;
  ~sosie_ctx() {
    for (auto* p : ws_slots) delete p;
    if (mailbox) cudaFreeHost(mailbox);
  }
  int32_t map_info[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int easy_sorted_lon = 0, easy_sorted_lat = 0;
  std::vector<double> h_axes;    // host copy of a 1-D source's axes: lon(nx), working lat(nyw)
  std::vector<double> axes_hint; // sosie_bilin_src_axes_hint: the caller's host copy of device-resident 1-D source axes
  bool axes_hint_dev = false;    // this set_grids call may use it (device-pointer entry only)
  bool axes_hint_verify = false; // SOSIE_VERIFY_HINTS: read the device axes back and compare
  // source tables of the previous set_grids, reusable while the hinted axes, k_ew and the pole extension stay the same
  bool tables_valid = false, tables_reused = false, easy_valid = false;
  int tables_kew = 0, tables_extra = 0, tables_lonm_same = 0;
  const double* tables_Xin = nullptr;
  std::vector<double> tables_axes;
  double easy_hint[4] = {0, 0, 0, 0};
  int l_last_y_row_missing = 0;

  // ---- akima state
  bool ak_set = false;
  int ak_kew = -1, ak_nx1 = 0, ak_ny1 = 0, ak_nx2 = 0, ak_ny2 = 0, ak_iorca = 0, ak_trg_1d = 0;
  DBuf<double> ak_zX, ak_zY;         // frame-extended coordinates (nx1+4, ny1+4)
  DBuf<double> ak_ZX8, ak_ZY8;       // twice-extended coordinates (nx1+8, ny1+8)
  DBuf<double> ak_X2, ak_Y2;         // target coordinates (2-D)
  DBuf<int32_t> ak_ixy;              // (nx2,ny2,2)
  bool ak_untiled_slopes = false;    // tests: the per-point slopes kernel instead of the shared-memory tiled one
  bool ak_sep8 = false;              // regular source: twice-extended coordinates are separable -> fused slopes + evaluation kernel
  bool ak_fused = false;             // sosie_akima_untiled_slopes(3): regular sources take the fused slopes + evaluation kernel
  DBuf<double> ak_ax8;               // the 1-D axes of the twice-extended grid: x8 (ni1+4), y8 (nj1+4)
  DBuf<uint8_t> ak_needed;           // (nx1+4,ny1+4): source points whose slopes some target reads
  DBuf<double> ak_ZF8, ak_sx, ak_sy, ak_sxy;  // per-slab work arrays
  DBuf<float> ak_stage1, ak_stage2;
  double ak_range[4] = {0, 0, 0, 0};
  DBuf<int32_t> ak_tiles;            // 32x8 tiles of the frame-extended slab that hold a needed point (k_slopes_tiled walks this list)
  int ak_ntiles_needed = -1;         // -1: no list (every tile)
  int ak_row_lo = 0, ak_row_hi = -1;  // rows of the (nx1+4,ny1+4) frame-extended slab holding needed points (1-based; empty: hi < lo)

  // ---- vertical stage (vert.cu)
  DBuf<float> v_vze, v_vz2, v_in, v_out;
  DBuf<double> v_ang, v_d1, v_d2, v_d3;
  DBuf<float> v_f1, v_f2, v_f3;
  DBuf<int32_t> v_plan, v_mask;

  // ---- drown work buffers
  DBuf<int8_t> dr_m0, dr_m1;
  DBuf<float> dr_x, dr_x2;
  DBuf<int32_t> dr_cnt, dr_mask32;
  DBuf<unsigned int> dr_Ri, dr_Re;   // land-only smoother: interior cells / periodic edge-column cells of dr_R0
  DBuf<unsigned int> dr_R0, dr_Ra, dr_Rb, dr_Lv, dr_ctr;   // level-list BDROWN: land lists, level list, counters
  DBuf<unsigned char> dr_tmp;                               // cub scratch
  DBuf<float> bd_stage_x;                                   // host-pointer BDROWN / SMOOTHER: persistent staging (one cudaMalloc per job, not per level)
  DBuf<int32_t> bd_stage_m;
  int force_general_bdrown = 0;  // tests: 0 auto, 1 level-list general path even for on-chip sized slabs, 2 legacy dense path
};

#define CK(call)                                                                      \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                 \
      return SOSIE_ERR_CUDA;                                                          \
    }                                                                                 \
  } while (0)

#define KLAUNCH_CHECK()                                                               \
  do {                                                                                \
    ctx->launches++;                                                                  \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      ctx->err = std::string("kernel launch: ") + cudaGetErrorString(e__);            \
      return SOSIE_ERR_CUDA;                                                          \
    }                                                                                 \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// exact unsigned 32-bit division by an invariant divisor (Granlund & Montgomery 1994): the list kernels decode
// (mask, j, i) from a linear cell index for every entry; three hardware divides cost more than the cell update itself
struct FastDiv {
  unsigned int d, M, s1, s2;
};
static inline FastDiv make_fastdiv(unsigned int d) {
  FastDiv f;
  f.d = d; f.M = 0; f.s1 = 0; f.s2 = 0;
  if (d <= 1) return f;
  unsigned int l = 0;
  while ((1ull << l) < d) l++;
  f.M = (unsigned int)((((1ull << l) - d) << 32) / d + 1);
  f.s1 = l < 1 ? l : 1; f.s2 = l > 1 ? l - 1 : 0;
  return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ unsigned int fdiv(unsigned int a, const FastDiv& f) {
  if (f.d <= 1) return a;
  const unsigned int t = __umulhi(f.M, a);
  return (t + ((a - t) >> f.s1)) >> f.s2;
}

#endif


// scratch slot #cursor of the context, grown to hold `count` T's (nullptr on allocation failure)
template <class T>
static inline T* ws_get(sosie_ctx* ctx, size_t count) {
  if (ctx->ws_cursor >= ctx->ws_slots.size()) ctx->ws_slots.push_back(new DBuf<unsigned char>());
  DBuf<unsigned char>* b = ctx->ws_slots[ctx->ws_cursor++];
  size_t bytes = count * sizeof(T) + 16;
  if (b->alloc(bytes) != cudaSuccess) return nullptr;
  return reinterpret_cast<T*>(b->p);
}
#define WS(T, name, count)                                                              \
  T* name##_p = ws_get<T>(ctx, (count));                                                \
  if (!name##_p) { ctx->err = "out of device memory (scratch)"; return SOSIE_ERR_CUDA; } \
  struct { T* p; } name = {name##_p}

// event brackets around the dominant kernels (ids: SOSIE_T_* in sosie_b200.h)
static inline void prof_begin(sosie_ctx* ctx, int id) {
  if (!ctx->prof) return;
  if (!ctx->ev0[id]) { cudaEventCreate(&ctx->ev0[id]); cudaEventCreate(&ctx->ev1[id]); }
  cudaEventRecord(ctx->ev0[id], ctx->stream);
}
static inline void prof_end(sosie_ctx* ctx, int id) {
  if (!ctx->prof) return;
  cudaEventRecord(ctx->ev1[id], ctx->stream);
  ctx->ev_used[id] = true;
}
// target index range [k0, k0+cnt) of this context's shard
// (sosie_bilin_set_shard and the entries that fix the target extents refuse a shard that leaves the grid: no silent
// fall-back to the whole grid)
static inline void shard_range(const sosie_ctx* ctx, size_t* k0, size_t* cnt) {
  size_t nx = (size_t)ctx->tg.nx, ny = (size_t)ctx->tg.ny;
  if (ctx->shard_j0 >= 1 && ctx->shard_j1 >= ctx->shard_j0 && (size_t)ctx->shard_j1 <= ny) {
    *k0 = (size_t)(ctx->shard_j0 - 1) * nx;
    *cnt = (size_t)(ctx->shard_j1 - ctx->shard_j0 + 1) * nx;
  } else { *k0 = 0; *cnt = nx * ny; }
}

// rows the mapping was built for = the shard active at that moment
static inline void map_rows_from_shard(sosie_ctx* ctx) { ctx->map_j0 = ctx->shard_j0; ctx->map_j1 = ctx->shard_j1; }
static inline int shard_check(sosie_ctx* ctx, const char* who) {
  if (ctx->shard_j0 >= 1 && ctx->tg.ny > 0 && ctx->shard_j1 > ctx->tg.ny) {
    ctx->err = std::string(who) + ": the shard set by sosie_bilin_set_shard exceeds the target rows";
    return SOSIE_ERR_ARG;
  }
  return SOSIE_OK;
}

// grid size for grid-stride kernels: a multiple of the SM count (148 on B200)
static inline int grid_for(const sosie_ctx* ctx, int64_t work_items, int block, int max_blocks_per_sm = 8) {
  int64_t need = (work_items + block - 1) / block;
  int64_t cap = (int64_t)ctx->num_sms * max_blocks_per_sm;
  if (need < 1) need = 1;
  if (need > cap) need = cap;
  return (int)need;
}

// result of the FIND_NEAREST_POINT prelude
struct MapPlan {
  int j_strt = 1, j_stop = 0, icr = 1, jlo = 1, jhi = 0;
  bool reg_s = true, reg_t = true;
  double y_min_s = 0, y_max_s = 0, y_min_t = 0, y_max_t = 0;
};

// queue.cu
void queue_destroy(sosie_ctx* ctx);
int queue_flush(sosie_ctx* ctx);          // submit the open batch, wait for every queued result
bool queue_pending(const sosie_ctx* ctx);
void pin_registry_add(const void* p, size_t bytes);   // page-locked host ranges the library knows about
void pin_registry_remove(const void* p);

#define SOSIE_MAILBOX_BYTES (512 * 1024)
// device -> host read-back of a small object through the mailbox (synchronises the context stream)
int fetch_small(sosie_ctx* ctx, const void* dsrc, void* hdst, size_t bytes);
int fetch_pair(sosie_ctx* ctx, const void* d1, void* h1, size_t b1, const void* d2, void* h2, size_t b2);

// entry points implemented in the other translation units
int bilin_materialize_mask(sosie_ctx* ctx);   // fills t_mask with 1 if it is still virtual
int bilin_map_impl(sosie_ctx* ctx, bool async = false);
int bilin_map_check(sosie_ctx* ctx);          // delivers the deferred status of sosie_bilin_map_async (one read-back)
void p2p_destroy(sosie_ctx* ctx);            // capi.cu: frees the peer-memory mailbox, closes the IPC mappings
int bilin_map_alloc(sosie_ctx* ctx);
int bilin_map_prelude(sosie_ctx* ctx, MapPlan* pl, int xrow_a, int xrow_b, bool* need_full_x);
// rows ja..jb (1-based) reduced here, the rest obtained from the other ranks; *undecided: fall back to the full prelude
int bilin_map_prelude_exchange(sosie_ctx* ctx, MapPlan* pl, int ja, int jb, bool* undecided);
int bilin_easy_prepare(sosie_ctx* ctx);
int bilin_easy_rows(sosie_ctx* ctx, int jlo, int jhi, size_t k0, size_t kcnt, int* derr);
int bilin_pack_range(sosie_ctx* ctx, size_t k0, size_t kcnt);
int bilin_src_rowrange(sosie_ctx* ctx, size_t k0, size_t kcnt, int* row_lo, int* row_hi);
int bilin_map_rows_d2h(sosie_ctx* ctx, cudaStream_t st, int j0, int j1, int32_t* metrics, double* alphabeta,
                       int16_t* id_problem, float* dist_np, int32_t* mask_out);
int bilin_first_pipelined(sosie_ctx* ctx, int k_ew, int nx1, int ny1, const double* X10, const double* Y10, const float* Z1,
                          int nx2, int ny2, const double* X20, const double* Y20, float* Z2, const int32_t* mask,
                          int l_reg_src, int i_orca_trg, int nslab, int32_t* metrics, double* alphabeta, int16_t* id_problem,
                          float* dist_np, bool* handled);
int bilin_pack_impl(sosie_ctx* ctx);
int bilin_apply_impl(sosie_ctx* ctx, int nslab, const float* dZ1, float* dZ2, int first_call, int use_clamp,
                     float vmin, float vmax, const int32_t* dmask_trg, float rmiss);
int bilin_apply_uv_impl(sosie_ctx* ctx, int nslab, const float* dU1, const float* dV1, float* dUo, float* dVo,
                        const double* dcos, const double* dsin);
int bilin_prepare_src(sosie_ctx* ctx);
int selftest_div_impl(sosie_ctx* ctx, unsigned long long n, unsigned long long seed, unsigned long long* mismatches);
int track_interp_impl(sosie_ctx* ctx, int ni, int nj, int nk, const float* F1, const float* F2, double vt1, double vt2, const int8_t* imask,
                      long long n, const double* rt, const int32_t* jiP, const int32_t* jjP, const int32_t* iqd, const double* xa,
                      const double* xb, const double* r_obs, double obs_lo, double obs_hi, int clip_missing, double rmissval, double* f_bl,
                      double* f_np, int8_t* fmask);
int interp_bl_impl(sosie_ctx* ctx, int k_ew, int nxi, int nyi, const float* dZ, int n, const int32_t* ji,
                   const int32_t* jj, const int32_t* iq, const double* xa, const double* xb, float* out);
int akima_prepare_impl(sosie_ctx* ctx);
int akima_apply_impl(sosie_ctx* ctx, int nslab, const float* dZ1, float* dZ2);
int bdrown_impl(sosie_ctx* ctx, int k_ew, int ni, int nj, int nslab, float* dX, const int32_t* dmask,
                int mask_per_slab, int nb_inc, int nb_smooth, float pmin, float pmax, float pval_land,
                int32_t* nsweeps_host);
int smoother_impl(sosie_ctx* ctx, int k_ew, int ni, int nj, int nslab, float* dX, int nb_smooth,
                  const int32_t* dmsk, int l_emp);
int drown_impl(sosie_ctx* ctx, int k_ew, int ni, int nj, int nslab, float* dX, const int8_t* dmask, int nb_inc,
               int32_t* nsweeps_host);
int akima_1d_3d_impl(sosie_ctx* ctx, int nx, int ny, int nk1, const float* vz1_host, const float* dxf1, int nk2, const float* vz2_host,
                     float* dxf2, float rfill_val);
int akima_1d_1d_impl(sosie_ctx* ctx, long long ncol, int nk1, const double* vz1, long long z1cs, long long z1ks, const double* vf1, long long f1cs,
                     long long f1ks, int nk2, const double* vz2, long long z2cs, long long z2ks, float* vf2, long long f2cs, long long f2ks,
                     int has_rmask, double rmask_val, int l_sp, int l_bp, int32_t* dstatus);
int interp3d_sigma_impl(sosie_ctx* ctx, int ni, int nj, int nk_src, const float* ddepth_src, const float* ddata3d_tmp, int nk_trg,
                        const float* ddepth_trg, const int32_t* dmask_trg, int lmout, int src_sigma, int trg_sigma, float* ddata3d_trg);
int lbc_lnk_impl(sosie_ctx* ctx, int nperio, int jpi, int jpj, int nslab, float* dX, int cd_type, double psgn, int has_pval, double pval);
int interp3d_post_impl(sosie_ctx* ctx, int ni, int nj, int nk, float* dX, const int32_t* dmask_trg, int ixtrpl_bot, float vmin, float vmax,
                       float rmiss_val, int i_orca_trg, int c_orca_trg, int lmout);
int extrp_hl_impl(sosie_ctx* ctx, int ni, int nj, int nslab, float* dX, int jj_ex_top, int jj_ex_btm, int nlat_icr_trg);
int angle2_impl(sosie_ctx* ctx, int nperio, int nx, int ny, const double* din, double* dout);
int rotate_impl(sosie_ctx* ctx, int64_t n, int nslab, const float* dU, const float* dV, const double* dcos,
                const double* dsin, float* dUo, float* dVo, int inverse);
