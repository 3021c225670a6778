/* sosie_b200.h -- C ABI of the B200-native replacement for SOSIE's horizontal-interpolation hot path.
 *
 * This is the drop-in boundary: exactly what the ISO_C_BINDING shim modules in
 * sosie_b200/fortran/ (mod_bilin_2d, mod_akima_2d, mod_bdrown, mod_drown) bind to.  Each entry
 * cites the reference interface it replaces (paths relative to the reference tree).
 *
 * Conventions (same as the reference):
 *   - every array is Fortran column-major, extents passed explicitly as int32;
 *   - coordinates REAL(8), fields REAL(4), masks INTEGER(4) (INTEGER(1) for mod_drown);
 *   - k_ew = -1: no east-west periodicity, >= 0: periodic with k_ew overlap points;
 *   - longitudes in [0,360), NaN already replaced by -9999;
 *   - return value: 0 = ok, non-zero = the condition on which the reference would PRINT+STOP
 *     (sosie_last_error() gives the text; the Fortran shim prints it and STOPs);
 *   - pointers are HOST pointers unless the entry name ends in _dev (device pointers, used by
 *     callers that keep slabs resident in HBM; all _dev work is enqueued on the context stream);
 *   - one host thread per context at a time (the reference is not re-entrant either).
 *
 * There is no CPU fallback: every entry fails with SOSIE_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef SOSIE_B200_H
#define SOSIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sosie_ctx sosie_ctx;

enum {
  SOSIE_OK = 0,
  SOSIE_ERR_CUDA = 1,      /* CUDA runtime / no device                                            */
  SOSIE_ERR_ARG = 2,       /* bad argument (shape mismatch: the reference's "do not agree" STOPs)  */
  SOSIE_ERR_STATE = 3,     /* call order (apply before map/load)                                   */
  SOSIE_ERR_REF_STOP = 4   /* a data condition on which the reference STOPs (see last_error)       */
};

/* ---- context ------------------------------------------------------------------------------ */
int sosie_create(sosie_ctx** ctx, int device);
int sosie_destroy(sosie_ctx* ctx);
const char* sosie_last_error(const sosie_ctx* ctx);
/* run everything on the caller's CUDA stream (a cudaStream_t passed as void*); NULL = own stream */
int sosie_set_stream(sosie_ctx* ctx, void* cuda_stream);
int sosie_sync(sosie_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int64_t sosie_launch_count(const sosie_ctx* ctx);
const char* sosie_version(void);
/* per-kernel device timing: when enabled, CUDA events bracket the dominant kernels on the context
 * stream; sosie_get_timing returns the duration (ms) of the LAST bracket of that id (syncs on it). */
enum { SOSIE_T_MAP = 0, SOSIE_T_PACK = 1, SOSIE_T_APPLY = 2, SOSIE_T_BDROWN = 3, SOSIE_T_SMOOTH = 4,
       SOSIE_T_AKIMA_PREP = 5, SOSIE_T_AKIMA_EVAL = 6, SOSIE_T_DROWN = 7 };
int sosie_profile(sosie_ctx* ctx, int32_t enable);
int sosie_get_timing(sosie_ctx* ctx, int32_t id, float* ms);
/* multi-GPU sharding of the bilinear path by target rows: this context maps / packs /
 * applies only target rows j0..j1 (1-based, inclusive; 0,0 = all rows).  The search prelude
 * (row range, regularity tests) still sees the whole target grid, so results equal the unsharded ones.
 * Only the regular-source (EASY) search shards; the curvilinear-source chain is not split.           */
int sosie_bilin_set_shard(sosie_ctx* ctx, int32_t j0, int32_t j1);

/* Row-sharded jobs, end to end from host arrays (sosie_bilin_2d_first on a context with sosie_bilin_set_shard): the
 * search prelude FIND_NEAREST_POINT (src/mod_manip.f90:888-947) reduces the WHOLE target latitude array (row range covered
 * by the source, monotony, regularity), which would make every rank upload every row.  With an all-gather callback the
 * ranks reduce their own rows and exchange one small record each (some scalars + 24 bytes per target column): the only
 * collective of the path, supplied by the host program (MPI_Allgather, torch.distributed, ...).  fn(user, send, recv,
 * bytes) must gather `bytes` bytes from every rank into recv (nranks * bytes, any rank order), return 0, and is called at
 * most once per first call, by all ranks or by none.  The ranks' shards must tile the target rows.  The combined prelude is
 * the sequential one bit for bit (min / max / first occurrence only); if no rank can prove the target grid irregular the
 * library falls back to the whole-array prelude.  fn = NULL switches the exchange off.
 * sosie_exchange_used: 1 when the last first call took the exchange path.                                          */
typedef int (*sosie_allgather_fn)(void* user, const void* send, void* recv, size_t bytes_per_rank);
int sosie_set_exchange(sosie_ctx* ctx, int32_t nranks, sosie_allgather_fn fn, void* user);
int sosie_exchange_used(sosie_ctx* ctx);   /* 0 no exchange, 1 callback (host or device), 2 peer memory */

/* ---- bilinear: grids -> mapping -> apply -------------------------------------------------------
 * Replaces BILIN_2D / MAPPING_BL / INTERP_BL, src/mod_bilin_2d.f90:44-361,366-691, and the search
 * FIND_NEAREST_POINT/_EASY/_TWISTED, src/mod_manip.f90:834-1440.
 *
 * sosie_bilin_set_grids: the coordinate part of BILIN_2D's prologue (:97-171): 1-D axes are kept
 *   1-D on device (src_1d/trg_1d != 0 -> X is (nx), Y is (ny)), else (nx,ny) arrays; when
 *   l_reg_src and the grid reaches the pole the two extra rows of EXT_NORTH_TO_90_REGG
 *   (src/mod_manip.f90:1751-1839) are added.  mask_domain_trg may be NULL (all 1).
 *   l_reg_src and i_orca_trg are the mod_conf globals the reference reads implicitly.            */
int sosie_bilin_set_grids(sosie_ctx* ctx, int32_t k_ew_per,
                          int32_t nx1, int32_t ny1, const double* X10, const double* Y10, int32_t src_1d,
                          int32_t l_reg_src,
                          int32_t nx2, int32_t ny2, const double* X20, const double* Y20, int32_t trg_1d,
                          const int32_t* mask_domain_trg, int32_t i_orca_trg);
/* same with device-resident coordinates (borrowed until the next set_grids / destroy) */
int sosie_bilin_set_grids_dev(sosie_ctx* ctx, int32_t k_ew_per,
                              int32_t nx1, int32_t ny1, const double* dX10, const double* dY10, int32_t src_1d,
                              int32_t l_reg_src,
                              int32_t nx2, int32_t ny2, const double* dX20, const double* dY20, int32_t trg_1d,
                              const int32_t* dmask_domain_trg, int32_t i_orca_trg);

/* A device-resident caller that also holds the (small) 1-D source axes on the host declares them here: the following
 * sosie_bilin_set_grids_dev calls with src_1d != 0 and the same extents then take the host-side decisions of BILIN_2D's
 * prologue (pole extension :113-148, sortedness, uniform-axis hints) from this copy instead of reading the device arrays back
 * -- with sosie_bilin_map_async the whole grids -> mapping -> apply sequence is then free of host round trips.  The copy
 * MUST equal the device arrays (SOSIE_VERIFY_HINTS=1 in the environment checks it on every call); NULL pointers clear it.  */
int sosie_bilin_src_axes_hint(sosie_ctx* ctx, int32_t nx1, int32_t ny1, const double* X10_host, const double* Y10_host);

/* FIND_NEAREST_POINT, src/mod_manip.f90:834-963, as the public procedure ij_from_lon_lat.f90:290 calls: for every target
 * point the (i,j) of the nearest source point, -1 where the target was not searched (masked, outside the source's
 * latitude range) or not found.  All four coordinate arrays are 2-D, (nx,ny) column-major, as the assumed-shape
 * dummies of the reference; the search picks EASY or TWISTED itself (L_IS_GRID_REGULAR on the source, :1632).
 * mask_domain_trg (INTENT(inout), OPTIONAL in the reference): NULL = absent; on return it holds mask_ignore_t (:958).
 * Leaves no mapping in the context.                                                                                 */
int sosie_find_nearest_point(sosie_ctx* ctx, int32_t nx_t, int32_t ny_t, const double* Xtrg, const double* Ytrg,
                             int32_t nx_s, int32_t ny_s, const double* Xsrc, const double* Ysrc,
                             int32_t* JIp, int32_t* JJp, int32_t* mask_domain_trg);

/* MAPPING_BL (:366-691): builds, on device, metrics(nx2,ny2,3)=(iP,jP,iqdrn), alphabeta(nx2,ny2,2),
 * ID_problem(nx2,ny2) and distance_to_np(nx2,ny2), and keeps them cached in the context.          */
int sosie_bilin_map(sosie_ctx* ctx);
/* the same for device-resident callers, without any host round trip when the source is regular and given as 1-D axes (the
 * EASY search): the prelude of FIND_NEAREST_POINT (:888-947) is reduced AND decided on the device, the search kernel reads
 * its row range from there.  Returns after enqueueing; what sosie_bilin_map would have reported (rows not covered by the
 * source, L_InPoly's STOP, nearest point not found) is returned by the next synchronising entry of the context: sosie_sync,
 * sosie_bilin_map_info, sosie_bilin_get_map, the host-pointer applies.  Other source kinds behave like sosie_bilin_map.  */
int sosie_bilin_map_async(sosie_ctx* ctx);
/* device-side prelude exchange of row-sharded mappings on resident grids (sosie_bilin_map[_async] on a context with
 * sosie_bilin_set_shard): each rank reduces only ITS target rows and the ranks all-gather one record each (32 bytes + 24 per
 * target column) GPU -> GPU.  fn(user, dsend, drecv, bytes, cuda_stream) must ENQUEUE on cuda_stream an all-gather of `bytes`
 * bytes from every rank into drecv (nranks * bytes, any rank order; ncclAllGather, torch.distributed.all_gather_into_tensor)
 * and return 0; it is called once per mapping, by every rank.  The merged prelude equals the one of a pass over the whole
 * array (min / max / first occurrence in row order).  fn = NULL switches it off (every rank then reduces the whole array). */
typedef int (*sosie_allgather_dev_fn)(void* user, const void* dsend, void* drecv, size_t bytes_per_rank, void* cuda_stream);
int sosie_set_exchange_dev(sosie_ctx* ctx, int32_t nranks, sosie_allgather_dev_fn fn, void* user);
/* The same exchange WITHOUT a collective call: one producer kernel and one consumer kernel over peer memory (NVLink P2P
 * stores).  sosie_p2p_init allocates this rank's mailbox (slot_bytes per rank and parity; 0 = 2 MB, enough for 87 000 target
 * columns) and returns its 64-byte CUDA IPC handle (ipc_handle_out, may be NULL) and its device address (mailbox_out, may be
 * NULL).  The ranks swap the handles any way they like (MPI_Allgather, torch.distributed) and call sosie_p2p_connect with
 * the nranks * 64 bytes -- or, for contexts of ONE process, with the nranks mailbox addresses.  From then on a sharded
 * sosie_bilin_map[_async] on resident grids writes its prelude record straight into every rank's mailbox and merges the
 * records once all have arrived; it takes precedence over sosie_set_exchange_dev.  Every rank must run the same number of
 * mappings; a record that does not arrive within ~10 s (SOSIE_P2P_TIMEOUT_MS, read by sosie_p2p_init) is reported by the
 * next synchronising entry (SOSIE_ERR_STATE).
 * Where CUDA IPC is not available the calls fail with SOSIE_ERR_CUDA and the callback exchange remains.                   */
int sosie_p2p_init(sosie_ctx* ctx, int32_t nranks, int32_t rank, size_t slot_bytes, void* ipc_handle_out, void** mailbox_out);
int sosie_p2p_connect(sosie_ctx* ctx, const void* ipc_handles, void* const* mailboxes);
/* copy the cached mapping to host arrays (what P2D_MAPPING_AB, src/io_ezcdf.f90:2195, writes).
 * Any pointer may be NULL.  mask_out receives mask_domain_trg as modified by the search (-1/-2).
 * A sharded context (sosie_bilin_set_shard) copies only its own target rows; the others are left untouched. */
int sosie_bilin_get_map(sosie_ctx* ctx, int32_t* metrics, double* alphabeta, int16_t* id_problem,
                        float* dist_np, int32_t* mask_out);
/* load a mapping read by RD_MAPPING_AB (src/io_ezcdf.f90:2280) instead of building it; may be called
 * before or after sosie_bilin_set_grids (a set_grids with the same target extents keeps it).        */
int sosie_bilin_load_map(sosie_ctx* ctx, int32_t nx2, int32_t ny2, const int32_t* metrics,
                         const double* alphabeta, const int16_t* id_problem);
/* device-pointer variants (enqueued on the context stream): the multi-GPU flow builds the mapping on one GPU, broadcasts the
 * arrays GPU -> GPU (NCCL over NVLink) and loads them on the others without touching the host              */
int sosie_bilin_get_map_dev(sosie_ctx* ctx, int32_t* dmetrics, double* dalphabeta, int16_t* did_problem, float* ddist_np);
int sosie_bilin_load_map_dev(sosie_ctx* ctx, int32_t nx2, int32_t ny2, const int32_t* dmetrics, const double* dalphabeta,
                             const int16_t* did_problem);
/* search diagnostics: info[0..7] = j_strt_t, j_stop_t, jlat_icr, l_is_reg_s, l_is_reg_t,
 * algorithm (0 EASY, 1 TWISTED), TWISTED chain iterations, l_add_extra_j                          */
int sosie_bilin_map_info(sosie_ctx* ctx, int32_t* info);

/* the apply loop of BILIN_2D (:209-263) for nslab source slabs Z1(nx1,ny1,nslab) ->
 * Z2(nx2,ny2,nslab): identical-point copy, INTERP_BL with its error codes, pole rows of
 * EXT_NORTH_TO_90_REGG recomputed per slab, ORCA last-row patch.  first_call: value of
 * l_first_call_interp_routine for the first slab (evaluates l_last_y_row_missing, :241-247).      */
int sosie_bilin_apply(sosie_ctx* ctx, int32_t nslab, const float* Z1, float* Z2, int32_t first_call);
int sosie_bilin_apply_dev(sosie_ctx* ctx, int32_t nslab, const float* dZ1, float* dZ2, int32_t first_call);

/* apply with the cheap per-record post-ops of INTERP_2D fused (src/mod_interp.f90:97-98,133-137):
 * clamp to [vmin,vmax] when use_clamp, then rmiss_val where mask_trg == 0 when dmask_trg != NULL.  */
int sosie_bilin_apply_ex_dev(sosie_ctx* ctx, int32_t nslab, const float* dZ1, float* dZ2, int32_t first_call,
                             int32_t use_clamp, float vmin, float vmax, const int32_t* dmask_trg, float rmiss_val);

/* apply to a (U,V) pair and rotate on the target grid in the same kernel (corr_vect,
 * src/corr_vect.f90:622-624,630,632): Uo = REAL(cos*U2 + sin*V2,4), Vo = REAL(cos*V2 - sin*U2,4)
 * with U2,V2 the interpolated slabs; xcos/xsin REAL(8)(nx2,ny2).                                   */
int sosie_bilin_apply_uv_dev(sosie_ctx* ctx, int32_t nslab, const float* dU1, const float* dV1, float* dUo,
                             float* dVo, const double* dxcos, const double* dxsin);

/* one-call mirror of BILIN_2D(k_ew_per, X10, Y10, Z1, X20, Y20, Z2, cnpat, mask_domain_trg)
 * (src/mod_bilin_2d.f90:44): on *l_first_call != 0 sets the grids and builds the mapping (or uses the
 * one given through sosie_bilin_load_map when have_mapping != 0), applies it, and clears
 * *l_first_call exactly as :267 does.  Host pointers.                                              */
int sosie_bilin_2d(sosie_ctx* ctx, int32_t k_ew_per, int32_t nx1, int32_t ny1, const double* X10,
                   const double* Y10, int32_t src_1d, const float* Z1, int32_t nx2, int32_t ny2,
                   const double* X20, const double* Y20, int32_t trg_1d, float* Z2,
                   const int32_t* mask_domain_trg, int32_t l_reg_src, int32_t i_orca_trg,
                   int32_t* l_first_call, int32_t have_mapping, int32_t nslab);

/* the FIRST call of BILIN_2D with MAPPING_BL's products handed back in the same call -- what the reference's first
 * call does end to end: MAPPING_BL (:196) -> P2D_MAPPING_AB writes metrics / alphabeta / iproblem / dist_np
 * (src/mod_bilin_2d.f90:686, src/io_ezcdf.f90:2195) -> apply loop (:209).  metrics(nx2,ny2,3), alphabeta(nx2,ny2,2),
 * id_problem(nx2,ny2), dist_np(nx2,ny2): any may be NULL.  With a regular source given as 1-D axes and a 2-D target
 * the call is a three-stream pipeline (coordinate upload | search+weights per block of target rows | mapping download,
 * then only the source rows the mapping reads are uploaded and applied): see csrc/bilin_pipeline.cu.  Host arrays
 * should be page-locked (sosie_host_register) for the copies to overlap.  Equivalent to
 * sosie_bilin_2d(first call) + sosie_bilin_get_map; a sharded context fills only its target rows.                 */
int sosie_bilin_2d_first(sosie_ctx* ctx, int32_t k_ew_per, int32_t nx1, int32_t ny1, const double* X10,
                         const double* Y10, int32_t src_1d, const float* Z1, int32_t nx2, int32_t ny2,
                         const double* X20, const double* Y20, int32_t trg_1d, float* Z2,
                         const int32_t* mask_domain_trg, int32_t l_reg_src, int32_t i_orca_trg, int32_t nslab,
                         int32_t* metrics, double* alphabeta, int16_t* id_problem, float* dist_np);
/* test hooks of the batched apply (results are identical whichever kernels run).  A batched apply splits the 32x8 target
 * tiles between up to three kernels by the shape of a tile's source box (csrc/bilin_apply.cu): k_apply_grp (groups of 8 or 4
 * slabs staged with 16-byte cp.async, one block barrier per group: the default for boxes up to 768 padded cells),
 * k_apply_tma (3-D tensor map of the slab stack, one cp.async.bulk.tensor per tile and group; OFF by default -- switch on
 * with sosie_apply_tma(1) or SOSIE_APPLY_TMA=1 -- because that instruction traps under the gVisor nvproxy driver of the
 * GPU pool this library was developed on, see profiles/r02_tma_nvproxy.md) and k_apply_staged (4-byte cp.async per slab,
 * direct gathers for boxes that fit nowhere; any alignment).  sosie_apply_tma(on): bit 0 TMA kernel on, bit 2 group
 * kernel off.  sosie_apply_tile_classes: counts[0..2] = tiles that fit the TMA box, the group kernel, neither (-1 before the
 * first batched apply of a mapping).  sosie_selftest_div: the reciprocal-based correctly rounded division of the apply
 * kernels against the IEEE division on n pseudo-random operand pairs; *mismatches must come back 0.                     */
int sosie_apply_tma(sosie_ctx* ctx, int32_t on);
int sosie_apply_tile_classes(sosie_ctx* ctx, int32_t* counts);
int sosie_selftest_div(sosie_ctx* ctx, uint64_t n, uint64_t seed, uint64_t* mismatches);
/* rows[0..1] = first / last physical source row (0-based) read by this context's packed mapping records: the
 * host-pointer apply paths upload only those rows of every slab (a regional target touches part of a global slab) */
int sosie_bilin_src_rows(sosie_ctx* ctx, int32_t* rows);
/* first calls with fewer targets than n take the sequential path (default 2^20; tests set 0 to force the pipeline) */
int sosie_set_pipeline_min_targets(sosie_ctx* ctx, int64_t n);
/* page-lock / release a caller-owned host array (cudaHostRegister): the Fortran shim pins the arrays it passes
 * repeatedly (coordinates, record buffers) so that the library's copies run asynchronously at full PCIe speed   */
int sosie_host_register(void* p, size_t bytes);
int sosie_host_unregister(void* p);

/* ---- asynchronous per-record / per-level calls: enqueue ... flush (csrc/queue.cu) ------------------------------------
 * The reference calls INTERP_2D once per time record (src/sosie.f90:133-165) and BDROWN, then bilin_2d / akima_2d, once per
 * level (src/mod_interp.f90:203-265, 288-337): one slab per call.  These entries take ONE slab like those call sites do,
 * but only queue it: the slab is uploaded on a copy stream at once, up to `depth` (default 32) queued slabs of the same
 * operation are coalesced into one batched launch, results are downloaded on a third stream while the next batch uploads.
 * The output arrays are complete after sosie_flush -- or after any synchronous (non-_dev, non-_enqueue) entry of the same
 * context, which flushes first.  Contract: a page-locked input (sosie_host_alloc, sosie_host_register) is BORROWED until
 * the flush and read by DMA in place; a pageable input is copied before the call returns and may be reused at once (the
 * reference loop re-reads data_src into the same array).  Outputs must stay valid until the flush; pageable outputs are
 * filled by a host copy at flush time, page-locked ones by DMA.  Errors of a queued operation are reported by the enqueue
 * or flush call that submits its batch.
 * sosie_bilin_apply_enqueue: the apply loop of BILIN_2D (:209-263) for one slab; the grids must be set and the mapping
 *   built or loaded (the first BILIN_2D call of a job stays synchronous: it builds the mapping and evaluates
 *   l_last_y_row_missing).  A sharded context fills only its target rows.
 * sosie_akima_apply_enqueue: AKIMA_2D for one slab after sosie_akima_set_grids.
 * sosie_bdrown_enqueue: BDROWN(k_ew, X, mask, nb_inc, nb_smooth, pmin, pmax, pval_land) on one slab, X(ni,nj) in place,
 *   mask(ni,nj) INTEGER(4) of that slab; nsweeps (nullable) receives the sweeps executed, at flush.
 * sosie_queue_stats: stats[0..3] = batched launches, slabs queued, inputs staged through the pinned ring (pageable),
 *   outputs staged (pageable) since the context was created.                                                        */
int sosie_bilin_apply_enqueue(sosie_ctx* ctx, const float* Z1, float* Z2);
int sosie_akima_apply_enqueue(sosie_ctx* ctx, const float* Z1, float* Z2);
int sosie_bdrown_enqueue(sosie_ctx* ctx, int32_t k_ew, int32_t ni, int32_t nj, float* X, const int32_t* mask, int32_t nb_inc,
                         int32_t nb_smooth, float pmin, float pmax, float pval_land, int32_t* nsweeps);
int sosie_flush(sosie_ctx* ctx);
int sosie_set_queue_depth(sosie_ctx* ctx, int32_t depth);
int sosie_queue_stats(sosie_ctx* ctx, int64_t* stats);
/* page-locked host memory for the record / level buffers of an enqueue loop (cudaHostAlloc; NULL on failure) */
void* sosie_host_alloc(size_t bytes);
int sosie_host_free(void* p);

/* scalar INTERP_BL(k_ew_per, jiP, jjP, iqd, xa, xb, Z_in) for n trajectory points
 * (src/mod_bilin_2d.f90:275; callers src/interp_to_ground_track.f90:750): host pointers.           */
int sosie_interp_bl(sosie_ctx* ctx, int32_t k_ew_per, int32_t nxi, int32_t nyi, const float* Z_in, int32_t n,
                    const int32_t* jiP, const int32_t* jjP, const int32_t* iqd, const double* xa,
                    const double* xb, float* out);

/* The per-point block of the trajectory front ends for n track points that fall between the same two
 * model records: src/interp_to_ground_track.f90:718-779 (nk = 1; F1, F2 = xvar1, xvar2 (ni,nj) REAL(4) at times vt1 < vt2;
 * rt(n) the track times; the reference rebuilds the whole time-interpolated field per point, here only the elements a
 * point reads are formed, with the same roundings) and src/interp_to_hydro_section.f90:567-607 (nk levels of one record:
 * F2 = NULL, rt unused).  imask(ni,nj,nk) INTEGER(1).  jiP, jjP, iqd, xa, xb: IMETRICS(1,:,1..3), RAB(1,:,1..2) of the
 * points (from sosie_bilin_map on the (1,N) target).  A point is valid when iP, jP > 0, imask(iP,jP,k) == 1, obs_lo < r_obs <
 * obs_hi (:738, -20/20; hydro: -20/50) and its 8 neighbours (indices clamped to the grid) are sea (:741-745).  For valid
 * points only: f_np(k,p) = nearest-point value, f_bl(k,p) = REAL(INTERP_BL(-1, ...),8) (set to rmissval when clip_missing
 * and < -9990, :787), fmask(k,p) = 1; the others keep the caller's pre-fill, as in the reference.  Level index fastest. */
int sosie_track_interp(sosie_ctx* ctx, int32_t ni, int32_t nj, int32_t nk, const float* F1, const float* F2, double vt1,
                       double vt2, const int8_t* imask, int64_t n, const double* rt, const int32_t* jiP,
                       const int32_t* jjP, const int32_t* iqd, const double* xa, const double* xb, const double* r_obs,
                       double obs_lo, double obs_hi, int32_t clip_missing, double rmissval, double* f_bl, double* f_np,
                       int8_t* fmask);
int sosie_track_interp_dev(sosie_ctx* ctx, int32_t ni, int32_t nj, int32_t nk, const float* dF1, const float* dF2,
                           double vt1, double vt2, const int8_t* dimask, int64_t n, const double* drt,
                           const int32_t* djiP, const int32_t* djjP, const int32_t* diqd, const double* dxa,
                           const double* dxb, const double* dr_obs, double obs_lo, double obs_hi, int32_t clip_missing,
                           double rmissval, double* df_bl, double* df_np, int8_t* dfmask);

/* ---- Akima 2D ------------------------------------------------------------------------------------
 * Replaces AKIMA_2D(k_ew_per, X10, Y10, Z1, X20, Y20, Z2, icall), src/mod_akima_2d.f90:59-239
 * (FILL_EXTRA_BANDS src/mod_manip.f90:69, SLOPES_AKIMA :483, build_pol :249, pol_val :445,
 * find_nearest_akima :670).  i_orca_src is the mod_conf global read at :168.                       */
int sosie_akima_set_grids(sosie_ctx* ctx, int32_t k_ew_per, int32_t nx1, int32_t ny1, const double* X10,
                          const double* Y10, int32_t src_1d, int32_t i_orca_src, int32_t nx2, int32_t ny2,
                          const double* X20, const double* Y20, int32_t trg_1d);
int sosie_akima_apply(sosie_ctx* ctx, int32_t nslab, const float* Z1, float* Z2);
int sosie_akima_apply_dev(sosie_ctx* ctx, int32_t nslab, const float* dZ1, float* dZ2);
/* test hook (identical results): 0 = default (slope arrays, shared-memory tiled slopes kernel, then the evaluation kernel;
 * a regular source reads its coordinates from 1-D axes), 1 = per-point slopes kernel, 3 = a regular source runs the fused
 * kernel (slopes of the source box under a 32x8 tile of targets built in shared memory, then the evaluation: no slope arrays
 * in HBM -- measured slower on B200, the stage is bound by its REAL(8) divisions: DESIGN.md) */
int sosie_akima_untiled_slopes(sosie_ctx* ctx, int32_t on);
/* the SAVE'd ixy_pos(nx2,ny2,2) table (src/mod_akima_2d.f90:41) */
int sosie_akima_get_ixy(sosie_ctx* ctx, int32_t* ixy_pos);

/* ---- BDROWN / SMOOTHER (src/mod_bdrown.f90:18-439, 444-608) --------------------------------------
 * X(ni,nj,nslab) in place; mask(ni,nj) shared by all slabs, or (ni,nj,nslab) when mask_per_slab.
 * nsweeps (nullable) receives the number of incursion sweeps actually executed per slab
 * (the device-side "any land left" flag replacing ANY(maskv==0), :125).                           */
int sosie_bdrown(sosie_ctx* ctx, int32_t k_ew, int32_t ni, int32_t nj, int32_t nslab, float* X,
                 const int32_t* mask, int32_t mask_per_slab, int32_t nb_inc, int32_t nb_smooth, float pmin,
                 float pmax, float pval_land, int32_t* nsweeps);
int sosie_bdrown_dev(sosie_ctx* ctx, int32_t k_ew, int32_t ni, int32_t nj, int32_t nslab, float* dX,
                     const int32_t* dmask, int32_t mask_per_slab, int32_t nb_inc, int32_t nb_smooth,
                     float pmin, float pmax, float pval_land, int32_t* nsweeps_host);
/* SMOOTHER(k_ew, X, nb_smooth, msk, l_exclude_mask_points): msk may be NULL; l_emp mirrors
 * PRESENT(l_exclude_mask_points) (presence, not value: :496).                                      */
/* test hook: which BDROWN implementation runs (results are identical): 0 = automatic (one-CTA-per-slab on-chip kernel
 * when a slab fits the shared memory of one SM, level lists otherwise), 1 = level-list general path for every size,
 * 2 = legacy dense path (one whole-array kernel per sweep).                                                        */
int sosie_bdrown_force_general(sosie_ctx* ctx, int32_t on);
int sosie_smoother(sosie_ctx* ctx, int32_t k_ew, int32_t ni, int32_t nj, int32_t nslab, float* X,
                   int32_t nb_smooth, const int32_t* msk, int32_t l_emp);

/* ---- DROWN (Gaussian), src/mod_drown.f90:18-198: mask is INTEGER(1) ------------------------------ */
int sosie_drown(sosie_ctx* ctx, int32_t k_ew, int32_t ni, int32_t nj, int32_t nslab, float* X,
                const int8_t* mask, int32_t nb_inc, int32_t* nsweeps);

/* ---- vector rotation alone, src/corr_vect.f90:622-624 (inverse != 0: :963,970) ------------------- */
int sosie_rotate_uv(sosie_ctx* ctx, int64_t n, int32_t nslab, const float* U, const float* V,
                    const double* xcos, const double* xsin, float* Uo, float* Vo, int32_t inverse);
int sosie_rotate_uv_dev(sosie_ctx* ctx, int64_t n, int32_t nslab, const float* dU, const float* dV,
                        const double* dxcos, const double* dxsin, float* dUo, float* dVo, int32_t inverse);

/* ---- vertical stage of INTERP_3D: the cube stays in HBM between the horizontal and the vertical
 * interpolation ----------------------------------------------------------------------------------------------------
 * AKIMA_1D_3D(vz1, xf1, vz2, xf2, rfill_val), src/mod_akima_1d.f90:243-389 (z -> z levels, REAL(4)): xf1(nx,ny,nk1) ->
 * xf2(nx,ny,nk2); vz1(nk1), vz2(nk2) are HOST arrays in both variants (the level search is done on the host once).   */
int sosie_akima_1d_3d(sosie_ctx* ctx, int32_t nx, int32_t ny, int32_t nk1, const float* vz1, const float* xf1, int32_t nk2,
                      const float* vz2, float* xf2, float rfill_val);
int sosie_akima_1d_3d_dev(sosie_ctx* ctx, int32_t nx, int32_t ny, int32_t nk1, const float* vz1, const float* dxf1,
                          int32_t nk2, const float* vz2, float* dxf2, float rfill_val);
/* AKIMA_1D_1D(vz1, vf1, vz2, vf2 [, rmask_val, l_surf_persist, l_botm_persist]), src/mod_akima_1d.f90:26-232 (the
 * per-column variant: REAL(8) in, REAL(4) out; callers src/mod_interp.f90:420, src/interp_to_hydro_section.f90:628-629),
 * over ncol independent columns in one launch.  Element (column c, level k), both 0-based, of an array lives at
 * base[c*cs + k*ks] (strides in elements, >= 0; cs = 0 shares one vector between all columns).  Targets outside the
 * source column come out as -6666 unless the persistence flags say otherwise, exactly as in the reference; a column
 * whose source values are all rmask_val is left at -6666 ("DOING NOTHING").  Inputs for which the reference itself is
 * undefined -- missing values elsewhere than in a leading run followed by a trailing run, or fewer than 3 valid source
 * points -- give SOSIE_ERR_REF_STOP.  status (ncol INTEGER(4), may be NULL; a DEVICE array in the _dev variant): per
 * column 0 done, 1 all missing, -1 / -2 the two undefined cases.                                                     */
int sosie_akima_1d_1d(sosie_ctx* ctx, int64_t ncol, int32_t nk1, const double* vz1, int64_t vz1_cs, int64_t vz1_ks,
                      const double* vf1, int64_t vf1_cs, int64_t vf1_ks, int32_t nk2, const double* vz2, int64_t vz2_cs,
                      int64_t vz2_ks, float* vf2, int64_t vf2_cs, int64_t vf2_ks, int32_t has_rmask, double rmask_val,
                      int32_t l_surf_persist, int32_t l_botm_persist, int32_t* status);
int sosie_akima_1d_1d_dev(sosie_ctx* ctx, int64_t ncol, int32_t nk1, const double* dvz1, int64_t vz1_cs, int64_t vz1_ks,
                          const double* dvf1, int64_t vf1_cs, int64_t vf1_ks, int32_t nk2, const double* dvz2,
                          int64_t vz2_cs, int64_t vz2_ks, float* dvf2, int64_t vf2_cs, int64_t vf2_ks, int32_t has_rmask,
                          double rmask_val, int32_t l_surf_persist, int32_t l_botm_persist, int32_t* dstatus);
/* the generic-coordinate (z <-> sigma) branch of INTERP_3D, src/mod_interp.f90:374-438: for every target column,
 * AKIMA_1D_1D from depth_src(ni,nj,nk_src) / data3d_tmp(ni,nj,nk_src) (the source depths and the field already carried
 * to the target grid) to depth_trg(ni,nj,nk_trg), sigma columns flipped (src_is_sigma / trg_is_sigma), the number of
 * target levels limited by the target mask when lmout (mask_trg(ni,nj,nk_trg) INTEGER(4), else may be NULL), persistence
 * below the last target level the source reaches.  data3d_trg(ni,nj,nk_trg) is INOUT (land columns and levels the
 * reference does not assign keep their values).  The reference flips its work arrays in place; here the three inputs
 * are left untouched.                                                                                                 */
int sosie_interp3d_sigma(sosie_ctx* ctx, int32_t ni, int32_t nj, int32_t nk_src, const float* depth_src,
                         const float* data3d_tmp, int32_t nk_trg, const float* depth_trg, const int32_t* mask_trg,
                         int32_t lmout, int32_t src_is_sigma, int32_t trg_is_sigma, float* data3d_trg);
int sosie_interp3d_sigma_dev(sosie_ctx* ctx, int32_t ni, int32_t nj, int32_t nk_src, const float* ddepth_src,
                             const float* ddata3d_tmp, int32_t nk_trg, const float* ddepth_trg, const int32_t* dmask_trg,
                             int32_t lmout, int32_t src_is_sigma, int32_t trg_is_sigma, float* ddata3d_trg);
/* lbc_lnk(nperio, pt2d, cd_type, psgn [, pval]) on nslab REAL(4) slabs X(jpi,jpj,nslab), src/mod_nemotools.f90:65-417
 * (E-W cyclic / closed, south row, north fold with T-point pivot nperio 3,4 or F-point pivot 5,6); cd_type is the
 * character code ('T','U','V','F','W').  Callers: src/mod_interp.f90:131,504, src/corr_vect.f90:629-632.            */
int sosie_lbc_lnk(sosie_ctx* ctx, int32_t nperio, int32_t jpi, int32_t jpj, int32_t nslab, float* X, int32_t cd_type,
                  double psgn, int32_t has_pval, double pval);
int sosie_lbc_lnk_dev(sosie_ctx* ctx, int32_t nperio, int32_t jpi, int32_t jpj, int32_t nslab, float* dX, int32_t cd_type,
                      double psgn, int32_t has_pval, double pval);
/* the post-interpolation block of INTERP_3D, src/mod_interp.f90:447-511, on X(ni,nj,nk) in place: persistence of the
 * last wet value into the sea-bed (ixtrpl_bot == 1), bounds vmin/vmax (rmiss_val untouched), lbc_lnk when
 * i_orca_trg > 0 (c_orca_trg = 'T','U','V','F'), target mask when lmout.  mask_trg(ni,nj,nk) INTEGER(4), may be NULL
 * when neither ixtrpl_bot == 1 nor lmout.  The optional SMOOTHER calls of that block are sosie_smoother.            */
int sosie_interp3d_post(sosie_ctx* ctx, int32_t ni, int32_t nj, int32_t nk, float* X, const int32_t* mask_trg,
                        int32_t ixtrpl_bot, float vmin, float vmax, float rmiss_val, int32_t i_orca_trg,
                        int32_t c_orca_trg, int32_t lmout);
int sosie_interp3d_post_dev(sosie_ctx* ctx, int32_t ni, int32_t nj, int32_t nk, float* dX, const int32_t* dmask_trg,
                            int32_t ixtrpl_bot, float vmin, float vmax, float rmiss_val, int32_t i_orca_trg,
                            int32_t c_orca_trg, int32_t lmout);

/* ANGLE2(nperio, plamt, pphit, plamu, pphiu, plamv, pphiv, plamf, pphif, gcost, gsint, ..., gcosf, gsinf),
 * src/mod_nemotools.f90:607-823: the rotation coefficients corr_vect feeds to the rotation
 * (src/corr_vect.f90:553-560).  All arrays REAL(8) (nx,ny).  nperio > 0: NEMO grid (lbc_lnk with psgn = -1), 0: Akima
 * extrapolation of the edges.  _dev: dcoords = the 8 coordinate arrays one after the other (lamt, phit, lamu, phiu, lamv,
 * phiv, lamf, phif), dangles = the 8 results in the reference's argument order.  Cells the reference leaves unassigned
 * are 0.                                                                                                            */
int sosie_angle2(sosie_ctx* ctx, int32_t nperio, int32_t nx, int32_t ny, const double* plamt, const double* pphit,
                 const double* plamu, const double* pphiu, const double* plamv, const double* pphiv, const double* plamf,
                 const double* pphif, double* gcost, double* gsint, double* gcosu, double* gsinu, double* gcosv,
                 double* gsinv, double* gcosf, double* gsinf);
int sosie_angle2_dev(sosie_ctx* ctx, int32_t nperio, int32_t nx, int32_t ny, const double* dcoords, double* dangles);

/* extrp_hl(X2d), src/mod_interp.f90:521-542: on a regular target the rows from jj_ex_top to the top and
 * from the bottom to jj_ex_btm (mod_conf globals, 0 = none; nlat_icr_trg = +1 / -1) take the last interpolated row.   */
int sosie_extrp_hl(sosie_ctx* ctx, int32_t ni, int32_t nj, int32_t nslab, float* X, int32_t jj_ex_top, int32_t jj_ex_btm,
                   int32_t nlat_icr_trg);
int sosie_extrp_hl_dev(sosie_ctx* ctx, int32_t ni, int32_t nj, int32_t nslab, float* dX, int32_t jj_ex_top,
                       int32_t jj_ex_btm, int32_t nlat_icr_trg);

/* ---- raw-binary twin of the mapping file (layout documented in csrc/mapfile.cu) ---------------------
 * Same variables, order, types and layout as the NetCDF file written by P2D_MAPPING_AB (src/io_ezcdf.f90:2195) and read
 * by RD_MAPPING_AB (:2280).  Host-side file I/O only (no context).  Return 0 or SOSIE_ERR_ARG (cannot open / bad file). */
int sosie_mapping_write_bin(const char* path, int32_t nx, int32_t ny, const double* lon, const double* lat,
                            const int32_t* metrics, const double* alphabeta, const int16_t* id_problem,
                            const float* dist_np /* nullable */, double vflag);
int sosie_mapping_read_header_bin(const char* path, int32_t* nx, int32_t* ny, int32_t* has_dist_np);
int sosie_mapping_read_bin(const char* path, int32_t nx, int32_t ny, double* lon /* nullable */, double* lat /* nullable */,
                           int32_t* metrics, double* alphabeta, int16_t* id_problem, float* dist_np /* nullable */);

#ifdef __cplusplus
}
#endif
#endif /* SOSIE_B200_H */
