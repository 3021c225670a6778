MODULE SOSIE_GPU_C
   !! ISO_C_BINDING view of include/sosie_b200.h -- the only thing the drop-in modules below need.
   !! NOT COMPILED IN THIS ENVIRONMENT (no Fortran compiler in the image); the C ABI itself is exercised
   !! from Python (ctypes) by tests/.  Link:  ... -L<repo>/sosie_b200/_build -lsosie_b200 -lcudart
   USE, INTRINSIC :: ISO_C_BINDING
   IMPLICIT NONE
   TYPE(C_PTR), SAVE :: gpu_ctx = C_NULL_PTR   !: one context per process (the reference is single-threaded)

   INTERFACE
      INTEGER(C_INT) FUNCTION sosie_create(ctx, device) BIND(C, name='sosie_create')
         IMPORT :: C_PTR, C_INT
         TYPE(C_PTR), INTENT(out) :: ctx
         INTEGER(C_INT), VALUE    :: device
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_destroy(ctx) BIND(C, name='sosie_destroy')
         IMPORT :: C_PTR, C_INT
         TYPE(C_PTR), VALUE :: ctx
      END FUNCTION
      TYPE(C_PTR) FUNCTION sosie_last_error(ctx) BIND(C, name='sosie_last_error')
         IMPORT :: C_PTR
         TYPE(C_PTR), VALUE :: ctx
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_bilin_set_grids(ctx, k_ew, nx1, ny1, X10, Y10, src_1d, l_reg_src, &
         &                                          nx2, ny2, X20, Y20, trg_1d, mask, i_orca_trg) BIND(C, name='sosie_bilin_set_grids')
         IMPORT :: C_PTR, C_INT, C_DOUBLE
         TYPE(C_PTR), VALUE    :: ctx, mask            ! mask: C_NULL_PTR or c_loc(INTEGER(4) array)
         INTEGER(C_INT), VALUE :: k_ew, nx1, ny1, src_1d, l_reg_src, nx2, ny2, trg_1d, i_orca_trg
         REAL(C_DOUBLE), INTENT(in) :: X10(*), Y10(*), X20(*), Y20(*)
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_bilin_2d_first(ctx, k_ew, nx1, ny1, X10, Y10, src_1d, Z1, nx2, ny2, X20, Y20, trg_1d, Z2, mask, &
         &                                         l_reg_src, i_orca_trg, nslab, metrics, alphabeta, id_problem, dist_np) &
         &                                         BIND(C, name='sosie_bilin_2d_first')
         !! first call of BILIN_2D: MAPPING_BL + mapping arrays for P2D_MAPPING_AB + apply, pipelined (include/sosie_b200.h)
         IMPORT :: C_PTR, C_INT, C_DOUBLE, C_FLOAT, C_SHORT
         TYPE(C_PTR), VALUE    :: ctx, mask
         INTEGER(C_INT), VALUE :: k_ew, nx1, ny1, src_1d, nx2, ny2, trg_1d, l_reg_src, i_orca_trg, nslab
         REAL(C_DOUBLE), INTENT(in)  :: X10(*), Y10(*), X20(*), Y20(*)
         REAL(C_FLOAT),  INTENT(in)  :: Z1(*)
         REAL(C_FLOAT),  INTENT(out) :: Z2(*)
         INTEGER(C_INT)   :: metrics(*)
         REAL(C_DOUBLE)   :: alphabeta(*)
         INTEGER(C_SHORT) :: id_problem(*)
         REAL(C_FLOAT)    :: dist_np(*)
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_host_register(p, nbytes) BIND(C, name='sosie_host_register')
         IMPORT :: C_PTR, C_INT, C_SIZE_T
         TYPE(C_PTR), VALUE       :: p
         INTEGER(C_SIZE_T), VALUE :: nbytes
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_host_unregister(p) BIND(C, name='sosie_host_unregister')
         IMPORT :: C_PTR, C_INT
         TYPE(C_PTR), VALUE :: p
      END FUNCTION
      ! FIND_NEAREST_POINT (mod_manip.f90:834) for its direct callers (ij_from_lon_lat.f90:290); mask = C_NULL_PTR when the
      ! OPTIONAL mask_domain_trg is absent, else C_LOC of the INTEGER(4) array (inout)
      INTEGER(C_INT) FUNCTION sosie_find_nearest_point(ctx, nx_t, ny_t, Xtrg, Ytrg, nx_s, ny_s, Xsrc, Ysrc, JIp, JJp, mask) &
         &           BIND(C, name='sosie_find_nearest_point')
         IMPORT :: C_PTR, C_INT, C_DOUBLE
         TYPE(C_PTR), VALUE    :: ctx, mask
         INTEGER(C_INT), VALUE :: nx_t, ny_t, nx_s, ny_s
         REAL(C_DOUBLE)        :: Xtrg(*), Ytrg(*), Xsrc(*), Ysrc(*)
         INTEGER(C_INT)        :: JIp(*), JJp(*)
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_bilin_map(ctx) BIND(C, name='sosie_bilin_map')
         IMPORT :: C_PTR, C_INT
         TYPE(C_PTR), VALUE :: ctx
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_bilin_get_map(ctx, metrics, alphabeta, id_problem, dist_np, mask_out) BIND(C, name='sosie_bilin_get_map')
         IMPORT :: C_PTR, C_INT, C_DOUBLE, C_SHORT, C_FLOAT
         TYPE(C_PTR), VALUE :: ctx, mask_out
         INTEGER(C_INT)   :: metrics(*)
         REAL(C_DOUBLE)   :: alphabeta(*)
         INTEGER(C_SHORT) :: id_problem(*)
         REAL(C_FLOAT)    :: dist_np(*)
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_bilin_load_map(ctx, nx2, ny2, metrics, alphabeta, id_problem) BIND(C, name='sosie_bilin_load_map')
         IMPORT :: C_PTR, C_INT, C_DOUBLE, C_SHORT
         TYPE(C_PTR), VALUE    :: ctx
         INTEGER(C_INT), VALUE :: nx2, ny2
         INTEGER(C_INT), INTENT(in)   :: metrics(*)
         REAL(C_DOUBLE), INTENT(in)   :: alphabeta(*)
         INTEGER(C_SHORT), INTENT(in) :: id_problem(*)
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_bilin_apply(ctx, nslab, Z1, Z2, first_call) BIND(C, name='sosie_bilin_apply')
         IMPORT :: C_PTR, C_INT, C_FLOAT
         TYPE(C_PTR), VALUE    :: ctx
         INTEGER(C_INT), VALUE :: nslab, first_call
         REAL(C_FLOAT), INTENT(in)  :: Z1(*)
         REAL(C_FLOAT), INTENT(out) :: Z2(*)
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_interp_bl(ctx, k_ew, nxi, nyi, Z_in, n, jiP, jjP, iqd, xa, xb, zout) BIND(C, name='sosie_interp_bl')
         IMPORT :: C_PTR, C_INT, C_FLOAT, C_DOUBLE
         TYPE(C_PTR), VALUE    :: ctx
         INTEGER(C_INT), VALUE :: k_ew, nxi, nyi, n
         REAL(C_FLOAT), INTENT(in)  :: Z_in(*)
         INTEGER(C_INT), INTENT(in) :: jiP(*), jjP(*), iqd(*)
         REAL(C_DOUBLE), INTENT(in) :: xa(*), xb(*)
         REAL(C_FLOAT), INTENT(out) :: zout(*)
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_akima_set_grids(ctx, k_ew, nx1, ny1, X10, Y10, src_1d, i_orca_src, &
         &                                          nx2, ny2, X20, Y20, trg_1d) BIND(C, name='sosie_akima_set_grids')
         IMPORT :: C_PTR, C_INT, C_DOUBLE
         TYPE(C_PTR), VALUE    :: ctx
         INTEGER(C_INT), VALUE :: k_ew, nx1, ny1, src_1d, i_orca_src, nx2, ny2, trg_1d
         REAL(C_DOUBLE), INTENT(in) :: X10(*), Y10(*), X20(*), Y20(*)
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_akima_apply(ctx, nslab, Z1, Z2) BIND(C, name='sosie_akima_apply')
         IMPORT :: C_PTR, C_INT, C_FLOAT
         TYPE(C_PTR), VALUE    :: ctx
         INTEGER(C_INT), VALUE :: nslab
         REAL(C_FLOAT), INTENT(in)  :: Z1(*)
         REAL(C_FLOAT), INTENT(out) :: Z2(*)
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_bdrown(ctx, k_ew, ni, nj, nslab, X, mask, mask_per_slab, nb_inc, nb_smooth, &
         &                                 pmin, pmax, pval_land, nsweeps) BIND(C, name='sosie_bdrown')
         IMPORT :: C_PTR, C_INT, C_FLOAT
         TYPE(C_PTR), VALUE    :: ctx, nsweeps
         INTEGER(C_INT), VALUE :: k_ew, ni, nj, nslab, mask_per_slab, nb_inc, nb_smooth
         REAL(C_FLOAT)         :: X(*)
         INTEGER(C_INT), INTENT(in) :: mask(*)
         REAL(C_FLOAT), VALUE  :: pmin, pmax, pval_land
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_smoother(ctx, k_ew, ni, nj, nslab, X, nb_smooth, msk, l_emp) BIND(C, name='sosie_smoother')
         IMPORT :: C_PTR, C_INT, C_FLOAT
         TYPE(C_PTR), VALUE    :: ctx, msk
         INTEGER(C_INT), VALUE :: k_ew, ni, nj, nslab, nb_smooth, l_emp
         REAL(C_FLOAT)         :: X(*)
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_drown(ctx, k_ew, ni, nj, nslab, X, mask, nb_inc, nsweeps) BIND(C, name='sosie_drown')
         IMPORT :: C_PTR, C_INT, C_FLOAT, C_SIGNED_CHAR
         TYPE(C_PTR), VALUE    :: ctx, nsweeps
         INTEGER(C_INT), VALUE :: k_ew, ni, nj, nslab, nb_inc
         REAL(C_FLOAT)         :: X(*)
         INTEGER(C_SIGNED_CHAR), INTENT(in) :: mask(*)
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_akima_1d_3d(ctx, nx, ny, nk1, vz1, xf1, nk2, vz2, xf2, rfill_val) BIND(C, name='sosie_akima_1d_3d')
         !! AKIMA_1D_3D (src/mod_akima_1d.f90:243): the z -> z vertical interpolation of INTERP_3D (src/mod_interp.f90:366)
         IMPORT :: C_PTR, C_INT, C_FLOAT
         TYPE(C_PTR), VALUE    :: ctx
         INTEGER(C_INT), VALUE :: nx, ny, nk1, nk2
         REAL(C_FLOAT), INTENT(in)  :: vz1(*), xf1(*), vz2(*)
         REAL(C_FLOAT), INTENT(out) :: xf2(*)
         REAL(C_FLOAT), VALUE  :: rfill_val
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_set_exchange(ctx, nranks, fn, user) BIND(C, name='sosie_set_exchange')
         !! prelude exchange of a row-sharded first call: fn = C_FUNLOC of an all-gather (e.g. a wrapper of MPI_Allgather on bytes)
         IMPORT :: C_PTR, C_INT, C_FUNPTR
         TYPE(C_PTR), VALUE    :: ctx, user
         INTEGER(C_INT), VALUE :: nranks
         TYPE(C_FUNPTR), VALUE :: fn
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_track_interp(ctx, ni, nj, nk, F1, F2, vt1, vt2, imask, n, rt, jiP, jjP, iqd, xa, xb, r_obs, &
         &                                       obs_lo, obs_hi, clip_missing, rmissval, f_bl, f_np, fmask) BIND(C, name='sosie_track_interp')
         !! per-point block of interp_to_ground_track.f90:718-779 (F2 = second model record) / interp_to_hydro_section.f90:567-607 (F2 = C_NULL_PTR)
         IMPORT :: C_PTR, C_INT, C_INT64_T, C_FLOAT, C_DOUBLE, C_SIGNED_CHAR
         TYPE(C_PTR), VALUE        :: ctx, F2, rt       !! F2, rt: C_LOC(array) or C_NULL_PTR
         INTEGER(C_INT), VALUE     :: ni, nj, nk, clip_missing
         INTEGER(C_INT64_T), VALUE :: n
         REAL(C_FLOAT), INTENT(in) :: F1(*)
         REAL(C_DOUBLE), VALUE     :: vt1, vt2, obs_lo, obs_hi, rmissval
         INTEGER(C_SIGNED_CHAR), INTENT(in) :: imask(*)
         INTEGER(C_INT), INTENT(in) :: jiP(*), jjP(*), iqd(*)
         REAL(C_DOUBLE), INTENT(in) :: xa(*), xb(*), r_obs(*)
         REAL(C_DOUBLE)            :: f_bl(*), f_np(*)
         INTEGER(C_SIGNED_CHAR)    :: fmask(*)
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_akima_1d_1d(ctx, ncol, nk1, vz1, vz1_cs, vz1_ks, vf1, vf1_cs, vf1_ks, nk2, vz2, vz2_cs, vz2_ks, &
         &                                      vf2, vf2_cs, vf2_ks, has_rmask, rmask_val, l_surf_persist, l_botm_persist, status) &
         &                                      BIND(C, name='sosie_akima_1d_1d')
         !! AKIMA_1D_1D (src/mod_akima_1d.f90:26) over ncol columns; element (c,k), 0-based, of an array at base(1 + c*cs + k*ks)
         IMPORT :: C_PTR, C_INT, C_INT64_T, C_FLOAT, C_DOUBLE
         TYPE(C_PTR), VALUE       :: ctx
         INTEGER(C_INT64_T), VALUE :: ncol, vz1_cs, vz1_ks, vf1_cs, vf1_ks, vz2_cs, vz2_ks, vf2_cs, vf2_ks
         INTEGER(C_INT), VALUE    :: nk1, nk2, has_rmask, l_surf_persist, l_botm_persist
         REAL(C_DOUBLE), INTENT(in) :: vz1(*), vf1(*), vz2(*)
         REAL(C_FLOAT)            :: vf2(*)
         REAL(C_DOUBLE), VALUE    :: rmask_val
         TYPE(C_PTR), VALUE       :: status    !! C_NULL_PTR or C_LOC of an INTEGER(4) (ncol) array
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_interp3d_sigma(ctx, ni, nj, nk_src, depth_src, data3d_tmp, nk_trg, depth_trg, mask_trg, lmout, &
         &                                         src_is_sigma, trg_is_sigma, data3d_trg) BIND(C, name='sosie_interp3d_sigma')
         !! the z <-> sigma branch of INTERP_3D (src/mod_interp.f90:374-438), all target columns in one call
         IMPORT :: C_PTR, C_INT, C_FLOAT
         TYPE(C_PTR), VALUE    :: ctx
         INTEGER(C_INT), VALUE :: ni, nj, nk_src, nk_trg, lmout, src_is_sigma, trg_is_sigma
         REAL(C_FLOAT), INTENT(in) :: depth_src(*), data3d_tmp(*), depth_trg(*)
         TYPE(C_PTR), VALUE    :: mask_trg     !! C_LOC of INTEGER(4) (ni,nj,nk_trg), or C_NULL_PTR when .NOT. lmout
         REAL(C_FLOAT)         :: data3d_trg(*)
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_lbc_lnk(ctx, nperio, jpi, jpj, nslab, X, cd_type, psgn, has_pval, pval) BIND(C, name='sosie_lbc_lnk')
         !! lbc_lnk (src/mod_nemotools.f90:65) on nslab REAL(4) slabs; cd_type = ICHAR('T'), ...
         IMPORT :: C_PTR, C_INT, C_FLOAT, C_DOUBLE
         TYPE(C_PTR), VALUE    :: ctx
         INTEGER(C_INT), VALUE :: nperio, jpi, jpj, nslab, cd_type, has_pval
         REAL(C_FLOAT)         :: X(*)
         REAL(C_DOUBLE), VALUE :: psgn, pval
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_interp3d_post(ctx, ni, nj, nk, X, mask_trg, ixtrpl_bot, vmin, vmax, rmiss_val, i_orca_trg, &
         &                                        c_orca_trg, lmout) BIND(C, name='sosie_interp3d_post')
         !! post-interpolation block of INTERP_3D (src/mod_interp.f90:447-511); mask_trg: C_NULL_PTR or c_loc(INTEGER(4) (ni,nj,nk))
         IMPORT :: C_PTR, C_INT, C_FLOAT
         TYPE(C_PTR), VALUE    :: ctx, mask_trg
         INTEGER(C_INT), VALUE :: ni, nj, nk, ixtrpl_bot, i_orca_trg, c_orca_trg, lmout
         REAL(C_FLOAT)         :: X(*)
         REAL(C_FLOAT), VALUE  :: vmin, vmax, rmiss_val
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_extrp_hl(ctx, ni, nj, nslab, X, jj_ex_top, jj_ex_btm, nlat_icr_trg) BIND(C, name='sosie_extrp_hl')
         IMPORT :: C_PTR, C_INT, C_FLOAT
         TYPE(C_PTR), VALUE    :: ctx
         INTEGER(C_INT), VALUE :: ni, nj, nslab, jj_ex_top, jj_ex_btm, nlat_icr_trg
         REAL(C_FLOAT)         :: X(*)
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_rotate_uv(ctx, n, nslab, U, V, xcos, xsin, Uo, Vo, inverse) BIND(C, name='sosie_rotate_uv')
         IMPORT :: C_PTR, C_INT, C_FLOAT, C_DOUBLE, C_INT64_T
         TYPE(C_PTR), VALUE        :: ctx
         INTEGER(C_INT64_T), VALUE :: n
         INTEGER(C_INT), VALUE     :: nslab, inverse
         REAL(C_FLOAT), INTENT(in)  :: U(*), V(*)
         REAL(C_DOUBLE), INTENT(in) :: xcos(*), xsin(*)
         REAL(C_FLOAT), INTENT(out) :: Uo(*), Vo(*)
      END FUNCTION
      !! ---- asynchronous per-record / per-level calls (include/sosie_b200.h, csrc/queue.cu): the slab arguments are passed as
      !! TYPE(C_PTR) because the library reads / writes them AFTER the call returned (no Fortran temporary may be made)
      INTEGER(C_INT) FUNCTION sosie_bilin_apply_enqueue(ctx, Z1, Z2) BIND(C, name='sosie_bilin_apply_enqueue')
         IMPORT :: C_PTR, C_INT
         TYPE(C_PTR), VALUE :: ctx, Z1, Z2
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_akima_apply_enqueue(ctx, Z1, Z2) BIND(C, name='sosie_akima_apply_enqueue')
         IMPORT :: C_PTR, C_INT
         TYPE(C_PTR), VALUE :: ctx, Z1, Z2
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_bdrown_enqueue(ctx, k_ew, ni, nj, X, mask, nb_inc, nb_smooth, pmin, pmax, pval_land, nsweeps) &
         &           BIND(C, name='sosie_bdrown_enqueue')
         IMPORT :: C_PTR, C_INT, C_FLOAT
         TYPE(C_PTR), VALUE    :: ctx, X, mask, nsweeps
         INTEGER(C_INT), VALUE :: k_ew, ni, nj, nb_inc, nb_smooth
         REAL(C_FLOAT), VALUE  :: pmin, pmax, pval_land
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_flush(ctx) BIND(C, name='sosie_flush')
         IMPORT :: C_PTR, C_INT
         TYPE(C_PTR), VALUE :: ctx
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_set_queue_depth(ctx, depth) BIND(C, name='sosie_set_queue_depth')
         IMPORT :: C_PTR, C_INT
         TYPE(C_PTR), VALUE    :: ctx
         INTEGER(C_INT), VALUE :: depth
      END FUNCTION
      TYPE(C_PTR) FUNCTION sosie_host_alloc(nbytes) BIND(C, name='sosie_host_alloc')
         IMPORT :: C_PTR, C_SIZE_T
         INTEGER(C_SIZE_T), VALUE :: nbytes
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_host_free(p) BIND(C, name='sosie_host_free')
         IMPORT :: C_PTR, C_INT
         TYPE(C_PTR), VALUE :: p
      END FUNCTION
      ! ---- entries a maintainer needs beyond the four drop-in modules ----------------------------------------------------
      INTEGER(C_INT) FUNCTION sosie_sync(ctx) BIND(C, name='sosie_sync')
         !! waits for the context's stream; reports the deferred status of an asynchronous mapping
         IMPORT :: C_PTR, C_INT
         TYPE(C_PTR), VALUE :: ctx
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_bilin_set_shard(ctx, j0, j1) BIND(C, name='sosie_bilin_set_shard')
         !! this process owns target rows j0..j1 (1-based, inclusive; 0, 0 = all): one process per GPU (INTEGRATION.md 4b)
         IMPORT :: C_PTR, C_INT
         TYPE(C_PTR), VALUE    :: ctx
         INTEGER(C_INT), VALUE :: j0, j1
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_bilin_map_info(ctx, info) BIND(C, name='sosie_bilin_map_info')
         !! what FIND_NEAREST_POINT / MAPPING_BL print: search kind, row range, counts of problem points (include/sosie_b200.h)
         IMPORT :: C_PTR, C_INT
         TYPE(C_PTR), VALUE :: ctx
         INTEGER(C_INT)     :: info(*)
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_angle2(ctx, nperio, nx, ny, plamt, pphit, plamu, pphiu, plamv, pphiv, plamf, pphif, &
         &                                 gcost, gsint, gcosu, gsinu, gcosv, gsinv, gcosf, gsinf) BIND(C, name='sosie_angle2')
         !! ANGLE2 (mod_nemotools.f90:607-823) for corr_vect: the eight rotation arrays of the T / U / V / F points
         IMPORT :: C_PTR, C_INT, C_DOUBLE
         TYPE(C_PTR), VALUE    :: ctx
         INTEGER(C_INT), VALUE :: nperio, nx, ny
         REAL(C_DOUBLE), INTENT(in)  :: plamt(*), pphit(*), plamu(*), pphiu(*), plamv(*), pphiv(*), plamf(*), pphif(*)
         REAL(C_DOUBLE), INTENT(out) :: gcost(*), gsint(*), gcosu(*), gsinu(*), gcosv(*), gsinv(*), gcosf(*), gsinf(*)
      END FUNCTION
      ! binary twin of P2D_MAPPING_AB / RD_MAPPING_AB (io_ezcdf.f90:2195-2304) for builds without NetCDF; path: a
      ! NUL-terminated CHARACTER(KIND=C_CHAR) array (TRIM(cf)//C_NULL_CHAR); dist_np may be C_NULL_PTR
      INTEGER(C_INT) FUNCTION sosie_mapping_write_bin(path, nx, ny, lon, lat, metrics, alphabeta, id_problem, dist_np, vflag) &
         &           BIND(C, name='sosie_mapping_write_bin')
         IMPORT :: C_PTR, C_INT, C_DOUBLE, C_SHORT, C_CHAR
         CHARACTER(KIND=C_CHAR), INTENT(in) :: path(*)
         INTEGER(C_INT), VALUE        :: nx, ny
         REAL(C_DOUBLE), INTENT(in)   :: lon(*), lat(*), alphabeta(*)
         INTEGER(C_INT), INTENT(in)   :: metrics(*)
         INTEGER(C_SHORT), INTENT(in) :: id_problem(*)
         TYPE(C_PTR), VALUE           :: dist_np
         REAL(C_DOUBLE), VALUE        :: vflag
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_mapping_read_header_bin(path, nx, ny, has_dist_np) BIND(C, name='sosie_mapping_read_header_bin')
         IMPORT :: C_INT, C_CHAR
         CHARACTER(KIND=C_CHAR), INTENT(in) :: path(*)
         INTEGER(C_INT), INTENT(out) :: nx, ny, has_dist_np
      END FUNCTION
      INTEGER(C_INT) FUNCTION sosie_mapping_read_bin(path, nx, ny, lon, lat, metrics, alphabeta, id_problem, dist_np) &
         &           BIND(C, name='sosie_mapping_read_bin')
         IMPORT :: C_PTR, C_INT, C_DOUBLE, C_SHORT, C_CHAR
         CHARACTER(KIND=C_CHAR), INTENT(in) :: path(*)
         INTEGER(C_INT), VALUE         :: nx, ny
         TYPE(C_PTR), VALUE            :: lon, lat, dist_np      ! C_NULL_PTR: not wanted
         INTEGER(C_INT), INTENT(out)   :: metrics(*)
         REAL(C_DOUBLE), INTENT(out)   :: alphabeta(*)
         INTEGER(C_SHORT), INTENT(out) :: id_problem(*)
      END FUNCTION
   END INTERFACE

   !! Asynchronous mode of the drop-in modules (GPU_ASYNC_BEGIN / GPU_FLUSH below).  While it is on, the non-first calls of
   !! BILIN_2D / AKIMA_2D and every call of BDROWN with CONTIGUOUS array arguments only QUEUE their slab
   !! (sosie_*_enqueue) and return; the output arrays are complete after GPU_FLUSH().  The caller decides where the results
   !! are needed: one CALL GPU_FLUSH() after each level loop of INTERP_3D, one per K records in the time loop (INTEGRATION.md 4).
   LOGICAL, SAVE :: l_gpu_async = .FALSE.

CONTAINS

   SUBROUTINE GPU_ASYNC_BEGIN(kdepth)
      !! kdepth: slabs coalesced into one batched launch (default 32)
      INTEGER, OPTIONAL, INTENT(in) :: kdepth
      CALL GPU_INIT()
      IF ( PRESENT(kdepth) ) CALL GPU_CHECK( sosie_set_queue_depth(gpu_ctx, INT(kdepth,C_INT)), 'GPU_ASYNC_BEGIN' )
      l_gpu_async = .TRUE.
   END SUBROUTINE GPU_ASYNC_BEGIN

   SUBROUTINE GPU_FLUSH()
      !! every slab queued so far is in its output array when this returns
      IF ( C_ASSOCIATED(gpu_ctx) ) CALL GPU_CHECK( sosie_flush(gpu_ctx), 'GPU_FLUSH' )
   END SUBROUTINE GPU_FLUSH

   SUBROUTINE GPU_ASYNC_END()
      CALL GPU_FLUSH()
      l_gpu_async = .FALSE.
   END SUBROUTINE GPU_ASYNC_END

   SUBROUTINE GPU_ALLOC_R4_3D(p, n1, n2, n3)
      !! page-locked REAL(4) (n1,n2,n3) array for record / level rings: enqueue reads it by DMA in place (no staging copy)
      REAL(4), DIMENSION(:,:,:), POINTER, INTENT(out) :: p
      INTEGER, INTENT(in) :: n1, n2, n3
      TYPE(C_PTR) :: cp
      cp = sosie_host_alloc( INT(n1,C_SIZE_T)*INT(n2,C_SIZE_T)*INT(n3,C_SIZE_T)*4_C_SIZE_T )
      IF ( .NOT. C_ASSOCIATED(cp) ) THEN
         PRINT *, 'ERROR (sosie_gpu_c): sosie_host_alloc failed'; STOP
      END IF
      CALL C_F_POINTER(cp, p, (/ n1, n2, n3 /))
   END SUBROUTINE GPU_ALLOC_R4_3D

   SUBROUTINE GPU_INIT()
      INTEGER(C_INT) :: ierr
      IF ( .NOT. C_ASSOCIATED(gpu_ctx) ) THEN
         ierr = sosie_create(gpu_ctx, 0_C_INT)
         IF ( ierr /= 0 ) THEN
            PRINT *, 'ERROR (sosie_gpu_c): no usable CUDA device -- this build has no CPU fallback'; STOP
         END IF
      END IF
   END SUBROUTINE GPU_INIT

   SUBROUTINE GPU_CHECK(ierr, cwhere)
      !! the C ABI returns codes where the reference PRINTs and STOPs: keep that behaviour
      INTEGER(C_INT),   INTENT(in) :: ierr
      CHARACTER(len=*), INTENT(in) :: cwhere
      CHARACTER(kind=C_CHAR), POINTER :: cmsg(:)
      INTEGER :: k
      IF ( ierr /= 0 ) THEN
         CALL C_F_POINTER(sosie_last_error(gpu_ctx), cmsg, (/ 512 /))
         k = 1
         DO WHILE ( (k < 512) .AND. (cmsg(k) /= C_NULL_CHAR) )
            k = k + 1
         END DO
         PRINT *, 'ERROR in ', TRIM(cwhere), ' (sosie_b200, code', ierr, '): ', cmsg(1:k-1)
         STOP
      END IF
   END SUBROUTINE GPU_CHECK

END MODULE SOSIE_GPU_C
