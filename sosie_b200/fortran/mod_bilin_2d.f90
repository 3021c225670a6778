MODULE MOD_BILIN_2D
   !! Drop-in replacement of SOSIE's src/mod_bilin_2d.f90: same module name, same PUBLIC procedures and
   !! argument lists (BILIN_2D :44, MAPPING_BL :366, INTERP_BL :275), bodies are calls into libsosie_b200.
   !! sosie.f90, mod_interp, mod_conf, mod_init, mod_grids and io_ezcdf are untouched.
   !! NOT COMPILED IN THIS ENVIRONMENT (no Fortran compiler); shipped as source.
   USE, INTRINSIC :: ISO_C_BINDING
   USE mod_conf          !: l_reg_src, i_orca_trg, l_first_call_interp_routine, rmissval  (implicit inputs, SURVEY 8b)
   USE io_ezcdf, ONLY : RD_MAPPING_AB, P2D_MAPPING_AB, TEST_XYZ
   USE sosie_gpu_c
   IMPLICIT NONE
   PRIVATE
   PUBLIC :: BILIN_2D, MAPPING_BL, INTERP_BL

CONTAINS

   SUBROUTINE BILIN_2D(k_ew_per, X10, Y10, Z1, X20, Y20, Z2, cnpat,  mask_domain_trg)
      INTEGER,                 INTENT(in)  :: k_ew_per
      REAL(8), DIMENSION(:,:), INTENT(in)  :: X10, Y10
      REAL(4), DIMENSION(:,:), INTENT(in),  TARGET :: Z1      ! TARGET: C_LOC of the caller's own storage in asynchronous mode
      REAL(8), DIMENSION(:,:), INTENT(in)  :: X20, Y20
      REAL(4), DIMENSION(:,:), INTENT(out), TARGET :: Z2
      CHARACTER(len=*),        INTENT(in)  :: cnpat
      INTEGER, OPTIONAL, DIMENSION(:,:), INTENT(in), TARGET :: mask_domain_trg
      !!
      INTEGER :: nx1, ny1, nx2, ny2, is1d, it1d
      INTEGER(C_INT) :: ierr
      LOGICAL :: lefw
      CHARACTER(len=400) :: cf_wght
      REAL(8),    DIMENSION(:,:),   ALLOCATABLE, TARGET :: xs, ys, xt, yt     ! contiguous copies (assumed-shape actuals may be sections)
      REAL(4),    DIMENSION(:,:),   ALLOCATABLE, TARGET :: z1c, z2c
      INTEGER(4), DIMENSION(:,:),   ALLOCATABLE, TARGET :: mskc
      INTEGER(4), DIMENSION(:,:,:), ALLOCATABLE, TARGET :: imetrics
      REAL(8),    DIMENSION(:,:,:), ALLOCATABLE, TARGET :: rab
      INTEGER(2), DIMENSION(:,:),   ALLOCATABLE :: ipb
      REAL(4),    DIMENSION(:,:),   ALLOCATABLE :: d2np
      TYPE(C_PTR) :: pmask
      !!
      CALL GPU_INIT()
      nx1 = SIZE(Z1,1) ; ny1 = SIZE(Z1,2) ; nx2 = SIZE(Z2,1) ; ny2 = SIZE(Z2,2)
      IF ( l_gpu_async .AND. (.NOT. l_first_call_interp_routine) .AND. IS_CONTIGUOUS(Z1) .AND. IS_CONTIGUOUS(Z2) ) THEN
         !! asynchronous mode (sosie_gpu_c: GPU_ASYNC_BEGIN): queue this slab and return.  No copies: the library reads Z1 and
         !! fills Z2 in the caller's storage -- Z1 is copied at once when it is pageable, borrowed until GPU_FLUSH when it is
         !! page-locked (GPU_ALLOC_R4_3D); Z2 is complete after GPU_FLUSH().  The level loop of INTERP_3D passes
         !! data3d_src(:,:,jk) / data3d_tmp(:,:,jk): contiguous sections, passed by descriptor without copy-in/out.
         CALL GPU_CHECK( sosie_bilin_apply_enqueue(gpu_ctx, C_LOC(Z1), C_LOC(Z2)), 'BILIN_2D/enqueue' )
         RETURN
      END IF
      is1d = 0 ; IF ( TEST_XYZ(X10,Y10,Z1) == '1d' ) is1d = 1
      it1d = 0 ; IF ( TEST_XYZ(X20,Y20,Z2) == '1d' ) it1d = 1
      ALLOCATE ( z1c(nx1,ny1), z2c(nx2,ny2) ) ; z1c = Z1

      IF ( l_first_call_interp_routine ) THEN
         ALLOCATE ( xs(SIZE(X10,1),SIZE(X10,2)), ys(SIZE(Y10,1),SIZE(Y10,2)), xt(SIZE(X20,1),SIZE(X20,2)), yt(SIZE(Y20,1),SIZE(Y20,2)) )
         xs = X10 ; ys = Y10 ; xt = X20 ; yt = Y20
         pmask = C_NULL_PTR
         IF ( PRESENT(mask_domain_trg) ) THEN
            ALLOCATE ( mskc(nx2,ny2) ) ; mskc = mask_domain_trg ; pmask = C_LOC(mskc)
         END IF
         ALLOCATE ( imetrics(nx2,ny2,3), rab(nx2,ny2,2), ipb(nx2,ny2), d2np(nx2,ny2) )
         WRITE(cf_wght,'("sosie_mapping_",a,".nc")') TRIM(cnpat)
         INQUIRE(FILE=cf_wght, EXIST=lefw )
         IF ( lefw ) THEN
            !! same cache file as stock SOSIE: a CPU-built mapping is reused as is
            CALL GPU_CHECK( sosie_bilin_set_grids(gpu_ctx, k_ew_per, nx1, ny1, xs, ys, is1d, MERGE(1,0,l_reg_src), &
               &                                  nx2, ny2, xt, yt, it1d, pmask, i_orca_trg), 'BILIN_2D/set_grids' )
            CALL RD_MAPPING_AB(cf_wght, imetrics, rab, ipb)
            CALL GPU_CHECK( sosie_bilin_load_map(gpu_ctx, nx2, ny2, imetrics, rab, ipb), 'BILIN_2D/load_map' )
            CALL GPU_CHECK( sosie_bilin_apply(gpu_ctx, 1, z1c, z2c, 1), 'BILIN_2D/apply' )
         ELSE
            !! MAPPING_BL + first apply in ONE call: coordinate upload, search + weights, mapping download and the
            !! apply overlap on three streams (csrc/bilin_pipeline.cu); the arrays are page-locked for the duration
            ierr = sosie_host_register(C_LOC(xt),  INT(SIZE(xt),  C_SIZE_T)*8_C_SIZE_T)   ! best effort: unpinned arrays
            ierr = sosie_host_register(C_LOC(yt),  INT(SIZE(yt),  C_SIZE_T)*8_C_SIZE_T)   ! still work, the copies then
            ierr = sosie_host_register(C_LOC(rab), INT(SIZE(rab), C_SIZE_T)*8_C_SIZE_T)   ! do not overlap
            ierr = sosie_host_register(C_LOC(imetrics), INT(SIZE(imetrics), C_SIZE_T)*4_C_SIZE_T)
            ierr = sosie_host_register(C_LOC(z1c), INT(SIZE(z1c), C_SIZE_T)*4_C_SIZE_T)
            CALL GPU_CHECK( sosie_bilin_2d_first(gpu_ctx, k_ew_per, nx1, ny1, xs, ys, is1d, z1c, nx2, ny2, xt, yt, it1d, z2c, &
               &                                 pmask, MERGE(1,0,l_reg_src), i_orca_trg, 1, imetrics, rab, ipb, d2np), 'BILIN_2D/first' )
            ierr = sosie_host_unregister(C_LOC(xt)) ; ierr = sosie_host_unregister(C_LOC(yt)) ; ierr = sosie_host_unregister(C_LOC(rab))
            ierr = sosie_host_unregister(C_LOC(imetrics)) ; ierr = sosie_host_unregister(C_LOC(z1c))
            IF ( it1d == 1 ) THEN
               CALL P2D_MAPPING_AB(cf_wght, SPREAD(xt(:,1),2,ny2), SPREAD(yt(:,1),1,nx2), imetrics, rab, REAL(rmissval,8), ipb, d2np=d2np)
            ELSE
               CALL P2D_MAPPING_AB(cf_wght, xt, yt, imetrics, rab, REAL(rmissval,8), ipb, d2np=d2np)
            END IF
         END IF
         DEALLOCATE ( imetrics, rab, ipb, d2np, xs, ys, xt, yt )
      ELSE
         CALL GPU_CHECK( sosie_bilin_apply(gpu_ctx, 1, z1c, z2c, 0), 'BILIN_2D/apply' )
      END IF

      Z2 = z2c
      DEALLOCATE ( z1c, z2c )
      l_first_call_interp_routine = .FALSE.
   END SUBROUTINE BILIN_2D


   SUBROUTINE MAPPING_BL(k_ew_per, X1, Y1, lon_trg, lat_trg, cf_w,  mask_domain_trg)
      !! stand-alone use (interp_to_ground_track.f90:627, interp_to_hydro_section.f90:464): X1,Y1 are full 2-D arrays.
      !! NOTE: MAPPING_BL and sosie_find_nearest_point set their own grids in gpu_ctx, i.e. they replace the mapping a
      !! previous BILIN_2D cached there (the reference keeps RAB/IMETRICS in module SAVE variables that MAPPING_BL also
      !! overwrites through its mapping file); a program that needs both creates a second context (sosie_create).
      INTEGER,                 INTENT(in) :: k_ew_per
      REAL(8), DIMENSION(:,:), INTENT(in) :: X1, Y1
      REAL(8), DIMENSION(:,:), INTENT(in) :: lon_trg, lat_trg
      CHARACTER(len=*)       , INTENT(in) :: cf_w
      INTEGER, OPTIONAL ,DIMENSION(:,:), INTENT(in) :: mask_domain_trg
      INTEGER :: nxi, nyi, nxo, nyo
      REAL(8),    DIMENSION(:,:),   ALLOCATABLE :: xs, ys, xt, yt
      INTEGER(4), DIMENSION(:,:),   ALLOCATABLE, TARGET :: mskc
      INTEGER(4), DIMENSION(:,:,:), ALLOCATABLE :: imetrics
      REAL(8),    DIMENSION(:,:,:), ALLOCATABLE :: rab
      INTEGER(2), DIMENSION(:,:),   ALLOCATABLE :: ipb
      REAL(4),    DIMENSION(:,:),   ALLOCATABLE :: d2np
      TYPE(C_PTR) :: pmask
      CALL GPU_INIT()
      nxi = SIZE(X1,1) ; nyi = SIZE(X1,2) ; nxo = SIZE(lon_trg,1) ; nyo = SIZE(lon_trg,2)
      ALLOCATE ( xs(nxi,nyi), ys(nxi,nyi), xt(nxo,nyo), yt(nxo,nyo), imetrics(nxo,nyo,3), rab(nxo,nyo,2), ipb(nxo,nyo), d2np(nxo,nyo) )
      xs = X1 ; ys = Y1 ; xt = lon_trg ; yt = lat_trg
      pmask = C_NULL_PTR
      IF ( PRESENT(mask_domain_trg) ) THEN
         ALLOCATE ( mskc(nxo,nyo) ) ; mskc = mask_domain_trg ; pmask = C_LOC(mskc)
      END IF
      !! l_reg_src = 0: the caller already extended the arrays (or does not want the pole rows)
      CALL GPU_CHECK( sosie_bilin_set_grids(gpu_ctx, k_ew_per, nxi, nyi, xs, ys, 0, 0, nxo, nyo, xt, yt, 0, pmask, 0), 'MAPPING_BL/set_grids' )
      CALL GPU_CHECK( sosie_bilin_map(gpu_ctx), 'MAPPING_BL' )
      CALL GPU_CHECK( sosie_bilin_get_map(gpu_ctx, imetrics, rab, ipb, d2np, C_NULL_PTR), 'MAPPING_BL/get_map' )
      CALL P2D_MAPPING_AB(cf_w, lon_trg, lat_trg, imetrics, rab, REAL(rmissval,8), ipb, d2np=d2np)
      DEALLOCATE ( xs, ys, xt, yt, imetrics, rab, ipb, d2np )
   END SUBROUTINE MAPPING_BL


   FUNCTION INTERP_BL(k_ew_per, jiP, jjP, iqd, xa, xb, Z_in)
      !! Scalar form (src/mod_bilin_2d.f90:275-361), called once per track point by interp_to_ground_track (:750) and
      !! interp_to_hydro_section: four loads and a weighted mean, computed HERE on the host -- a device round trip (and an
      !! upload of the whole Z_in) per point would be orders of magnitude slower than the reference, and would disturb the
      !! mapping cached in gpu_ctx.  Same operations, order and REAL(4) roundings as the reference and as the packed
      !! records of the GPU apply kernels (csrc/pack.cuh).  Whole tracks go through sosie_track_interp (INTEGRATION.md 5b).
      INTEGER,                 INTENT(in) :: k_ew_per
      INTEGER,                 INTENT(in) :: jiP, jjP, iqd
      REAL(8),                 INTENT(in) :: xa, xb
      REAL(4), DIMENSION(:,:), INTENT(in) :: Z_in
      REAL(4) :: INTERP_BL
      REAL(4) :: wup, w1, w2, w3, w4
      INTEGER :: nxi, nyi, jiPm1, jiPp1, i1, j1, i2, j2, i3, j3, i4, j4
      nxi = SIZE(Z_in,1) ; nyi = SIZE(Z_in,2)
      jiPm1 = jiP-1 ; jiPp1 = jiP+1
      IF ( (jiPm1 ==   0  ).AND.(k_ew_per>=0) )  jiPm1 = nxi - k_ew_per
      IF ( (jiPp1 == nxi+1).AND.(k_ew_per>=0) )  jiPp1 = 1   + k_ew_per
      i1=0 ; j1=0 ; i2=0 ; j2=0 ; i3=0 ; j3=0 ; i4=0 ; j4=0
      SELECT CASE (iqd)
      CASE (1)
         i1=jiP ; j1=jjP ; i2=jiPp1 ; j2=jjP   ; i3=jiPp1 ; j3=jjP+1 ; i4=jiP   ; j4=jjP+1
      CASE (2)
         i1=jiP ; j1=jjP ; i2=jiP   ; j2=jjP-1 ; i3=jiPp1 ; j3=jjP-1 ; i4=jiPp1 ; j4=jjP
      CASE (3)
         i1=jiP ; j1=jjP ; i2=jiPm1 ; j2=jjP   ; i3=jiPm1 ; j3=jjP-1 ; i4=jiP   ; j4=jjP-1
      CASE (4)
         i1=jiP ; j1=jjP ; i2=jiP   ; j2=jjP+1 ; i3=jiPm1 ; j3=jjP+1 ; i4=jiPm1 ; j4=jjP
      END SELECT
      w1 = REAL( (1.d0 - xa)*(1.d0 - xb) , 4)
      w2 = REAL(       xa   *(1.d0 - xb) , 4)
      w3 = REAL(       xa   *    xb      , 4)
      w4 = REAL( (1.d0 - xa)*    xb      , 4)
      wup = w1 + w2 + w3 + w4
      IF     ( wup == 0. ) THEN
         INTERP_BL = -9998.
      ELSEIF ( (i1 < 1).OR.(i2 < 1).OR.(i3 < 1).OR.(i4 < 1) ) THEN
         INTERP_BL = -9997.
      ELSEIF ( (j1 < 1).OR.(j2 < 1).OR.(j3 < 1).OR.(j4 < 1) ) THEN
         INTERP_BL = -9996.
      ELSEIF ( (j1 > nyi).OR.(j2 > nyi).OR.(j3 > nyi).OR.(j4 > nyi) ) THEN
         INTERP_BL = -9995.
      ELSEIF ( (i1 > nxi).OR.(i2 > nxi).OR.(i3 > nxi).OR.(i4 > nxi) ) THEN
         INTERP_BL = -9994.
      ELSE
         INTERP_BL = ( Z_in(i1,j1)*w1 + Z_in(i2,j2)*w2 + Z_in(i3,j3)*w3 + Z_in(i4,j4)*w4 )/wup
      END IF
   END FUNCTION INTERP_BL

END MODULE MOD_BILIN_2D
