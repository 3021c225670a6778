MODULE MOD_BDROWN
   !! Drop-in replacement of SOSIE's src/mod_bdrown.f90 (BDROWN :18, SMOOTHER :444).  NOT COMPILED HERE.
   USE, INTRINSIC :: ISO_C_BINDING
   USE sosie_gpu_c
   IMPLICIT NONE
   PRIVATE
   PUBLIC :: bdrown, smoother
CONTAINS
   SUBROUTINE BDROWN(k_ew, X, mask,   nb_inc, nb_smooth, pmin, pmax, pval_land )
      INTEGER,                    INTENT(in)    :: k_ew
      REAL(4),    DIMENSION(:,:), INTENT(inout), TARGET :: X
      INTEGER,    DIMENSION(:,:), INTENT(in),    TARGET :: mask
      INTEGER,    OPTIONAL,       INTENT(in)    :: nb_inc, nb_smooth
      REAL(4),    OPTIONAL,       INTENT(in)    :: pmin, pmax, pval_land
      !! zmin/zmax/zval_land keep the reference's implicit SAVE (initialised locals, mod_bdrown.f90:56)
      REAL(4), SAVE :: zmin=-1.E24, zmax=1.E24, zval_land=0.
      INTEGER :: ninc_max, nsmooth, ni, nj
      REAL(4),    DIMENSION(:,:), ALLOCATABLE :: xc
      INTEGER(4), DIMENSION(:,:), ALLOCATABLE :: mc
      ninc_max = 200 ; IF ( PRESENT(nb_inc) )    ninc_max = nb_inc
      nsmooth  = 0   ; IF ( PRESENT(nb_smooth) ) nsmooth  = nb_smooth
      IF( PRESENT(pmin)      ) zmin      = pmin
      IF( PRESENT(pmax)      ) zmax      = pmax
      IF( PRESENT(pval_land) ) zval_land = pval_land
      IF ( (SIZE(X,1)/=SIZE(mask,1)) .OR. (SIZE(X,2)/=SIZE(mask,2)) ) THEN
         PRINT *, 'ERROR, mod_bdrown.F90 => `BDROWN()` : data and mask arrays have different shapes !!!'; STOP
      END IF
      CALL GPU_INIT()
      ni = SIZE(X,1) ; nj = SIZE(X,2)
      IF ( l_gpu_async .AND. IS_CONTIGUOUS(X) .AND. IS_CONTIGUOUS(mask) .AND. (KIND(mask) == 4) ) THEN
         !! asynchronous mode: queue the level (mod_interp.f90:213 passes data3d_src(:,:,jk), mask_src(:,:,jk)); up to
         !! `depth` levels run in ONE launch (one CTA per slab for ORCA-sized slabs); X is drowned after GPU_FLUSH()
         CALL GPU_CHECK( sosie_bdrown_enqueue(gpu_ctx, k_ew, ni, nj, C_LOC(X), C_LOC(mask), ninc_max, nsmooth, zmin, zmax, zval_land, &
            &                                 C_NULL_PTR), 'BDROWN/enqueue' )
         RETURN
      END IF
      ALLOCATE ( xc(ni,nj), mc(ni,nj) ) ; xc = X ; mc = mask      ! contiguous copies (callers pass array sections)
      CALL GPU_CHECK( sosie_bdrown(gpu_ctx, k_ew, ni, nj, 1, xc, mc, 0, ninc_max, nsmooth, zmin, zmax, zval_land, C_NULL_PTR), 'BDROWN' )
      X = xc
      DEALLOCATE ( xc, mc )
   END SUBROUTINE BDROWN

   SUBROUTINE SMOOTHER(k_ew, X,  nb_smooth, msk, l_exclude_mask_points)
      INTEGER,                 INTENT(in)           :: k_ew
      REAL(4), DIMENSION(:,:), INTENT(inout)        :: X
      INTEGER, OPTIONAL,                 INTENT(in) :: nb_smooth
      INTEGER, OPTIONAL, DIMENSION(:,:), INTENT(in) :: msk
      LOGICAL, OPTIONAL                , INTENT(in) :: l_exclude_mask_points
      INTEGER :: nsmooth, ni, nj, lemp
      REAL(4),    DIMENSION(:,:), ALLOCATABLE :: xc
      INTEGER(4), DIMENSION(:,:), ALLOCATABLE, TARGET :: mc
      TYPE(C_PTR) :: pm
      nsmooth = 10 ; IF ( PRESENT(nb_smooth) ) nsmooth = nb_smooth
      lemp = 0     ; IF ( PRESENT(l_exclude_mask_points) ) lemp = 1    ! presence, not value (mod_bdrown.f90:496)
      CALL GPU_INIT()
      ni = SIZE(X,1) ; nj = SIZE(X,2)
      ALLOCATE ( xc(ni,nj) ) ; xc = X
      pm = C_NULL_PTR
      IF ( PRESENT(msk) ) THEN
         ALLOCATE ( mc(ni,nj) ) ; mc = msk ; pm = C_LOC(mc)
      END IF
      CALL GPU_CHECK( sosie_smoother(gpu_ctx, k_ew, ni, nj, 1, xc, nsmooth, pm, lemp), 'SMOOTHER' )
      X = xc
      DEALLOCATE ( xc )
   END SUBROUTINE SMOOTHER
END MODULE MOD_BDROWN
