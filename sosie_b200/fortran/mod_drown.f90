MODULE MOD_DROWN
   !! Drop-in replacement of SOSIE's src/mod_drown.f90 (DROWN :18; SMOOTHER is re-exported from mod_bdrown,
   !! the two reference copies are identical).  NOT COMPILED HERE.
   USE, INTRINSIC :: ISO_C_BINDING
   USE sosie_gpu_c
   USE mod_bdrown, ONLY : smoother
   IMPLICIT NONE
   PRIVATE
   PUBLIC :: drown, smoother
CONTAINS
   SUBROUTINE DROWN(k_ew, X, mask,   nb_inc)
      INTEGER,                       INTENT(in)    :: k_ew
      REAL(4),    DIMENSION(:,:),    INTENT(inout) :: X
      INTEGER(1), DIMENSION(:,:),    INTENT(in)    :: mask
      INTEGER,    OPTIONAL,          INTENT(in)    :: nb_inc
      INTEGER :: ninc_max, ni, nj
      REAL(4),    DIMENSION(:,:), ALLOCATABLE :: xc
      INTEGER(1), DIMENSION(:,:), ALLOCATABLE :: mc
      ninc_max = 200 ; IF ( PRESENT(nb_inc) ) ninc_max = nb_inc
      IF ( (SIZE(X,1) /= SIZE(mask,1)).OR.(SIZE(X,2) /= SIZE(mask,2)) ) THEN
         PRINT *, 'ERROR, mod_drown.F90 => DROWN : size of data and mask do not match!!!'; STOP
      END IF
      CALL GPU_INIT()
      ni = SIZE(X,1) ; nj = SIZE(X,2)
      ALLOCATE ( xc(ni,nj), mc(ni,nj) ) ; xc = X ; mc = mask
      CALL GPU_CHECK( sosie_drown(gpu_ctx, k_ew, ni, nj, 1, xc, mc, ninc_max, C_NULL_PTR), 'DROWN' )
      X = xc
      DEALLOCATE ( xc, mc )
   END SUBROUTINE DROWN
END MODULE MOD_DROWN
