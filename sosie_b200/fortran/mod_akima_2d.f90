MODULE MOD_AKIMA_2D
   !! Drop-in replacement of SOSIE's src/mod_akima_2d.f90 (AKIMA_2D :59).  NOT COMPILED HERE (no Fortran compiler).
   USE, INTRINSIC :: ISO_C_BINDING
   USE mod_conf          !: i_orca_src, l_first_call_interp_routine
   USE io_ezcdf, ONLY : TEST_XYZ
   USE sosie_gpu_c
   IMPLICIT NONE
   LOGICAL, PUBLIC, SAVE :: l_always_first_call = .FALSE.
   PRIVATE
   PUBLIC :: AKIMA_2D
CONTAINS
   SUBROUTINE AKIMA_2D(k_ew_per, X10, Y10, Z1, X20, Y20, Z2,    icall)
      INTEGER,                 INTENT(in)  :: k_ew_per
      REAL(8), DIMENSION(:,:), INTENT(in)  :: X10, Y10
      REAL(4), DIMENSION(:,:), INTENT(in),  TARGET :: Z1
      REAL(8), DIMENSION(:,:), INTENT(in)  :: X20, Y20
      REAL(4), DIMENSION(:,:), INTENT(out), TARGET :: Z2
      INTEGER,       OPTIONAL, INTENT(in)  :: icall
      INTEGER :: nx1, ny1, nx2, ny2, is1d, it1d
      REAL(8), DIMENSION(:,:), ALLOCATABLE :: xs, ys, xt, yt
      REAL(4), DIMENSION(:,:), ALLOCATABLE :: z1c, z2c
      IF ( PRESENT(icall) ) THEN
         IF ( icall == 1 ) THEN
            l_first_call_interp_routine = .TRUE. ; l_always_first_call = .TRUE.
         END IF
      END IF
      CALL GPU_INIT()
      nx1 = SIZE(Z1,1) ; ny1 = SIZE(Z1,2) ; nx2 = SIZE(Z2,1) ; ny2 = SIZE(Z2,2)
      IF ( l_first_call_interp_routine ) THEN
         !! coordinate frames, source range and the ixy_pos table are built once and cached on the device
         is1d = 0 ; IF ( TEST_XYZ(X10,Y10,Z1) == '1d' ) is1d = 1
         it1d = 0 ; IF ( TEST_XYZ(X20,Y20,Z2) == '1d' ) it1d = 1
         ALLOCATE ( xs(SIZE(X10,1),SIZE(X10,2)), ys(SIZE(Y10,1),SIZE(Y10,2)), xt(SIZE(X20,1),SIZE(X20,2)), yt(SIZE(Y20,1),SIZE(Y20,2)) )
         xs = X10 ; ys = Y10 ; xt = X20 ; yt = Y20
         CALL GPU_CHECK( sosie_akima_set_grids(gpu_ctx, k_ew_per, nx1, ny1, xs, ys, is1d, i_orca_src, nx2, ny2, xt, yt, it1d), 'AKIMA_2D/set_grids' )
         DEALLOCATE ( xs, ys, xt, yt )
      END IF
      IF ( l_gpu_async .AND. IS_CONTIGUOUS(Z1) .AND. IS_CONTIGUOUS(Z2) ) THEN
         !! asynchronous mode: queue the slab (grids are set above on the first call); Z2 is complete after GPU_FLUSH()
         CALL GPU_CHECK( sosie_akima_apply_enqueue(gpu_ctx, C_LOC(Z1), C_LOC(Z2)), 'AKIMA_2D/enqueue' )
         l_first_call_interp_routine = .FALSE.
         IF ( l_always_first_call ) l_first_call_interp_routine = .TRUE.
         RETURN
      END IF
      ALLOCATE ( z1c(nx1,ny1), z2c(nx2,ny2) ) ; z1c = Z1
      CALL GPU_CHECK( sosie_akima_apply(gpu_ctx, 1, z1c, z2c), 'AKIMA_2D' )
      Z2 = z2c
      DEALLOCATE ( z1c, z2c )
      l_first_call_interp_routine = .FALSE.
      IF ( l_always_first_call ) l_first_call_interp_routine = .TRUE.
   END SUBROUTINE AKIMA_2D
END MODULE MOD_AKIMA_2D
