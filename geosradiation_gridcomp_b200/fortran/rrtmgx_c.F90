! rrtmgx_c.F90 -- ISO_C_BINDING view of include/rrtmgx.h (the B200 RRTMG LW + SW + McICA library).
!
! Source only: no Fortran compiler exists in the build image, so this shim is exercised through
! the same C ABI from the C++/Python tests.  A GEOS build adds this directory to the
! GEOSirrad_GridComp / GEOSsolar_GridComp CMake targets in place of the RRTMG src/ trees and
! links librrtmgx.so (see INTEGRATION.md).
!
! Default `real` may be 8 bytes (-fdefault-real-8 / -r8: the fp64 contract the parity tests hold the
! library to) or 4 bytes (the production kind of GEOS): rrtmgx_real_flags then carries
! RRTMGX_F32_ARRAYS, the library widens the arrays exactly while it stages them, computes in fp64 and
! rounds the outputs once.  Any other kind is a compile-time error.
module rrtmgx_c
   use, intrinsic :: iso_c_binding
   implicit none
   public

   integer, parameter :: rrtmgx_real_kind = kind(1.0)
   ! compile-time trap: the array size is negative unless default real is c_double or c_float
   integer, parameter, private :: real_is_c_real(2*merge(1, -1, rrtmgx_real_kind == c_double .or. &
                                                                 rrtmgx_real_kind == c_float) - 1) = 0

   integer(c_int), parameter :: RRTMGX_DEVICE_PTRS = 1, RRTMGX_NO_SYNC = 2, RRTMGX_SKIP_CHECKS = 4, &
                                RRTMGX_KEEP_STATUS = 8, RRTMGX_REUSE_CLOUDS = 16, RRTMGX_F32_ARRAYS = 32, &
                                RRTMGX_LIT_ONLY = 64
   ! what every shim ORs into `flags`: the element kind of the caller's real arrays
   integer, parameter :: RRTMGX_NRADVAL = 120   ! the SOLAR_RADVAL dummies of rrtmg_sw (rrtmg_sw_rad.F90:85-122)
   integer(c_int), parameter :: rrtmgx_real_flags = merge(0_c_int, RRTMGX_F32_ARRAYS, rrtmgx_real_kind == c_double)

   type, bind(C) :: rrtmgx_config
      type(c_ptr)    :: table_blob = c_null_ptr
      integer(c_int) :: device = -1
      integer(c_int) :: inhomogeneity = -1   ! -1: leave the McICA state alone (the *_ini shims); 0..2 sets ih at the first init
      type(c_ptr)    :: corr = c_null_ptr
   end type

   ! field order == RrtmgxLwArgs in include/rrtmgx.h
   type, bind(C) :: rrtmgx_lw_args
      integer(c_int) :: ncol, nlay, psize, dudTs, iceflglw, liqflglw, dyofyr, cloudLM, cloudMH, flags
      type(c_ptr) :: stream
      type(c_ptr) :: play, plev, tlay, tlev, tsfc, emis
      type(c_ptr) :: h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr
      type(c_ptr) :: cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr
      type(c_ptr) :: cldf, ciwp, clwp, rei, rel
      type(c_ptr) :: tauaer, zm, alat
      type(c_ptr) :: band_output
      type(c_ptr) :: clearCounts
      type(c_ptr) :: uflx, dflx, uflxc, dflxc, duflx_dTs, duflxc_dTs, olrb, dolrb_dTs
   end type

   ! field order == RrtmgxLwVariants in include/rrtmgx.h; gas: 1 H2O 2 O3 3 CO2 4 CH4 5 N2O 6 CFC11 7 CFC12 8 HCFC22
   type, bind(C) :: rrtmgx_lw_variants
      integer(c_int) :: nvar
      type(c_ptr) :: gas
      type(c_ptr) :: uflx, dflx, duflx_dTs
   end type

   ! field order == RrtmgxSwNoAerosol in include/rrtmgx.h
   type, bind(C) :: rrtmgx_sw_no_aerosol
      type(c_ptr) :: swuflx, swdflx, swuflxc, swdflxc, fswband
   end type

   ! field order == RrtmgxSwArgs in include/rrtmgx.h
   type, bind(C) :: rrtmgx_sw_args
      integer(c_int) :: ncol, nlay, rpart, isolvar, iceflgsw, liqflgsw, dyofyr, cloudLM, cloudMH
      integer(c_int) :: iaer, normFlx, do_drfband, flags
      type(c_ptr) :: stream
      real(c_double) :: scon, adjes
      type(c_ptr) :: bndscl, indsolvar, solcycfrac
      type(c_ptr) :: coszen, play, plev, tlay
      type(c_ptr) :: h2ovmr, o3vmr, co2vmr, ch4vmr, o2vmr
      type(c_ptr) :: cld, ciwp, clwp, rei, rel, zm, alat
      type(c_ptr) :: tauaer, ssaaer, asmaer
      type(c_ptr) :: asdir, asdif, aldir, aldif
      type(c_ptr) :: clearCounts
      type(c_ptr) :: swuflx, swdflx, swuflxc, swdflxc
      type(c_ptr) :: nirr, nirf, parr, parf, uvrr, uvrf
      type(c_ptr) :: fswband
      type(c_ptr) :: cotdtp, cotdhp, cotdmp, cotdlp, cotntp, cotnhp, cotnmp, cotnlp
      type(c_ptr) :: drband, dfband
      type(c_ptr) :: radval   ! c_null_ptr, or (ncol,RRTMGX_NRADVAL): the SOLAR_RADVAL build
   end type

   ! fused Run-phase glue; field order == RrtmgxIrradArgs in include/rrtmgx.h
   type, bind(C) :: rrtmgx_irrad_args
      integer(c_int) :: ncol, lm, iceflg, liqflg, doy, lcldmh, lcldlm, flags
      type(c_ptr) :: stream
      real(c_double) :: co2_fixed, o2, ccl4, airmw, h2omw, o3mw, rgas, grav
      type(c_ptr) :: ple, pl, t, q, o3, ch4, n2o, co2, cfc11, cfc12, hcfc22, fcld
      type(c_ptr) :: qliq, qice, rliq, rice
      type(c_ptr) :: ts, t2m, emis, lats
      type(c_ptr) :: taua, ssaa
      type(c_ptr) :: band_output
      type(c_ptr) :: flxu, flxd, flcu, flcd, dfdts, dfdtsc
      type(c_ptr) :: sfcem
      type(c_ptr) :: cldtt, cldhi, cldmd, cldlo
      type(c_ptr) :: olrb, dolrb_dts
   end type

   ! field order == RrtmgxSolarArgs in include/rrtmgx.h
   type, bind(C) :: rrtmgx_solar_args
      integer(c_int) :: ncol, lm, iceflg, liqflg, doy, isolvar, lcldmh, lcldlm, flags
      type(c_ptr) :: stream
      real(c_double) :: sc, dist, co2, o2, airmw, h2omw, o3mw, rgas, grav, undef
      type(c_ptr) :: solcycfrac
      type(c_ptr) :: ple, pl, t, q, o3, ch4, cl
      type(c_ptr) :: qliq, qice, rliq, rice
      type(c_ptr) :: ts, zt, lats
      type(c_ptr) :: albvr, albvf, albnr, albnf
      type(c_ptr) :: taua, ssaa, asya
      type(c_ptr) :: fsw, fsc, fswu, fscu
      type(c_ptr) :: nirr, nirf, parr, parf, uvrr, uvrf
      type(c_ptr) :: fswband
      type(c_ptr) :: cldts, cldhs, cldms, cldls
      type(c_ptr) :: cottp, cothp, cotmp, cotlp
   end type

   interface
      integer(c_int) function rrtmgx_init(cfg) bind(C, name='rrtmgx_init')
         import :: c_int, rrtmgx_config
         type(rrtmgx_config), intent(in) :: cfg
      end function
      integer(c_int) function rrtmgx_set_mcica(ih, corr) bind(C, name='rrtmgx_set_mcica')
         import :: c_int, c_double
         integer(c_int), value :: ih
         real(c_double), intent(in) :: corr(8)
      end function
      integer(c_int) function rrtmgx_finalize() bind(C, name='rrtmgx_finalize')
         import :: c_int
      end function
      integer(c_int) function rrtmgx_lw_run(a) bind(C, name='rrtmgx_lw_run')
         import :: c_int, rrtmgx_lw_args
         type(rrtmgx_lw_args), intent(in) :: a
      end function
      ! removed-gas loop + main call in one (IRR:3405-3478); v == RrtmgxLwVariants
      integer(c_int) function rrtmgx_lw_run_variants(a, v) bind(C, name='rrtmgx_lw_run_variants')
         import :: c_int, rrtmgx_lw_args, rrtmgx_lw_variants
         type(rrtmgx_lw_args), intent(in) :: a
         type(rrtmgx_lw_variants), intent(in) :: v
      end function
      integer(c_int) function rrtmgx_sw_run(a) bind(C, name='rrtmgx_sw_run')
         import :: c_int, rrtmgx_sw_args
         type(rrtmgx_sw_args), intent(in) :: a
      end function
      integer(c_int) function rrtmgx_irrad_refresh(a) bind(C, name='rrtmgx_irrad_refresh')
         import :: c_int, rrtmgx_irrad_args
         type(rrtmgx_irrad_args), intent(in) :: a
      end function
      integer(c_int) function rrtmgx_solar_refresh(a) bind(C, name='rrtmgx_solar_refresh')
         import :: c_int, rrtmgx_solar_args
         type(rrtmgx_solar_args), intent(in) :: a
      end function
      ! no-aerosol + regular pass in one (SOL:3249-3287)
      integer(c_int) function rrtmgx_sw_run_with_clean(a, na) bind(C, name='rrtmgx_sw_run_with_clean')
         import :: c_int, rrtmgx_sw_args, rrtmgx_sw_no_aerosol
         type(rrtmgx_sw_args), intent(in) :: a
         type(rrtmgx_sw_no_aerosol), intent(in) :: na
      end function
      integer(c_int) function rrtmgx_heating_rate(ncol, nlay, fnet, plev, hr, grav, cp, flags, stream) &
            bind(C, name='rrtmgx_heating_rate')
         import :: c_int, c_double, c_ptr
         integer(c_int), value :: ncol, nlay, flags
         real(c_double), intent(in) :: fnet(*), plev(*)
         real(c_double), intent(out) :: hr(*)
         real(c_double), value :: grav, cp
         type(c_ptr), value :: stream
      end function
      type(c_ptr) function rrtmgx_strerror(status) bind(C, name='rrtmgx_strerror')
         import :: c_int, c_ptr
         integer(c_int), value :: status
      end function
   end interface

contains

   ! message text of a status code as a Fortran string
   function rrtmgx_message(status) result(msg)
      integer(c_int), intent(in) :: status
      character(len=:), allocatable :: msg
      character(kind=c_char), pointer :: p(:)
      integer :: n
      call c_f_pointer(rrtmgx_strerror(status), p, [256])
      n = 0
      do while (n < 256)
         if (p(n+1) == c_null_char) exit
         n = n + 1
      end do
      allocate(character(len=n) :: msg)
      msg = transfer(p(1:n), msg)
   end function

end module rrtmgx_c
