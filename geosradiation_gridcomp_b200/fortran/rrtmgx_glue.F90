! rrtmgx_glue.F90 -- driver-level entry points: one call per refresh from the GEOS-native state.
!
! What LW_Driver (GEOS_IrradGridComp.F90:3237-3547) and SORADCORE (GEOS_SolarGridComp.F90:6113-6447)
! do around their RRTMG calls - flip in the vertical, Pa -> hPa, q -> vmr, content -> path, radius
! limits, TLEV, layer heights, negative clean-up; then unflip, sign convention, SFCEM, FSW = down - up,
! clear counts -> cloud fractions, COT ratios - runs on the device inside rrtmgx_irrad_refresh /
! rrtmgx_solar_refresh.  The driver passes its own arrays (IM*JM columns flattened, levels top-down,
! SI units) and gets the INTERNAL / EXPORT arrays back; the *_R work arrays, their allocation and the
! two FLIP timers disappear from the driver.  Source only (no Fortran compiler in the build image).
module rrtmgx_glue
   use, intrinsic :: iso_c_binding
   use rrtmgx_c
   implicit none
   private
   public :: rrtmgx_irrad, rrtmgx_solar

contains

   ! Replaces GEOS_IrradGridComp.F90:3237-3371 + the RRTMG_LW call :3471-3478 + :3486-3547.
   ! CWC / REFF are passed per species (KLIQUID, KICE slices); CO2_3d is optional as in the driver.
   subroutine rrtmgx_irrad(ncol, LM, PLE, PL, T, Q, O3, CH4, N2O, CO2_FIXED, O2, CCL4, CFC11, CFC12, HCFC22, &
                           FCLD, QLIQ, QICE, RLIQ, RICE, TS, T2M, EMIS, LATS, TAUA, SSAA, &
                           ICEFLGLW, LIQFLGLW, DOY, LCLDMH, LCLDLM, BAND_OUTPUT, &
                           AIRMW, H2OMW, O3MW, RGAS, GRAV, &
                           FLXU_INT, FLXD_INT, FLCU_INT, FLCD_INT, DFDTS, DFDTSC, SFCEM_INT, &
                           CLDTTLW, CLDHILW, CLDMDLW, CLDLOLW, OLRBRG, DOLRBRG_DTS, CO2_3d)
      integer, intent(in) :: ncol, LM, ICEFLGLW, LIQFLGLW, DOY, LCLDMH, LCLDLM
      real, intent(in), target :: PLE(ncol,0:LM)
      real, intent(in), target, dimension(ncol,LM) :: PL, T, Q, O3, CH4, N2O, CFC11, CFC12, HCFC22, FCLD
      real, intent(in), target, dimension(ncol,LM) :: QLIQ, QICE, RLIQ, RICE
      real, intent(in), target, dimension(ncol) :: TS, T2M, EMIS, LATS
      real, intent(in), target, dimension(ncol,LM,16) :: TAUA, SSAA
      real, intent(in) :: CO2_FIXED, O2, CCL4, AIRMW, H2OMW, O3MW, RGAS, GRAV
      logical, intent(in) :: BAND_OUTPUT(16)
      real, intent(out), target, dimension(ncol,0:LM) :: FLXU_INT, FLXD_INT, FLCU_INT, FLCD_INT, DFDTS, DFDTSC
      real, intent(out), target, dimension(ncol) :: SFCEM_INT, CLDTTLW, CLDHILW, CLDMDLW, CLDLOLW
      real, intent(inout), target, dimension(16,ncol) :: OLRBRG, DOLRBRG_DTS
      real, intent(in), target, optional :: CO2_3d(ncol,LM)
      type(rrtmgx_irrad_args) :: a
      integer(c_int), target :: bo(16)
      integer(c_int) :: status

      bo = merge(1_c_int, 0_c_int, BAND_OUTPUT)
      a%ncol = ncol; a%lm = LM; a%iceflg = ICEFLGLW; a%liqflg = LIQFLGLW; a%doy = DOY
      a%lcldmh = LCLDMH; a%lcldlm = LCLDLM; a%flags = rrtmgx_real_flags; a%stream = c_null_ptr
      a%co2_fixed = CO2_FIXED; a%o2 = O2; a%ccl4 = CCL4
      a%airmw = AIRMW; a%h2omw = H2OMW; a%o3mw = O3MW; a%rgas = RGAS; a%grav = GRAV
      a%ple = c_loc(PLE); a%pl = c_loc(PL); a%t = c_loc(T); a%q = c_loc(Q); a%o3 = c_loc(O3)
      a%ch4 = c_loc(CH4); a%n2o = c_loc(N2O); a%co2 = c_null_ptr
      if (present(CO2_3d)) a%co2 = c_loc(CO2_3d)
      a%cfc11 = c_loc(CFC11); a%cfc12 = c_loc(CFC12); a%hcfc22 = c_loc(HCFC22); a%fcld = c_loc(FCLD)
      a%qliq = c_loc(QLIQ); a%qice = c_loc(QICE); a%rliq = c_loc(RLIQ); a%rice = c_loc(RICE)
      a%ts = c_loc(TS); a%t2m = c_loc(T2M); a%emis = c_loc(EMIS); a%lats = c_loc(LATS)
      a%taua = c_loc(TAUA); a%ssaa = c_loc(SSAA); a%band_output = c_loc(bo)
      a%flxu = c_loc(FLXU_INT); a%flxd = c_loc(FLXD_INT); a%flcu = c_loc(FLCU_INT); a%flcd = c_loc(FLCD_INT)
      a%dfdts = c_loc(DFDTS); a%dfdtsc = c_loc(DFDTSC); a%sfcem = c_loc(SFCEM_INT)
      a%cldtt = c_loc(CLDTTLW); a%cldhi = c_loc(CLDHILW); a%cldmd = c_loc(CLDMDLW); a%cldlo = c_loc(CLDLOLW)
      a%olrb = c_loc(OLRBRG); a%dolrb_dts = c_loc(DOLRBRG_DTS)
      status = rrtmgx_irrad_refresh(a)
      if (status /= 0) then   ! rrtmg_lw stops on its traps (LW/src/rrtmg_lw_rad.F90:209-318)
         write(*,*) 'rrtmgx_irrad: ', rrtmgx_message(status)
         error stop 'rrtmgx_irrad'
      end if
   end subroutine

   ! Replaces GEOS_SolarGridComp.F90:6113-6223 + the RRTMG_SW call :6331-6387 + :6395-6447.
   subroutine rrtmgx_solar(ncol, LM, SC, DIST, ZT, ISOLVAR, PLE, PL, T, Q, O3, CH4, CO2, O2, CL, &
                           QLIQ, QICE, RLIQ, RICE, TS, LATS, ALBVR, ALBVF, ALBNR, ALBNF, TAUA, SSAA, ASYA, &
                           ICEFLGSW, LIQFLGSW, DOY, LCLDMH, LCLDLM, SOLCYCFRAC, &
                           AIRMW, H2OMW, O3MW, RGAS, GRAV, UNDEF, &
                           FSW, FSC, FSWU, FSCU, NIRR, NIRF, PARR, PARF, UVRR, UVRF, FSWBAND, &
                           CLDTS, CLDHS, CLDMS, CLDLS, COTTP, COTHP, COTMP, COTLP, RC)
      integer, intent(in) :: ncol, LM, ISOLVAR, ICEFLGSW, LIQFLGSW, DOY, LCLDMH, LCLDLM
      real, intent(in) :: SC, DIST, CO2, O2, AIRMW, H2OMW, O3MW, RGAS, GRAV, UNDEF
      real, intent(in), target :: SOLCYCFRAC
      real, intent(in), target :: PLE(ncol,LM+1)
      real, intent(in), target, dimension(ncol,LM) :: PL, T, Q, O3, CH4, CL, QLIQ, QICE, RLIQ, RICE
      real, intent(in), target, dimension(ncol) :: ZT, TS, LATS, ALBVR, ALBVF, ALBNR, ALBNF
      real, intent(in), target, dimension(ncol,LM,14) :: TAUA, SSAA, ASYA   ! un-normalised (SOL:6116-6126)
      real, intent(out), target, dimension(ncol,LM+1) :: FSW, FSC, FSWU, FSCU
      real, intent(out), target, dimension(ncol) :: NIRR, NIRF, PARR, PARF, UVRR, UVRF
      real, intent(out), target :: FSWBAND(ncol,14)
      real, intent(out), target, dimension(ncol) :: CLDTS, CLDHS, CLDMS, CLDLS, COTTP, COTHP, COTMP, COTLP
      integer, intent(out), optional :: RC
      type(rrtmgx_solar_args) :: a
      integer(c_int) :: status
      real(c_double), target :: solcycfrac_d

      a%ncol = ncol; a%lm = LM; a%iceflg = ICEFLGSW; a%liqflg = LIQFLGSW; a%doy = DOY; a%isolvar = ISOLVAR
      a%lcldmh = LCLDMH; a%lcldlm = LCLDLM; a%flags = rrtmgx_real_flags; a%stream = c_null_ptr
      a%sc = SC; a%dist = DIST; a%co2 = CO2; a%o2 = O2
      a%airmw = AIRMW; a%h2omw = H2OMW; a%o3mw = O3MW; a%rgas = RGAS; a%grav = GRAV; a%undef = UNDEF
      solcycfrac_d = SOLCYCFRAC; a%solcycfrac = c_loc(solcycfrac_d)
      a%ple = c_loc(PLE); a%pl = c_loc(PL); a%t = c_loc(T); a%q = c_loc(Q); a%o3 = c_loc(O3); a%ch4 = c_loc(CH4)
      a%cl = c_loc(CL); a%qliq = c_loc(QLIQ); a%qice = c_loc(QICE); a%rliq = c_loc(RLIQ); a%rice = c_loc(RICE)
      a%ts = c_loc(TS); a%zt = c_loc(ZT); a%lats = c_loc(LATS)
      a%albvr = c_loc(ALBVR); a%albvf = c_loc(ALBVF); a%albnr = c_loc(ALBNR); a%albnf = c_loc(ALBNF)
      a%taua = c_loc(TAUA); a%ssaa = c_loc(SSAA); a%asya = c_loc(ASYA)
      a%fsw = c_loc(FSW); a%fsc = c_loc(FSC); a%fswu = c_loc(FSWU); a%fscu = c_loc(FSCU)
      a%nirr = c_loc(NIRR); a%nirf = c_loc(NIRF); a%parr = c_loc(PARR); a%parf = c_loc(PARF)
      a%uvrr = c_loc(UVRR); a%uvrf = c_loc(UVRF); a%fswband = c_loc(FSWBAND)
      a%cldts = c_loc(CLDTS); a%cldhs = c_loc(CLDHS); a%cldms = c_loc(CLDMS); a%cldls = c_loc(CLDLS)
      a%cottp = c_loc(COTTP); a%cothp = c_loc(COTHP); a%cotmp = c_loc(COTMP); a%cotlp = c_loc(COTLP)
      status = rrtmgx_solar_refresh(a)
      if (present(RC)) then   ! rrtmg_sw reports through RC (_ASSERT, SW/src/rrtmg_sw_rad.F90:365-383)
         RC = status
      else if (status /= 0) then
         write(*,*) 'rrtmgx_solar: ', rrtmgx_message(status)
         error stop 'rrtmgx_solar'
      end if
   end subroutine

end module rrtmgx_glue
