! Drop-in `module rrtmg_lw_rad`: the reference interface of
! GEOSirrad_GridComp/RRTMG/rrtmg_lw/gcm_model/src/rrtmg_lw_rad.F90:15-23,113-201 over the
! B200 library.  GEOS_IrradGridComp.F90:3439,3471 compile against it unchanged.
module rrtmg_lw_rad
   use, intrinsic :: iso_c_binding
   use, intrinsic :: iso_fortran_env, only : error_unit
   use rrtmgx_c
   implicit none
   private
   public :: rrtmg_lw
contains

   subroutine rrtmg_lw( &
      ncol, nlay, psize, dudTs, &
      play, plev, tlay, tlev, tsfc, emis, &
      h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, &
      cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, &
      cldf, ciwp, clwp, rei, rel, iceflglw, liqflglw, &
      tauaer, zm, alat, dyofyr, cloudLM, cloudMH, clearCounts, &
      uflx, dflx, uflxc, dflxc, duflx_dTs, duflxc_dTs, &
      band_output, olrb, dolrb_dTs)

      integer, intent(in) :: ncol, nlay, psize
      logical, intent(in) :: dudTs
      real, intent(in), target :: play(ncol,nlay), plev(ncol,0:nlay), tlay(ncol,nlay), tlev(ncol,0:nlay)
      real, intent(in), target :: tsfc(ncol), emis(ncol,16)
      real, intent(in), target, dimension(ncol,nlay) :: h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr
      real, intent(in), target, dimension(ncol,nlay) :: cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr
      real, intent(in), target, dimension(ncol,nlay) :: cldf, ciwp, clwp, rei, rel
      integer, intent(in) :: iceflglw, liqflglw
      real, intent(in), target :: tauaer(ncol,nlay,16), zm(ncol,nlay), alat(ncol)
      integer, intent(in) :: dyofyr, cloudLM, cloudMH
      integer, intent(out), target :: clearCounts(ncol,4)
      real, intent(out), target, dimension(ncol,nlay+1) :: uflx, dflx, uflxc, dflxc, duflx_dTs, duflxc_dTs
      logical, intent(in) :: band_output(16)
      real, intent(out), target :: olrb(16,ncol), dolrb_dTs(16,ncol)

      type(rrtmgx_lw_args) :: a
      integer(c_int), target :: bo(16)
      integer(c_int) :: status

      bo = merge(1_c_int, 0_c_int, band_output)          ! logical -> int32
      a%ncol = ncol; a%nlay = nlay; a%psize = psize      ! psize: cache blocking of the CPU code, ignored
      a%dudTs = merge(1_c_int, 0_c_int, dudTs)
      a%iceflglw = iceflglw; a%liqflglw = liqflglw
      a%dyofyr = dyofyr; a%cloudLM = cloudLM; a%cloudMH = cloudMH
      a%flags = rrtmgx_real_flags                        ! host arrays of default real; the library stages them
      a%stream = c_null_ptr
      a%play = c_loc(play); a%plev = c_loc(plev); a%tlay = c_loc(tlay); a%tlev = c_loc(tlev)
      a%tsfc = c_loc(tsfc); a%emis = c_loc(emis)
      a%h2ovmr = c_loc(h2ovmr); a%o3vmr = c_loc(o3vmr); a%co2vmr = c_loc(co2vmr); a%ch4vmr = c_loc(ch4vmr)
      a%n2ovmr = c_loc(n2ovmr); a%o2vmr = c_loc(o2vmr)
      a%cfc11vmr = c_loc(cfc11vmr); a%cfc12vmr = c_loc(cfc12vmr); a%cfc22vmr = c_loc(cfc22vmr)
      a%ccl4vmr = c_loc(ccl4vmr)
      a%cldf = c_loc(cldf); a%ciwp = c_loc(ciwp); a%clwp = c_loc(clwp); a%rei = c_loc(rei); a%rel = c_loc(rel)
      a%tauaer = c_loc(tauaer); a%zm = c_loc(zm); a%alat = c_loc(alat)
      a%band_output = c_loc(bo)
      a%clearCounts = c_loc(clearCounts)
      a%uflx = c_loc(uflx); a%dflx = c_loc(dflx); a%uflxc = c_loc(uflxc); a%dflxc = c_loc(dflxc)
      a%duflx_dTs = c_loc(duflx_dTs); a%duflxc_dTs = c_loc(duflxc_dTs)
      a%olrb = c_loc(olrb); a%dolrb_dTs = c_loc(dolrb_dTs)

      status = rrtmgx_lw_run(a)
      if (status /= 0) then                              ! the reference `error stop`s on every trap
         write(error_unit,*) 'file:', __FILE__, ', line:', __LINE__
         write(error_unit,*) 'rrtmg_lw: ', rrtmgx_message(status)
         error stop 'RRTMG_LW (rrtmgx) failed'
      end if
   end subroutine rrtmg_lw

end module rrtmg_lw_rad
