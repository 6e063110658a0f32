! Drop-in `module rrtmg_sw_rad`: the reference interface of
! GEOSsolar_GridComp/RRTMG/rrtmg_sw/gcm_model/src/rrtmg_sw_rad.F90:68-124,130-357 (default
! build, SOLAR_RADVAL off) over the B200 library.  GEOS_SolarGridComp.F90:6331-6387 compiles
! against it unchanged; the MAPL handle (timers/asserts only in the reference) is accepted and
! not used.
#include "MAPL_Generic.h"
module rrtmg_sw_rad
   use, intrinsic :: iso_c_binding
   use ESMF
   use MAPL
   use rrtmgx_c
   implicit none
   private
   public :: rrtmg_sw
contains

   subroutine rrtmg_sw(MAPL, &
      rpart, ncol, nlay, &
      scon, adjes, coszen, isolvar, &
      play, plev, tlay, &
      h2ovmr, o3vmr, co2vmr, ch4vmr, o2vmr, &
      iceflgsw, liqflgsw, &
      cld, ciwp, clwp, rei, rel, &
      dyofyr, zm, alat, &
      iaer, tauaer, ssaaer, asmaer, &
      asdir, asdif, aldir, aldif, &
      cloudLM, cloudMH, normFlx, &
      clearCounts, swuflx, swdflx, swuflxc, swdflxc, &
      nirr, nirf, parr, parf, uvrr, uvrf, fswband, &
      cotdtp, cotdhp, cotdmp, cotdlp, &
      cotntp, cotnhp, cotnmp, cotnlp, &
      do_drfband, drband, dfband, &
      bndscl, indsolvar, solcycfrac, &
      RC)

      type(MAPL_MetaComp), pointer, intent(inout) :: MAPL
      integer, intent(in) :: rpart, ncol, nlay
      real, intent(in) :: scon, adjes
      real, intent(in), target :: coszen(ncol)
      integer, intent(in) :: isolvar
      real, intent(in), target :: play(ncol,nlay), plev(ncol,nlay+1), tlay(ncol,nlay)
      real, intent(in), target, dimension(ncol,nlay) :: h2ovmr, o3vmr, co2vmr, ch4vmr, o2vmr
      integer, intent(in) :: iceflgsw, liqflgsw
      real, intent(in), target, dimension(ncol,nlay) :: cld, ciwp, clwp, rei, rel
      integer, intent(in) :: dyofyr
      real, intent(in), target :: zm(ncol,nlay), alat(ncol)
      integer, intent(in) :: iaer
      real, intent(in), target, dimension(ncol,nlay,14) :: tauaer, ssaaer, asmaer
      real, intent(in), target, dimension(ncol) :: asdir, asdif, aldir, aldif
      integer, intent(in) :: cloudLM, cloudMH, normFlx
      integer, intent(out), target :: clearCounts(ncol,4)
      real, intent(out), target, dimension(ncol,nlay+1) :: swuflx, swdflx, swuflxc, swdflxc
      real, intent(out), target, dimension(ncol) :: nirr, nirf, parr, parf, uvrr, uvrf
      real, intent(out), target :: fswband(ncol,14)
      real, intent(out), target, dimension(ncol) :: cotdtp, cotdhp, cotdmp, cotdlp, cotntp, cotnhp, cotnmp, cotnlp
      logical, intent(in) :: do_drfband
      real, pointer, dimension(:,:) :: drband, dfband        ! (ncol,14), touched only if do_drfband
      real, intent(in), optional, target :: bndscl(14), indsolvar(2), solcycfrac
      integer, intent(out), optional :: RC

      type(rrtmgx_sw_args) :: a
      integer(c_int) :: status
      real(c_double), target :: bndscl_d(14), indsolvar_d(2), solcycfrac_d   ! always double in the C ABI

      a%ncol = ncol; a%nlay = nlay; a%rpart = rpart      ! rpart: cache blocking of the CPU code, ignored
      a%isolvar = isolvar; a%iceflgsw = iceflgsw; a%liqflgsw = liqflgsw
      a%dyofyr = dyofyr; a%cloudLM = cloudLM; a%cloudMH = cloudMH
      a%iaer = iaer; a%normFlx = normFlx
      a%do_drfband = merge(1_c_int, 0_c_int, do_drfband)
      a%flags = rrtmgx_real_flags
      a%stream = c_null_ptr
      a%scon = scon; a%adjes = adjes
      a%bndscl = c_null_ptr; a%indsolvar = c_null_ptr; a%solcycfrac = c_null_ptr   ! absent optionals -> NULL
      if (present(bndscl)) then
         bndscl_d = bndscl; a%bndscl = c_loc(bndscl_d)
      end if
      if (present(indsolvar)) then
         indsolvar_d = indsolvar; a%indsolvar = c_loc(indsolvar_d)
      end if
      if (present(solcycfrac)) then
         solcycfrac_d = solcycfrac; a%solcycfrac = c_loc(solcycfrac_d)
      end if
      a%coszen = c_loc(coszen); a%play = c_loc(play); a%plev = c_loc(plev); a%tlay = c_loc(tlay)
      a%h2ovmr = c_loc(h2ovmr); a%o3vmr = c_loc(o3vmr); a%co2vmr = c_loc(co2vmr); a%ch4vmr = c_loc(ch4vmr)
      a%o2vmr = c_loc(o2vmr)
      a%cld = c_loc(cld); a%ciwp = c_loc(ciwp); a%clwp = c_loc(clwp); a%rei = c_loc(rei); a%rel = c_loc(rel)
      a%zm = c_loc(zm); a%alat = c_loc(alat)
      a%tauaer = c_loc(tauaer); a%ssaaer = c_loc(ssaaer); a%asmaer = c_loc(asmaer)
      a%asdir = c_loc(asdir); a%asdif = c_loc(asdif); a%aldir = c_loc(aldir); a%aldif = c_loc(aldif)
      a%clearCounts = c_loc(clearCounts)
      a%swuflx = c_loc(swuflx); a%swdflx = c_loc(swdflx); a%swuflxc = c_loc(swuflxc); a%swdflxc = c_loc(swdflxc)
      a%nirr = c_loc(nirr); a%nirf = c_loc(nirf); a%parr = c_loc(parr); a%parf = c_loc(parf)
      a%uvrr = c_loc(uvrr); a%uvrf = c_loc(uvrf); a%fswband = c_loc(fswband)
      a%cotdtp = c_loc(cotdtp); a%cotdhp = c_loc(cotdhp); a%cotdmp = c_loc(cotdmp); a%cotdlp = c_loc(cotdlp)
      a%cotntp = c_loc(cotntp); a%cotnhp = c_loc(cotnhp); a%cotnmp = c_loc(cotnmp); a%cotnlp = c_loc(cotnlp)
      a%drband = c_null_ptr; a%dfband = c_null_ptr
      if (do_drfband) then
         a%drband = c_loc(drband); a%dfband = c_loc(dfband)
      end if

      status = rrtmgx_sw_run(a)
      ! the reference reports through MAPL's _ASSERT/_FAIL -> RC (rrtmg_sw_rad.F90:365-383,910,1033)
      _ASSERT(status == 0, 'rrtmg_sw (rrtmgx): ' // rrtmgx_message(status))
      _RETURN(_SUCCESS)
   end subroutine rrtmg_sw

end module rrtmg_sw_rad
