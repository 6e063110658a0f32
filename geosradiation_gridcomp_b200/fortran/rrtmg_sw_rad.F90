! Drop-in `module rrtmg_sw_rad`: the reference interface of
! GEOSsolar_GridComp/RRTMG/rrtmg_sw/gcm_model/src/rrtmg_sw_rad.F90:68-124,130-357 (the default
! build, and the SOLAR_RADVAL build when compiled with -DSOLAR_RADVAL like the reference,
! GEOSsolar_GridComp/CMakeLists.txt:18-20) over the B200 library.  GEOS_SolarGridComp.F90:6331-6387 compiles
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
#ifdef SOLAR_RADVAL
      cdsdtp, cdsdhp, cdsdmp, cdsdlp, &
      cdsntp, cdsnhp, cdsnmp, cdsnlp, &
      cotldtp, cotldhp, cotldmp, cotldlp, &
      cotlntp, cotlnhp, cotlnmp, cotlnlp, &
      cdsldtp, cdsldhp, cdsldmp, cdsldlp, &
      cdslntp, cdslnhp, cdslnmp, cdslnlp, &
      cotidtp, cotidhp, cotidmp, cotidlp, &
      cotintp, cotinhp, cotinmp, cotinlp, &
      cdsidtp, cdsidhp, cdsidmp, cdsidlp, &
      cdsintp, cdsinhp, cdsinmp, cdsinlp, &
      ssaldtp, ssaldhp, ssaldmp, ssaldlp, &
      ssalntp, ssalnhp, ssalnmp, ssalnlp, &
      sdsldtp, sdsldhp, sdsldmp, sdsldlp, &
      sdslntp, sdslnhp, sdslnmp, sdslnlp, &
      ssaidtp, ssaidhp, ssaidmp, ssaidlp, &
      ssaintp, ssainhp, ssainmp, ssainlp, &
      sdsidtp, sdsidhp, sdsidmp, sdsidlp, &
      sdsintp, sdsinhp, sdsinmp, sdsinlp, &
      asmldtp, asmldhp, asmldmp, asmldlp, &
      asmlntp, asmlnhp, asmlnmp, asmlnlp, &
      adsldtp, adsldhp, adsldmp, adsldlp, &
      adslntp, adslnhp, adslnmp, adslnlp, &
      asmidtp, asmidhp, asmidmp, asmidlp, &
      asmintp, asminhp, asminmp, asminlp, &
      adsidtp, adsidhp, adsidmp, adsidlp, &
      adsintp, adsinhp, adsinmp, adsinlp, &
      forldtp, forldhp, forldmp, forldlp, &
      forlntp, forlnhp, forlnmp, forlnlp, &
      foridtp, foridhp, foridmp, foridlp, &
      forintp, forinhp, forinmp, forinlp, &
#endif
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
#ifdef SOLAR_RADVAL
      ! the developer-validation diagnostics of the reference's SOLAR_RADVAL build (:85-122, :306-345)
      real, intent(out), dimension(ncol) :: cdsdtp, cdsdhp, cdsdmp, cdsdlp, &
                                            cdsntp, cdsnhp, cdsnmp, cdsnlp
      real, intent(out), dimension(ncol) :: cotldtp, cotldhp, cotldmp, cotldlp, &
                                            cotlntp, cotlnhp, cotlnmp, cotlnlp
      real, intent(out), dimension(ncol) :: cdsldtp, cdsldhp, cdsldmp, cdsldlp, &
                                            cdslntp, cdslnhp, cdslnmp, cdslnlp
      real, intent(out), dimension(ncol) :: cotidtp, cotidhp, cotidmp, cotidlp, &
                                            cotintp, cotinhp, cotinmp, cotinlp
      real, intent(out), dimension(ncol) :: cdsidtp, cdsidhp, cdsidmp, cdsidlp, &
                                            cdsintp, cdsinhp, cdsinmp, cdsinlp
      real, intent(out), dimension(ncol) :: ssaldtp, ssaldhp, ssaldmp, ssaldlp, &
                                            ssalntp, ssalnhp, ssalnmp, ssalnlp
      real, intent(out), dimension(ncol) :: sdsldtp, sdsldhp, sdsldmp, sdsldlp, &
                                            sdslntp, sdslnhp, sdslnmp, sdslnlp
      real, intent(out), dimension(ncol) :: ssaidtp, ssaidhp, ssaidmp, ssaidlp, &
                                            ssaintp, ssainhp, ssainmp, ssainlp
      real, intent(out), dimension(ncol) :: sdsidtp, sdsidhp, sdsidmp, sdsidlp, &
                                            sdsintp, sdsinhp, sdsinmp, sdsinlp
      real, intent(out), dimension(ncol) :: asmldtp, asmldhp, asmldmp, asmldlp, &
                                            asmlntp, asmlnhp, asmlnmp, asmlnlp
      real, intent(out), dimension(ncol) :: adsldtp, adsldhp, adsldmp, adsldlp, &
                                            adslntp, adslnhp, adslnmp, adslnlp
      real, intent(out), dimension(ncol) :: asmidtp, asmidhp, asmidmp, asmidlp, &
                                            asmintp, asminhp, asminmp, asminlp
      real, intent(out), dimension(ncol) :: adsidtp, adsidhp, adsidmp, adsidlp, &
                                            adsintp, adsinhp, adsinmp, adsinlp
      real, intent(out), dimension(ncol) :: forldtp, forldhp, forldmp, forldlp, &
                                            forlntp, forlnhp, forlnmp, forlnlp
      real, intent(out), dimension(ncol) :: foridtp, foridhp, foridmp, foridlp, &
                                            forintp, forinhp, forinmp, forinlp
#endif
      logical, intent(in) :: do_drfband
      real, pointer, dimension(:,:) :: drband, dfband        ! (ncol,14), touched only if do_drfband
      real, intent(in), optional, target :: bndscl(14), indsolvar(2), solcycfrac
      integer, intent(out), optional :: RC

      type(rrtmgx_sw_args) :: a
      integer(c_int) :: status
      real(c_double), target :: bndscl_d(14), indsolvar_d(2), solcycfrac_d   ! always double in the C ABI
#ifdef SOLAR_RADVAL
      real, allocatable, target :: radval(:,:)   ! (ncol,RRTMGX_NRADVAL): the 120 dummies in list order, one column each
#endif

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

#ifdef SOLAR_RADVAL
      allocate(radval(ncol,RRTMGX_NRADVAL))   ! on the heap: ncol is a whole partition of the grid
      a%radval = c_loc(radval)
#else
      a%radval = c_null_ptr
#endif

      status = rrtmgx_sw_run(a)
      ! the reference reports through MAPL's _ASSERT/_FAIL -> RC (rrtmg_sw_rad.F90:365-383,910,1033)
      _ASSERT(status == 0, 'rrtmg_sw (rrtmgx): ' // rrtmgx_message(status))
#ifdef SOLAR_RADVAL
      cdsdtp = radval(:,1)
      cdsdhp = radval(:,2)
      cdsdmp = radval(:,3)
      cdsdlp = radval(:,4)
      cdsntp = radval(:,5)
      cdsnhp = radval(:,6)
      cdsnmp = radval(:,7)
      cdsnlp = radval(:,8)
      cotldtp = radval(:,9)
      cotldhp = radval(:,10)
      cotldmp = radval(:,11)
      cotldlp = radval(:,12)
      cotlntp = radval(:,13)
      cotlnhp = radval(:,14)
      cotlnmp = radval(:,15)
      cotlnlp = radval(:,16)
      cdsldtp = radval(:,17)
      cdsldhp = radval(:,18)
      cdsldmp = radval(:,19)
      cdsldlp = radval(:,20)
      cdslntp = radval(:,21)
      cdslnhp = radval(:,22)
      cdslnmp = radval(:,23)
      cdslnlp = radval(:,24)
      cotidtp = radval(:,25)
      cotidhp = radval(:,26)
      cotidmp = radval(:,27)
      cotidlp = radval(:,28)
      cotintp = radval(:,29)
      cotinhp = radval(:,30)
      cotinmp = radval(:,31)
      cotinlp = radval(:,32)
      cdsidtp = radval(:,33)
      cdsidhp = radval(:,34)
      cdsidmp = radval(:,35)
      cdsidlp = radval(:,36)
      cdsintp = radval(:,37)
      cdsinhp = radval(:,38)
      cdsinmp = radval(:,39)
      cdsinlp = radval(:,40)
      ssaldtp = radval(:,41)
      ssaldhp = radval(:,42)
      ssaldmp = radval(:,43)
      ssaldlp = radval(:,44)
      ssalntp = radval(:,45)
      ssalnhp = radval(:,46)
      ssalnmp = radval(:,47)
      ssalnlp = radval(:,48)
      sdsldtp = radval(:,49)
      sdsldhp = radval(:,50)
      sdsldmp = radval(:,51)
      sdsldlp = radval(:,52)
      sdslntp = radval(:,53)
      sdslnhp = radval(:,54)
      sdslnmp = radval(:,55)
      sdslnlp = radval(:,56)
      ssaidtp = radval(:,57)
      ssaidhp = radval(:,58)
      ssaidmp = radval(:,59)
      ssaidlp = radval(:,60)
      ssaintp = radval(:,61)
      ssainhp = radval(:,62)
      ssainmp = radval(:,63)
      ssainlp = radval(:,64)
      sdsidtp = radval(:,65)
      sdsidhp = radval(:,66)
      sdsidmp = radval(:,67)
      sdsidlp = radval(:,68)
      sdsintp = radval(:,69)
      sdsinhp = radval(:,70)
      sdsinmp = radval(:,71)
      sdsinlp = radval(:,72)
      asmldtp = radval(:,73)
      asmldhp = radval(:,74)
      asmldmp = radval(:,75)
      asmldlp = radval(:,76)
      asmlntp = radval(:,77)
      asmlnhp = radval(:,78)
      asmlnmp = radval(:,79)
      asmlnlp = radval(:,80)
      adsldtp = radval(:,81)
      adsldhp = radval(:,82)
      adsldmp = radval(:,83)
      adsldlp = radval(:,84)
      adslntp = radval(:,85)
      adslnhp = radval(:,86)
      adslnmp = radval(:,87)
      adslnlp = radval(:,88)
      asmidtp = radval(:,89)
      asmidhp = radval(:,90)
      asmidmp = radval(:,91)
      asmidlp = radval(:,92)
      asmintp = radval(:,93)
      asminhp = radval(:,94)
      asminmp = radval(:,95)
      asminlp = radval(:,96)
      adsidtp = radval(:,97)
      adsidhp = radval(:,98)
      adsidmp = radval(:,99)
      adsidlp = radval(:,100)
      adsintp = radval(:,101)
      adsinhp = radval(:,102)
      adsinmp = radval(:,103)
      adsinlp = radval(:,104)
      forldtp = radval(:,105)
      forldhp = radval(:,106)
      forldmp = radval(:,107)
      forldlp = radval(:,108)
      forlntp = radval(:,109)
      forlnhp = radval(:,110)
      forlnmp = radval(:,111)
      forlnlp = radval(:,112)
      foridtp = radval(:,113)
      foridhp = radval(:,114)
      foridmp = radval(:,115)
      foridlp = radval(:,116)
      forintp = radval(:,117)
      forinhp = radval(:,118)
      forinmp = radval(:,119)
      forinlp = radval(:,120)
#endif
      _RETURN(_SUCCESS)
   end subroutine rrtmg_sw

end module rrtmg_sw_rad
