! C-callable entry points over the UNMODIFIED reference routines, compiled by oracle/build_ref.sh from the sources
! where they lie under $REFERENCE_ROOT (never copied into this repository).  Test infrastructure: the golden-vector
! generator tests/golden/make_golden_from_ref.py calls these through ctypes to pin the C restatement (oracle/*.c)
! by reference OUTPUT.  `real` is the default real kind, i.e. whatever the build promotes it to
! (-fdefault-real-8 for the fp64 contract, plain real(4) for a production-kind run); ref_real_bytes() tells the caller.
module ref_capi
   use, intrinsic :: iso_c_binding
   use MAPL, only : MAPL_MetaComp
   use rrtmg_lw_init, only : rrtmg_lw_ini
   use rrtmg_sw_init, only : rrtmg_sw_ini
   use rrtmg_lw_rad, only : rrtmg_lw
   use rrtmg_sw_rad, only : rrtmg_sw
   use cloud_condensate_inhomogeneity, only : set_inhomogeneity, unset_inhomogeneity
   use cloud_subcol_gen, only : initialize_cloud_subcol_gen
   implicit none
contains

   integer(c_int) function ref_real_bytes() bind(C, name='ref_real_bytes')
      real :: x
      ref_real_bytes = int(storage_size(x) / 8, c_int)
   end function ref_real_bytes

   ! rrtmg_lw_ini + rrtmg_sw_ini + the McICA module state (RAD:565-578: inhomogeneity option ih, default lengths)
   subroutine ref_init(ih) bind(C, name='ref_init')
      integer(c_int), value :: ih
      call rrtmg_lw_ini
      call rrtmg_sw_ini
      call unset_inhomogeneity
      if (ih > 0) call set_inhomogeneity(int(ih))
   end subroutine ref_init

   subroutine ref_initialize_cloud_subcol_gen(am) bind(C, name='ref_initialize_cloud_subcol_gen')
      real, intent(in) :: am(8)
      call initialize_cloud_subcol_gen(am(1), am(2), am(3), am(4), am(5), am(6), am(7), am(8))
   end subroutine ref_initialize_cloud_subcol_gen

   ! LW/src/rrtmg_lw_rad.F90:15-23
   subroutine ref_rrtmg_lw(ncol, nlay, psize, dudTs, play, plev, tlay, tlev, tsfc, emis, h2ovmr, o3vmr, co2vmr, &
         ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, cldf, ciwp, clwp, rei, rel, iceflglw, &
         liqflglw, tauaer, zm, alat, dyofyr, cloudLM, cloudMH, clearCounts, uflx, dflx, uflxc, dflxc, duflx_dTs, &
         duflxc_dTs, band_output, olrb, dolrb_dTs) bind(C, name='ref_rrtmg_lw')
      integer(c_int), value :: ncol, nlay, psize, dudTs, iceflglw, liqflglw, dyofyr, cloudLM, cloudMH
      real, intent(in) :: play(ncol,nlay), plev(ncol,0:nlay), tlay(ncol,nlay), tlev(ncol,0:nlay), tsfc(ncol), emis(ncol,16)
      real, intent(in), dimension(ncol,nlay) :: h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr, &
         cfc22vmr, ccl4vmr, cldf, ciwp, clwp, rei, rel
      real, intent(in) :: tauaer(ncol,nlay,16), zm(ncol,nlay), alat(ncol)
      integer(c_int), intent(out) :: clearCounts(ncol,4)
      real, intent(out), dimension(ncol,nlay+1) :: uflx, dflx, uflxc, dflxc, duflx_dTs, duflxc_dTs
      integer(c_int), intent(in) :: band_output(16)
      real, intent(out) :: olrb(16,ncol), dolrb_dTs(16,ncol)
      logical :: bo(16)
      integer :: cc(ncol,4)
      bo = band_output /= 0
      call rrtmg_lw(int(ncol), int(nlay), int(psize), dudTs /= 0, play, plev, tlay, tlev, tsfc, emis, h2ovmr, o3vmr, &
         co2vmr, ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, cldf, ciwp, clwp, rei, rel, &
         int(iceflglw), int(liqflglw), tauaer, zm, alat, int(dyofyr), int(cloudLM), int(cloudMH), cc, uflx, dflx, &
         uflxc, dflxc, duflx_dTs, duflxc_dTs, bo, olrb, dolrb_dTs)
      clearCounts = int(cc, c_int)
   end subroutine ref_rrtmg_lw

   ! SW/src/rrtmg_sw_rad.F90:68-124 (no SOLAR_RADVAL).  have_opt: bit 0 bndscl, bit 1 indsolvar, bit 2 solcycfrac.
   ! Returns RC (0, or the stand-in's trap code, oracle/ref_recipe/mapl_stub.F90).
   integer(c_int) function ref_rrtmg_sw(rpart, ncol, nlay, scon, adjes, coszen, isolvar, play, plev, tlay, h2ovmr, &
         o3vmr, co2vmr, ch4vmr, o2vmr, iceflgsw, liqflgsw, cld, ciwp, clwp, rei, rel, dyofyr, zm, alat, iaer, tauaer, &
         ssaaer, asmaer, asdir, asdif, aldir, aldif, cloudLM, cloudMH, normFlx, clearCounts, swuflx, swdflx, swuflxc, &
         swdflxc, nirr, nirf, parr, parf, uvrr, uvrf, fswband, cot, do_drfband, drband, dfband, have_opt, bndscl, &
         indsolvar, solcycfrac) bind(C, name='ref_rrtmg_sw') result(rc_out)
      integer(c_int), value :: rpart, ncol, nlay, isolvar, iceflgsw, liqflgsw, dyofyr, iaer, cloudLM, cloudMH, normFlx, &
         do_drfband, have_opt
      real, value :: scon, adjes, solcycfrac
      real, intent(in) :: coszen(ncol), play(ncol,nlay), plev(ncol,nlay+1), tlay(ncol,nlay)
      real, intent(in), dimension(ncol,nlay) :: h2ovmr, o3vmr, co2vmr, ch4vmr, o2vmr, cld, ciwp, clwp, rei, rel, zm
      real, intent(in) :: alat(ncol), tauaer(ncol,nlay,14), ssaaer(ncol,nlay,14), asmaer(ncol,nlay,14)
      real, intent(in), dimension(ncol) :: asdir, asdif, aldir, aldif
      integer(c_int), intent(out) :: clearCounts(ncol,4)
      real, intent(out), dimension(ncol,nlay+1) :: swuflx, swdflx, swuflxc, swdflxc
      real, intent(out), dimension(ncol) :: nirr, nirf, parr, parf, uvrr, uvrf
      real, intent(out) :: fswband(ncol,14), cot(ncol,8)
      real, intent(out), target :: drband(ncol,14), dfband(ncol,14)
      real, intent(in) :: bndscl(14), indsolvar(2)
      type(MAPL_MetaComp), pointer :: M
      real, pointer :: pdr(:,:), pdf(:,:)
      integer :: cc(ncol,4), rc
      allocate(M)
      pdr => drband
      pdf => dfband
      rc = 0
      select case (iand(int(have_opt), 7))
      case (0)
         call rrtmg_sw(M, int(rpart), int(ncol), int(nlay), scon, adjes, coszen, int(isolvar), play, plev, tlay, h2ovmr, &
            o3vmr, co2vmr, ch4vmr, o2vmr, int(iceflgsw), int(liqflgsw), cld, ciwp, clwp, rei, rel, int(dyofyr), zm, alat, &
            int(iaer), tauaer, ssaaer, asmaer, asdir, asdif, aldir, aldif, int(cloudLM), int(cloudMH), int(normFlx), cc, &
            swuflx, swdflx, swuflxc, swdflxc, nirr, nirf, parr, parf, uvrr, uvrf, fswband, cot(:,1), cot(:,2), cot(:,3), &
            cot(:,4), cot(:,5), cot(:,6), cot(:,7), cot(:,8), do_drfband /= 0, pdr, pdf, RC=rc)
      case (1)
         call rrtmg_sw(M, int(rpart), int(ncol), int(nlay), scon, adjes, coszen, int(isolvar), play, plev, tlay, h2ovmr, &
            o3vmr, co2vmr, ch4vmr, o2vmr, int(iceflgsw), int(liqflgsw), cld, ciwp, clwp, rei, rel, int(dyofyr), zm, alat, &
            int(iaer), tauaer, ssaaer, asmaer, asdir, asdif, aldir, aldif, int(cloudLM), int(cloudMH), int(normFlx), cc, &
            swuflx, swdflx, swuflxc, swdflxc, nirr, nirf, parr, parf, uvrr, uvrf, fswband, cot(:,1), cot(:,2), cot(:,3), &
            cot(:,4), cot(:,5), cot(:,6), cot(:,7), cot(:,8), do_drfband /= 0, pdr, pdf, bndscl=bndscl, RC=rc)
      case (2, 3)
         call rrtmg_sw(M, int(rpart), int(ncol), int(nlay), scon, adjes, coszen, int(isolvar), play, plev, tlay, h2ovmr, &
            o3vmr, co2vmr, ch4vmr, o2vmr, int(iceflgsw), int(liqflgsw), cld, ciwp, clwp, rei, rel, int(dyofyr), zm, alat, &
            int(iaer), tauaer, ssaaer, asmaer, asdir, asdif, aldir, aldif, int(cloudLM), int(cloudMH), int(normFlx), cc, &
            swuflx, swdflx, swuflxc, swdflxc, nirr, nirf, parr, parf, uvrr, uvrf, fswband, cot(:,1), cot(:,2), cot(:,3), &
            cot(:,4), cot(:,5), cot(:,6), cot(:,7), cot(:,8), do_drfband /= 0, pdr, pdf, bndscl=bndscl, &
            indsolvar=indsolvar, RC=rc)
      case default
         call rrtmg_sw(M, int(rpart), int(ncol), int(nlay), scon, adjes, coszen, int(isolvar), play, plev, tlay, h2ovmr, &
            o3vmr, co2vmr, ch4vmr, o2vmr, int(iceflgsw), int(liqflgsw), cld, ciwp, clwp, rei, rel, int(dyofyr), zm, alat, &
            int(iaer), tauaer, ssaaer, asmaer, asdir, asdif, aldir, aldif, int(cloudLM), int(cloudMH), int(normFlx), cc, &
            swuflx, swdflx, swuflxc, swdflxc, nirr, nirf, parr, parf, uvrr, uvrf, fswband, cot(:,1), cot(:,2), cot(:,3), &
            cot(:,4), cot(:,5), cot(:,6), cot(:,7), cot(:,8), do_drfband /= 0, pdr, pdf, bndscl=bndscl, &
            indsolvar=indsolvar, solcycfrac=solcycfrac, RC=rc)
      end select
      clearCounts = int(cc, c_int)
      deallocate(M)
      rc_out = int(rc, c_int)
   end function ref_rrtmg_sw

end module ref_capi
