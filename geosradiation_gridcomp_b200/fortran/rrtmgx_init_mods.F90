! Drop-in initialisation modules: the reference entry points called from the Initialize / Run
! phases (GEOS_IrradGridComp.F90:3381, GEOS_SolarGridComp.F90:6225,
! GEOS_RadiationGridComp.F90:565,578) mapped onto rrtmgx_init / rrtmgx_set_mcica.  The k-table
! reduction and upload happen once, in the first call; later calls are no-ops (GEOS calls the
! _ini routines on every refresh).
module rrtmg_lw_init
   use rrtmgx_c
   implicit none
   private
   public :: rrtmg_lw_ini
contains
   subroutine rrtmg_lw_ini()            ! LW/src/rrtmg_lw_init.F90:22
      type(rrtmgx_config) :: cfg
      if (rrtmgx_init(cfg) /= 0) error stop 'rrtmg_lw_ini: rrtmgx_init failed (no CUDA device or table blob)'
   end subroutine
end module rrtmg_lw_init

module rrtmg_sw_init
   use rrtmgx_c
   implicit none
   private
   public :: rrtmg_sw_ini
contains
   subroutine rrtmg_sw_ini()            ! SW/src/rrtmg_sw_init.F90:49
      type(rrtmgx_config) :: cfg
      if (rrtmgx_init(cfg) /= 0) error stop 'rrtmg_sw_ini: rrtmgx_init failed (no CUDA device or table blob)'
   end subroutine
end module rrtmg_sw_init

module cloud_condensate_inhomogeneity
   use, intrinsic :: iso_c_binding
   use rrtmgx_c
   implicit none
   private
   public :: set_inhomogeneity, rrtmgx_mcica_state
   integer(c_int), save :: ih_now = 1
   real(c_double), save :: corr_now(8) = [1.4315d0, 2.1219d0, 7.d0, -25.584d0, 0.72192d0, 0.78996d0, 8.5d0, 40.404d0]
contains
   subroutine set_inhomogeneity(ih)     ! SH/cloud_condensate_inhomogeneity.F90:45
      integer, intent(in) :: ih
      type(rrtmgx_config) :: cfg
      ih_now = ih
      if (rrtmgx_init(cfg) /= 0) error stop 'set_inhomogeneity: rrtmgx_init failed'
      if (rrtmgx_set_mcica(ih_now, corr_now) /= 0) error stop 'set_inhomogeneity: unknown inhomogeneity type'
   end subroutine
   subroutine rrtmgx_mcica_state(ih, corr, set)
      integer(c_int), intent(inout) :: ih
      real(c_double), intent(inout) :: corr(8)
      logical, intent(in) :: set
      if (set) then
         corr_now = corr
      else
         ih = ih_now; corr = corr_now
      end if
   end subroutine
end module cloud_condensate_inhomogeneity

module cloud_subcol_gen
   use, intrinsic :: iso_c_binding
   use rrtmgx_c
   use cloud_condensate_inhomogeneity, only : rrtmgx_mcica_state
   implicit none
   private
   public :: initialize_cloud_subcol_gen
contains
   subroutine initialize_cloud_subcol_gen(adl_am1, adl_am2, adl_am30, adl_am4, &
                                          rdl_am1, rdl_am2, rdl_am30, rdl_am4)   ! SH/cloud_subcol_gen.F90:108
      real, intent(in) :: adl_am1, adl_am2, adl_am30, adl_am4, rdl_am1, rdl_am2, rdl_am30, rdl_am4
      type(rrtmgx_config) :: cfg
      integer(c_int) :: ih
      real(c_double) :: corr(8), dummy(8)
      corr = [adl_am1, adl_am2, adl_am30, adl_am4, rdl_am1, rdl_am2, rdl_am30, rdl_am4]
      call rrtmgx_mcica_state(ih, corr, .true.)
      call rrtmgx_mcica_state(ih, dummy, .false.)
      if (rrtmgx_init(cfg) /= 0) error stop 'initialize_cloud_subcol_gen: rrtmgx_init failed'
      if (rrtmgx_set_mcica(ih, corr) /= 0) error stop 'initialize_cloud_subcol_gen: rrtmgx_set_mcica failed'
   end subroutine
end module cloud_subcol_gen
