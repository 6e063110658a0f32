! Stand-ins for the two modules rrtmg_sw_rad.F90 / rrtmg_sw_spcvmc.F90 `use` (ESMF, MAPL): the only MAPL entities the
! RRTMG SW sources touch are the MAPL_MetaComp handle they pass around and the MAPL_TimerOn/Off profiling calls
! (SW/src/rrtmg_sw_rad.F90:130,1181-1200; rrtmg_sw_spcvmc.F90:382-567).  Test infrastructure (oracle/build_ref.sh).
module ESMF
   implicit none
end module ESMF

module MAPL
   implicit none
   type MAPL_MetaComp
      integer :: unused = 0
   end type MAPL_MetaComp
contains
   subroutine MAPL_TimerOn(M, name, RC)
      type(MAPL_MetaComp), pointer, intent(inout) :: M
      character(len=*), intent(in) :: name
      integer, optional, intent(out) :: RC
      if (present(RC)) RC = 0
   end subroutine MAPL_TimerOn
   subroutine MAPL_TimerOff(M, name, RC)
      type(MAPL_MetaComp), pointer, intent(inout) :: M
      character(len=*), intent(in) :: name
      integer, optional, intent(out) :: RC
      if (present(RC)) RC = 0
   end subroutine MAPL_TimerOff
   ! what the _ASSERT / _FAIL stand-ins return: -(100 + k) for the k-th "negative values in input" assertion in the
   ! reference's order (the library's RRTMGX_ENEGATIVE codes), -1 otherwise
   integer function ref_trap(msg)
      character(len=*), intent(in) :: msg
      integer, save :: dummy = 0
      character(len=6), parameter :: names(19) = [character(len=6) :: 'play','plev','tlay','h2ovmr','o3vmr','co2vmr', &
         'ch4vmr','o2vmr','asdir','aldir','asdif','aldif','cld','ciwp','clwp','rei','rel','tauaer','ssaaer']
      integer :: k
      ref_trap = -1
      if (index(msg, 'negative values in input') == 0) return
      do k = 1, 19
         if (index(msg, ': '//trim(names(k))) > 0 .or. index(msg, ' '//trim(names(k))) == len_trim(msg) - len_trim(names(k))) then
            ref_trap = -(100 + k)
            return
         end if
      end do
   end function ref_trap
end module MAPL
