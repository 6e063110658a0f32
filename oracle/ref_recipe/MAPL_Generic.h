! Minimal stand-in for MAPL's error-handling macros, enough to compile
! GEOSsolar_GridComp/RRTMG/rrtmg_sw/gcm_model/src/rrtmg_sw_rad.F90 and rrtmg_sw_spcvmc.F90 from the reference tree
! WITHOUT MAPL/ESMF (oracle/build_ref.sh).  Test infrastructure only: it turns _ASSERT/_FAIL into an early return
! with RC = a negative line-independent code so that the golden-vector generator can see which trap fired.
#define _SUCCESS 0
#define _FAILURE 1
#define _VERIFY(A) if ((A) /= 0) then; if (present(RC)) RC = (A); return; endif
#define __RC__ RC=STATUS); _VERIFY(STATUS
#define _RC __RC__
#define _ASSERT(A,msg) if (.not.(A)) then; if (present(RC)) RC = ref_trap(msg); return; endif
#define _FAIL(msg) if (present(RC)) RC = ref_trap(msg); return
#define _RETURN(A) if (present(RC)) RC = A; return
