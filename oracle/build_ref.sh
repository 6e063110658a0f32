#!/bin/bash
# oracle/_ref: the UNMODIFIED reference RRTMG LW + SW + McICA (Fortran) compiled from the sources where they lie under
# $REFERENCE_ROOT (default /root/reference), with stand-ins for the two MAPL/ESMF modules and the MAPL_Generic.h
# macros the SW driver uses (oracle/ref_recipe/), plus C-callable wrappers (oracle/ref_recipe/ref_capi.F90).
# Output: oracle/_ref/libgeosref.so (+ .mod files under oracle/_ref/mod); nothing else is written, no reference source
# is copied.  Test infrastructure: tests/golden/make_golden_from_ref.py turns its output into golden vectors that pin
# the C restatement (oracle/*.c) - and through it the CUDA path - by reference OUTPUT.
#
#   oracle/build_ref.sh            fp64 contract: default real promoted to 8 bytes
#   REF_REAL=4 oracle/build_ref.sh production kind: default real = real(4)  -> oracle/_ref/libgeosref_r4.so
#   FC=/path/to/compiler           override the compiler search (gfortran, ifx, ifort, flang, nvfortran)
#   oracle/build_ref.sh --dry-run  print the compile lines, compile nothing
# Exit status: 0 built, 3 no Fortran compiler (the state of this image and of the GPU boxes so far), 4 no reference tree.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
DRY=0; [ "$1" = "--dry-run" ] && DRY=1
if [ -z "$FC" ]; then
  for c in gfortran ifx ifort flang flang-new nvfortran; do
    if command -v "$c" >/dev/null 2>&1; then FC="$c"; break; fi
  done
fi
if [ -z "$FC" ]; then echo "oracle/build_ref.sh: no Fortran compiler found (gfortran ifx ifort flang nvfortran): oracle/_ref not built" >&2; exit 3; fi
if [ ! -d "$REF/GEOSirrad_GridComp" ]; then echo "oracle/build_ref.sh: no reference tree at $REF" >&2; exit 4; fi
LW="$REF/GEOSirrad_GridComp/RRTMG/rrtmg_lw/gcm_model"
SW="$REF/GEOSsolar_GridComp/RRTMG/rrtmg_sw/gcm_model"
SH="$REF/GEOS_RadiationShared"
RK="${REF_REAL:-8}"
case "$(basename "$FC")" in
  gfortran*) FLAGS="-O2 -fPIC -cpp -ffree-line-length-none -ffp-contract=off -fno-fast-math -J$OUT/mod"; [ "$RK" = 8 ] && FLAGS="$FLAGS -fdefault-real-8 -fdefault-double-8" ;;
  ifx*|ifort*) FLAGS="-O2 -fPIC -fpp -fp-model=strict -no-fma -module $OUT/mod"; [ "$RK" = 8 ] && FLAGS="$FLAGS -r8" ;;
  nvfortran*) FLAGS="-O2 -fPIC -Mpreprocess -Kieee -Mnofma -module $OUT/mod"; [ "$RK" = 8 ] && FLAGS="$FLAGS -r8" ;;
  *) FLAGS="-O2 -fPIC -cpp -ffp-contract=off -J$OUT/mod"; [ "$RK" = 8 ] && FLAGS="$FLAGS -fdefault-real-8" ;;
esac
LIB="$OUT/libgeosref.so"; [ "$RK" = 4 ] && LIB="$OUT/libgeosref_r4.so"
# compile order = module dependency order
SRCS=("$HERE/ref_recipe/mapl_stub.F90"
      "$SH/cloud_condensate_inhomogeneity.F90" "$SH/cloud_subcol_gen.F90"
      "$LW/modules/parrrtm.F90" "$LW/modules/rrlw_cld.F90" "$LW/modules/rrlw_con.F90")
for k in 01 02 03 04 05 06 07 08 09 10 11 12 13 14 15 16; do SRCS+=("$LW/modules/rrlw_kg$k.F90"); done
SRCS+=("$LW/modules/rrlw_ncpar.F90" "$LW/modules/rrlw_ref.F90" "$LW/modules/rrlw_tbl.F90" "$LW/modules/rrlw_vsn.F90" "$LW/modules/rrlw_wvn.F90")
for k in 01 02 03 04 05 06 07 08 09 10 11 12 13 14 15 16; do SRCS+=("$LW/src/rrtmg_lw_k_g_$k.F90"); done
SRCS+=("$LW/src/rrtmg_lw_init.F90" "$LW/src/rrtmg_lw_cldprmc.F90" "$LW/src/rrtmg_lw_setcoef.F90" "$LW/src/rrtmg_lw_taumol.F90"
       "$LW/src/rrtmg_lw_rtrnmc.F90" "$LW/src/rrtmg_lw_rad.F90"
       "$SW/modules/parrrsw.F90" "$SW/modules/rrsw_aer.F90" "$SW/modules/rrsw_cld.F90" "$SW/modules/rrsw_con.F90")
for k in 16 17 18 19 20 21 22 23 24 25 26 27 28 29; do SRCS+=("$SW/modules/rrsw_kg$k.F90"); done
SRCS+=("$SW/modules/rrsw_ref.F90" "$SW/modules/rrsw_tbl.F90" "$SW/modules/rrsw_vsn.F90" "$SW/modules/rrsw_wvn.F90")
for k in 16 17 18 19 20 21 22 23 24 25 26 27 28 29; do SRCS+=("$SW/src/rrtmg_sw_k_g_$k.F90"); done
SRCS+=("$SW/src/NRLSSI2.F90" "$SW/src/rrtmg_sw_init.F90" "$SW/src/rrtmg_sw_cldprmc.F90" "$SW/src/rrtmg_sw_setcoef.F90"
       "$SW/src/rrtmg_sw_taumol.F90" "$SW/src/rrtmg_sw_spcvmc.F90" "$SW/src/rrtmg_sw_rad.F90"
       "$HERE/ref_recipe/ref_capi.F90")
for f in "${SRCS[@]}"; do
  if [ ! -f "$f" ]; then echo "oracle/build_ref.sh: missing source $f" >&2; exit 4; fi
done
echo "oracle/build_ref.sh: $FC, default real = $RK bytes, ${#SRCS[@]} sources -> $LIB"
if [ "$DRY" = 1 ]; then
  for f in "${SRCS[@]}"; do echo "$FC $FLAGS -I$HERE/ref_recipe -c $f -o $OUT/obj/$(basename "${f%.F90}").o"; done
  echo "$FC -shared -o $LIB $OUT/obj/*.o"
  exit 0
fi
mkdir -p "$OUT/mod" "$OUT/obj"
# modules must be compiled before their users: the list above is in dependency order as read from the sources, and
# the loop below retries what failed until a pass makes no progress, so a mistake in the order is not fatal
TODO=("${SRCS[@]}")
OBJS=()
while [ ${#TODO[@]} -gt 0 ]; do
  NEXT=()
  for f in "${TODO[@]}"; do
    o="$OUT/obj/$(basename "${f%.F90}").o"
    if $FC $FLAGS -I"$HERE/ref_recipe" -c "$f" -o "$o" 2> "$OUT/obj/last_error.txt"; then OBJS+=("$o"); else NEXT+=("$f"); fi
  done
  if [ ${#NEXT[@]} -eq ${#TODO[@]} ]; then
    echo "oracle/build_ref.sh: ${#NEXT[@]} sources do not compile; first: ${NEXT[0]}" >&2
    $FC $FLAGS -I"$HERE/ref_recipe" -c "${NEXT[0]}" -o /dev/null >&2 || true
    exit 5
  fi
  TODO=("${NEXT[@]}")
done
$FC -shared -o "$LIB" "${OBJS[@]}"
echo "built $LIB"
