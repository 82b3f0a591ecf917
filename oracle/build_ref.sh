#!/bin/bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
# Compiles the reference's own CPU hot path from the sources WHERE THEY LIE under
# $MILC_REF (default /root/reference) into oracle/_ref/ (git-ignored, travels to
# the GPU box).  This is our own short recipe, not the reference's build system;
# the flags are the ones its default vanilla build uses (Makefile:814-818:
# -DDBLSTORE_FN -DFEWSUMS -DD_FN_GATHER13, libraries/Make_vanilla: -O3 -DFAST).
# Outputs:
#   oracle/_ref/libmilcref.so       double precision, single thread
#   oracle/_ref/libmilcref_omp.so   double precision, OpenMP site loops (-DOMP)
#   oracle/_ref/libmilcref_f.so     single precision (MILC_PRECISION=1), single thread
# Nothing is written to $MILC_REF and no reference source is copied into the repo.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${MILC_REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/generic_ks" ]; then
  echo "build_ref.sh: reference tree not found at $REF (nothing to do)"; exit 0
fi
mkdir -p "$OUT/gen"
# the reference's own make copies the chosen action header to quark_action.h
# (ks_imp_utilities/Make_template:118-119); it is a build artefact, not source.
cp "$REF/generic_ks/imp_actions/hisq/hisq_u3_action.h" "$OUT/gen/quark_action.h"

LIBSRC=$(cd "$REF/libraries" && ls *.c | grep -v "^prefetch32.c$\|^prefetch64.c$")
GEN="com_vanilla.c layout_hyper_prime.c make_lattice.c field_utilities.c ranstuff.c"
GKS="dslash_fn_dblstore.c fn_links_milc.c d_congrad5_fn_milc.c ks_multicg_offset.c fermion_links_fn_twist_milc.c"
# HISQ link construction: U -> V (fat7) -> W (U(3) projection) -> fat, long
GEN="$GEN general_staple.c path_product.c gauge_utilities.c project_su3_hit.c reunitarize2.c stout_smear.c"
GKS="$GKS fermion_links_hisq_load_milc.c fermion_links_fn_load_milc.c ks_action_paths_hisq.c su3_mat_op.c rephase.c"
# the HISQ fermion force (oracle only so far), ks_imp_rhmc's flags
# (ks_imp_rhmc/Make_template: -DHISQ_FF_MULTI_WRAPPER -DHISQ_FORCE_FILTER=5.0e-5, KS_MULTIFF=FNMAT)
GKS="$GKS fermion_force_hisq_multi.c fermion_links_hisq_milc.c fermion_links.c ff_opt.c path_transport.c"
# meson tie-ups: the contraction and the sink spin-taste operators
GKS="$GKS ks_meson_mom.c spin_taste_ops.c"
GWI="gammas.c"
# the UML propagator-solve sequence
GEN="$GEN report_invert_status.c"
GKS="$GKS mat_invert.c d_congrad5_fn.c"

# incremental eigCG: generic_ks/inc_eigcg.c needs LAPACK/BLAS; the image has none
# installed system-wide, but the OpenBLAS inside its Python packages exports the plain Fortran symbols
LAPACK_LIB="${MILC_REF_LAPACK:-$(ls /opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs/libopenblas*.so 2>/dev/null | head -1)}"
if [ -n "$LAPACK_LIB" ] && [ -f "$LAPACK_LIB" ]; then
  echo "eigCG reference: linking $LAPACK_LIB"
else
  LAPACK_LIB=""
  echo "eigCG reference: no LAPACK found, inc_eigcg.c left out (eigCG oracle unpinned)"
fi

build_variant() {  # name precision extra-flags
  local name="$1" prec="$2" extra="$3"
  local obj="$OUT/obj_$name"
  mkdir -p "$obj"
  local CF="-O3 -fPIC -std=gnu99 -w $extra -DMILC_PRECISION=$prec -DFAST"
  # (-DHISQ_REUNIT_*: the reunitarisation flags of the RHMC build, ks_imp_rhmc/Make_template:204-206)
  local AF="$CF -DDBLSTORE_FN -DFEWSUMS -DD_FN_GATHER13 -DC_GLOBAL_INLINE -DFN -DHAVE_KS \
            -DHISQ_REUNIT_ALLOW_SVD -DHISQ_REUNIT_SVD_REL_ERROR=1e-8 -DHISQ_REUNIT_SVD_ABS_ERROR=1e-8 \
            -DKS_MULTIFF=FNMAT -DHISQ_FF_MULTI_WRAPPER -DHISQ_FORCE_FILTER=5.0e-5 \
            -D_FILE_OFFSET_BITS=64 -I$HERE/ref_harness -I$OUT/gen"
  local jobs=()
  for f in $LIBSRC; do
    echo "gcc -c $CF $REF/libraries/$f -o $obj/lib_${f%.c}.o"
  done > "$obj/cmds.txt"
  for f in $GEN; do echo "gcc -c $AF -I$REF/generic $REF/generic/$f -o $obj/gen_${f%.c}.o"; done >> "$obj/cmds.txt"
  for f in $GKS; do echo "gcc -c $AF -I$REF/generic_ks $REF/generic_ks/$f -o $obj/gks_${f%.c}.o"; done >> "$obj/cmds.txt"
  for f in $GWI; do echo "gcc -c $AF -I$REF/generic_wilson $REF/generic_wilson/$f -o $obj/gwi_${f%.c}.o"; done >> "$obj/cmds.txt"
  echo "gcc -c $AF -I$REF/generic_ks $HERE/ref_harness/meson_harness.c -o $obj/meson_harness.o" >> "$obj/cmds.txt"
  echo "gcc -c $AF -I$REF/generic_ks $HERE/ref_harness/harness.c -o $obj/harness.o" >> "$obj/cmds.txt"
  local lapack=""
  if [ -n "$LAPACK_LIB" ] && [ "$prec" = "2" ]; then   # (inc_eigcg.c requires double precision)
    echo "gcc -c $AF -I$REF/generic_ks $REF/generic_ks/inc_eigcg.c -o $obj/gks_inc_eigcg.o" >> "$obj/cmds.txt"
    echo "gcc -c $AF -I$REF/generic_ks $HERE/ref_harness/eigcg_harness.c -o $obj/eigcg_harness.o" >> "$obj/cmds.txt"
    lapack="$LAPACK_LIB -Wl,--disable-new-dtags -Wl,-rpath,$(dirname "$LAPACK_LIB")"   # (DT_RPATH: also for OpenBLAS's own libgfortran)
  fi
  # a few library files are platform-specific and may not compile; they are not on the path
  xargs -P "$(nproc)" -I{} sh -c '{} 2>/dev/null || echo "skip: {}" | cut -c1-200 >&2' < "$obj/cmds.txt"
  gcc -shared $extra -o "$OUT/libmilcref$name.so" "$obj"/*.o $lapack -lm
  echo "built $OUT/libmilcref$name.so"
}

build_variant ""     2 ""
build_variant "_omp" 2 "-fopenmp -DOMP"
build_variant "_f"   1 ""
