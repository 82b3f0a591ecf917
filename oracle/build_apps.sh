#!/bin/bash
# oracle/build_apps.sh -- TEST INFRASTRUCTURE (end-to-end drop-in check).
#
# Builds two of the reference's own applications from the sources WHERE THEY LIE under
# $MILC_REF (default /root/reference), each in two flavours, into oracle/_ref/apps/:
#
#   ks_spectrum_hisq_cpu   su3_rhmc_hisq_cpu     the reference's default vanilla CPU build
#                                                (-DDBLSTORE_FN -DFEWSUMS -DD_FN_GATHER13)
#   ks_spectrum_hisq_b200  su3_rhmc_hisq_b200    the SAME unmodified sources built the way the
#                                                reference builds against QUDA (-DHAVE_QUDA
#                                                -DUSE_CG_GPU, Makefile:419-474) but linked to
#                                                milc_qcd_b200/libb200ks.so through
#                                                include/quda_milc_interface.h  (route 2)
#   ks_spectrum_hisq_b200fl su3_rhmc_hisq_b200fl as _b200 plus -DUSE_FL_GPU (WANT_FL_GPU=true): the
#                                                HISQ links are built on the GPU as well
#   su3_rhmc_hisq_b200ff                         as _b200fl plus -DUSE_FF_GPU (WANT_FF_GPU=true): the
#                                                HISQ fermion force on the GPU too
#
# It also stages the sample inputs, golden outputs, tolerance files and sample lattices the
# reference's own regression uses (ks_spectrum/test, ks_imp_rhmc/test, binary_samples) into
# oracle/_ref/samples/ so the same regression can run on the GPU box, where /root/reference
# does not exist.  Everything lands in the git-ignored oracle/_ref/; no reference source is
# copied (one file needs a one-word gcc-13 fix and is piped through sed into the compiler).
# This is our own recipe: the object lists mirror Make_template_combos / the applications'
# Make_template, the flags mirror the top-level Makefile.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${MILC_REF:-/root/reference}"
OUT="$HERE/_ref"
[ -d "$REF/generic_ks" ] || { echo "build_apps.sh: no reference tree at $REF (nothing to do)"; exit 0; }
[ -f "$ROOT/milc_qcd_b200/libb200ks.so" ] || { echo "build libb200ks.so first (python -m milc_qcd_b200.build)"; exit 1; }
mkdir -p "$OUT/gen" "$OUT/apps" "$OUT/samples"
cp "$REF/generic_ks/imp_actions/hisq/hisq_u3_action.h" "$OUT/gen/quark_action.h"
cp "$REF/generic/imp_actions/symanzik_1loop_hisq_action.h" "$OUT/gen/gauge_action.h"

COMMON="-O3 -std=c99 -w -DSINGLE -DCGTIME -DCG_OK -DREMAP_STDIO_APPEND -DKS_MULTIFF=FNMAT -DCL_CG=BICG \
 -DC_GLOBAL_INLINE -DMILC_PRECISION=2 -D_FILE_OFFSET_BITS=64 -D_LARGEFILE64_SOURCE -DFN \
 -DHISQ_REUNIT_ALLOW_SVD -DHISQ_REUNIT_SVD_REL_ERROR=1e-8 -DHISQ_REUNIT_SVD_ABS_ERROR=1e-8"
CPU_FLAGS="-DDBLSTORE_FN -DFEWSUMS -DD_FN_GATHER13"
GPU_FLAGS="-DFEWSUMS -DHAVE_QUDA -DUSE_CG_GPU -DSET_QUDA_SUMMARIZE -I$ROOT/include -I/usr/local/cuda/include"

GENERIC_BASE="ape_smear check_unitarity d_plaq4 gaugefix2 io_lat4 momentum_twist nersc_cksum path_product \
 project_su3_hit reunitarize2 show_generic_opts show_scidac_opts layout_hyper_prime field_translation \
 field_utilities gauge_utilities io_detect io_helpers io_lat_utils make_lattice ranstuff remap_stdio_from_args \
 io_ansi com_vanilla general_staple stout_smear report_invert_status"
GKS_BASE="charge_utilities fermion_links_from_site f_meas gauss_smear_ks grsource_imp naik_eps_utilities \
 path_transport rephase show_generic_ks_opts show_hisq_links_opts fermion_links_hisq_milc \
 fermion_links_hisq_load_milc fermion_links fermion_links_fn_load_milc fermion_links_fn_twist_milc fn_links_milc \
 ks_action_paths_hisq su3_mat_op d_congrad5_two_src d_congrad5_fn_milc mat_invert ks_invert d_congrad5_fn \
 d_congrad_opt ks_multicg ks_multicg_offset"

SPEC_APP="control gauge_info ks_source_info ksprop_info make_prop setup spectrum_ks"
SPEC_GENERIC="restrict_fourier discretize_wf io_source_cmplx_fm phases quark_source quark_source_io quark_source_sink_op"
SPEC_GKS="io_helpers_ks io_prop_ks spin_taste_ops ks_baryon ks_meson_mom eigen_stuff_helpers io_helpers_ks_eigen \
 io_ks_eigen jacobi eigen_stuff_Ritz eigen_stuff_PRIMME"
SPEC_DEFS="-DKS_MULTICG=HYBRID -DMULTISOURCE -DHAVE_KS -DKalkreuter_Ritz"

RHMC_APP="d_action_rhmc eo_fermion_force_rhmc gauge_info grsource_rhmc ks_ratinv load_rhmc_params setup \
 update_h_rhmc update_rhmc update_u control"
RHMC_GENERIC="ploop3 gauge_force_imp gauge_stuff ranmom"
RHMC_GKS="gauge_force_imp_ks reunitarize_ks show_generic_ks_md_opts fermion_force_hisq_multi show_hisq_force_opts ff_opt"
RHMC_DEFS="-DKS_MULTICG=HYBRID -DINT_ALG=INT_3G1F -DHISQ_FF_MULTI_WRAPPER -DHISQ_FORCE_FILTER=5.0e-5 -DHMC"

# libraries (su3 + complex), shared by all flavours
LIBOBJ="$OUT/obj_apps_lib"
if [ ! -f "$LIBOBJ/.done" ]; then
  mkdir -p "$LIBOBJ"
  (cd "$REF/libraries" && ls *.c | grep -v "^prefetch32.c$\|^prefetch64.c$") | \
    xargs -P "$(nproc)" -I{} sh -c "gcc -c -O3 -w -DFAST -DMILC_PRECISION=2 $REF/libraries/{} -o $LIBOBJ/{}.o 2>/dev/null || true"
  touch "$LIBOBJ/.done"
fi

build_app() {  # name appdir flavour(cpu|b200) appfiles generic gks defs
  local name="$1" appdir="$2" flav="$3" appf="$4" genf="$5" gksf="$6" defs="$7"
  local obj="$OUT/obj_${name}_${flav}"
  mkdir -p "$obj"
  local fl="$COMMON $defs -I$REF/$appdir -I$OUT/gen"
  local dsl="dslash_fn_dblstore" extra_gen="" extra_gks=""
  if [ "$flav" = "b200" ]; then
    fl="$fl $GPU_FLAGS"; dsl="dslash_fn"; extra_gen="milc_to_quda_utilities"; extra_gks="d_congrad5_fn_gpu ks_multicg_offset_gpu"
  elif [ "$flav" = "b200fl" ]; then
    # additionally WANT_FL_GPU=true (Makefile:453-455, Make_template_combos:167-171,190): the fermion
    # links are built through qudaLoadUnitarizedLink / qudaLoadKSLink
    fl="$fl $GPU_FLAGS -DUSE_FL_GPU"; dsl="dslash_fn"; extra_gen="milc_to_quda_utilities"
    extra_gks="d_congrad5_fn_gpu ks_multicg_offset_gpu fermion_links_fn_load_gpu"
  elif [ "$flav" = "b200ff" ]; then
    # additionally WANT_FF_GPU=true (Makefile:458-461): the HISQ fermion force goes through qudaHisqForce
    fl="$fl $GPU_FLAGS -DUSE_FL_GPU -DUSE_FF_GPU"; dsl="dslash_fn"; extra_gen="milc_to_quda_utilities"
    extra_gks="d_congrad5_fn_gpu ks_multicg_offset_gpu fermion_links_fn_load_gpu"
  else
    fl="$fl $CPU_FLAGS"
  fi
  : > "$obj/cmds.txt"
  for f in $appf; do echo "gcc -c $fl $REF/$appdir/$f.c -o $obj/app_$f.o" >> "$obj/cmds.txt"; done
  for f in $GENERIC_BASE $genf $extra_gen; do echo "gcc -c $fl $REF/generic/$f.c -o $obj/gen_$f.o" >> "$obj/cmds.txt"; done
  for f in $GKS_BASE $gksf $extra_gks $dsl; do
    if [ "$f" = "io_helpers_ks_eigen" ]; then
      # gcc 13: "static declaration follows non-static" vs include/io_ks_eigen.h:133 -- drop the
      # keyword on the fly, nothing is written next to the reference
      echo "sed '82s/^static //' $REF/generic_ks/$f.c | gcc -c $fl -I$REF/generic_ks -x c - -o $obj/gks_$f.o" >> "$obj/cmds.txt"
    else
      echo "gcc -c $fl $REF/generic_ks/$f.c -o $obj/gks_$f.o" >> "$obj/cmds.txt"
    fi
  done
  [ "$name" = "ks_spectrum_hisq" ] && echo "gcc -c $fl $REF/generic_wilson/gammas.c -o $obj/gw_gammas.o" >> "$obj/cmds.txt"
  xargs -d '\n' -P "$(nproc)" -I{} bash -c {} < "$obj/cmds.txt"
  local exe="$OUT/apps/${name}_${flav}"
  if [ "$flav" != "cpu" ]; then
    g++ -o "$exe" "$obj"/*.o "$LIBOBJ"/*.o -L"$ROOT/milc_qcd_b200" -lb200ks \
        -Wl,-rpath,'$ORIGIN/../../../milc_qcd_b200' -L/usr/local/cuda/lib64 -lcudart -lm
  else
    gcc -o "$exe" "$obj"/*.o "$LIBOBJ"/*.o -lm
  fi
  echo "built $exe"
}

for flav in cpu b200 b200fl; do
  build_app ks_spectrum_hisq ks_spectrum "$flav" "$SPEC_APP" "$SPEC_GENERIC" "$SPEC_GKS" "$SPEC_DEFS"
  build_app su3_rhmc_hisq ks_imp_rhmc "$flav" "$RHMC_APP" "$RHMC_GENERIC" "$RHMC_GKS" "$RHMC_DEFS"
done
# links, solves AND the fermion force on the GPU (RHMC only: spectroscopy has no force)
build_app su3_rhmc_hisq ks_imp_rhmc b200ff "$RHMC_APP" "$RHMC_GENERIC" "$RHMC_GKS" "$RHMC_DEFS"

# stage the reference's regression fixtures (data, not source)
S="$OUT/samples"
mkdir -p "$S/ks_spectrum/test" "$S/ks_imp_rhmc/test" "$S/binary_samples"
cp "$REF"/ks_spectrum/test/ks_spectrum_hisq.*.2.sample-in "$REF"/ks_spectrum/test/ks_spectrum_hisq.*.2.sample-out \
   "$REF"/ks_spectrum/test/ks_spectrum_hisq.*.2.errtol "$REF"/ks_spectrum/test/ks_spectrum_hisq.*.2.corrfile_t0.* \
   "$S/ks_spectrum/test/" 2>/dev/null || true
cp "$REF"/ks_imp_rhmc/test/su3_rhmc_hisq.2.* "$REF"/ks_imp_rhmc/test/rationals.sample.su3_rhmc_hisq "$S/ks_imp_rhmc/test/"
for f in lat.sample.l8888 lat.sample.l6666.hisq lat.sample.l4448.gf lat.sample.l6666; do
  [ -f "$REF/binary_samples/$f" ] && cp "$REF/binary_samples/$f" "$S/binary_samples/"
done
echo "staged samples in $S"
