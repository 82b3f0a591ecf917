#!/bin/bash
# builds profiles/variants/libb200ks_<name>.so (see profiles/variants/README)
set -e
cd "$(dirname "$0")/.."
FLAGS="--threads 3 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -Xcompiler -fvisibility=default -I include"
SRC=$(ls milc_qcd_b200/csrc/*.cu)
build() { nvcc $FLAGS $2 -o profiles/variants/libb200ks_$1.so $SRC; echo built $1; }
build mb5 -DB200KS_HALF_MINBLOCKS=5 &
build mb7 -DB200KS_HALF_MINBLOCKS=7 &
wait
build probe_nomath -DB200KS_PROBE_NOMATH &
build probe_nolinkload -DB200KS_PROBE_NOLINKLOAD &
wait
# how boundary CTAs of a partitioned stencil poll the arrival flags (dslash.cuh acquire_halo_cta)
for a in 0 1 2 3; do build acq$a -DB200KS_ACQ=$a & done
wait
# fused backward staple body of the fermion force compiled for 2 / 3 / 4 CTAs per SM (254 / 168 / 128 registers)
for mb in 2 3 4; do build fbwd$mb -DB200KS_FORCE_BWD_MINB=$mb & done
wait
