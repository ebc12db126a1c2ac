#!/bin/bash
# Same-box A/B of builds of the library: tools/session_ab.sh <out-dir> "<bench flags>" <variant> [<variant> ...]
# (variant X = lib/liblbm_b200_X.so, "base" = the library as built; base runs first and last)
L=mpilattice-boltzmann_b200/lib
OUT=$1; FLAGS=$2; shift 2
cp $L/liblbm_b200.so $L/liblbm_b200_base.so
for v in base "$@" base; do
  cp $L/liblbm_b200_$v.so $L/liblbm_b200.so
  echo "## $v" | tee -a $OUT.log
  tools/sweep_r02.sh $OUT "$FLAGS" | grep '^{' | cut -c1-120 | tee -a $OUT.log
done
cp $L/liblbm_b200_base.so $L/liblbm_b200.so
