#!/bin/bash
# Elapsed time of the four shipped decks through the drop-in CLI, with the environment given on the command line:
#   tools/deck_times.sh [VAR=value ...]
cd "$(dirname "$0")/.."
D=$PWD/tests/golden/decks
EXE=$PWD/mpilattice-boltzmann_b200/bin/d2q9-bgk
for name in 128x128 128x256 256x256 1024x1024; do
  W=$(mktemp -d)
  ( cd "$W" && env "$@" LBM_VERBOSE=1 LBM_FINAL_STATE=0 "$EXE" "$D/input_$name.params" "$D/obstacles_$name.dat" > run.out 2> run.err
    echo "$name $* $(grep 'Elapsed time' run.out | tr -s '\t' ' ') $(grep Reynolds run.out | tr -s '\t' ' ') $(cat run.err)" )
  rm -rf "$W"
done
