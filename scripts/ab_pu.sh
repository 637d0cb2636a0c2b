#!/bin/bash
# same-box A/B of project / unproject / round-trip builds: scripts/ab_pu.sh <models> <lib1> <lib2> ...  ("default" = in-tree libacm.so)
models=$1; shift
for rep in 1 2; do for lib in "$@"; do
  if [ "$lib" = default ]; then unset ACM_LIB_PATH; else export ACM_LIB_PATH=$PWD/build/ab/libacm_$lib.so; fi
  echo "== $lib (rep $rep)"; MODELS=$models REPS=20 python scripts/pu_bench.py
done; done
