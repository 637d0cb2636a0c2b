#!/bin/bash
# same-box A/B of linearize builds, one model per group:  scripts/ab_lin2.sh "<model>:<lib>,<lib>,..." ...   ("default" = in-tree libacm.so)
for rep in 1 2; do for grp in "$@"; do
  model=${grp%%:*}; libs=${grp#*:}
  for lib in ${libs//,/ }; do
    if [ "$lib" = default ]; then unset ACM_LIB_PATH; else export ACM_LIB_PATH=$PWD/build/ab/libacm_$lib.so; fi
    echo -n "rep$rep $lib: "; MODELS=$model REPS=${REPS:-30} python scripts/lin_bench.py | tail -${TAILN:-1} | tr "\n" "|"; echo
  done
done; done
