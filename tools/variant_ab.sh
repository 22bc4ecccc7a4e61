#!/bin/bash
# time every library variant under cosmopower_b200/csrc/variants/ (profiling builds) for the given schedule(s) and models
models=${1:-cmb_PP_PCAplusNN,PKLIN_NN,cmb_TT_NN}
scheds=${2:-simple}
for lib in cosmopower_b200/csrc/variants/lib_*.so; do
  for s in $scheds; do
    export CPB200_LIB=$PWD/$lib
    if [ "$s" = default ]; then unset CPB200_SCHEDULE; else export CPB200_SCHEDULE=$s; fi
    echo "== $(basename $lib) schedule $s"
    timeout 200 python tools/time_models.py 1048576 $models 2>&1 | grep rows=
  done
done
