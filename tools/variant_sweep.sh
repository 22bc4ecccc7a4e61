#!/bin/bash
# Tuning aid: time every kernel variant built into cosmopower_b200/csrc/variants/ (clock-independent SM cycles of the
# slowest CTA, 1M rows) - usage: tools/variant_sweep.sh [models]
MODELS=${1:-cmb_TT_NN,cmb_PP_PCAplusNN,PKLIN_NN}
for lib in cosmopower_b200/csrc/variants/lib_*.so; do
  echo "== $lib"
  CPB200_LIB=$PWD/$lib timeout 300 python tools/time_models.py 1048576 $MODELS 2>&1 | grep rows=
done
