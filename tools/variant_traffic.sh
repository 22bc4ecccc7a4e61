#!/bin/bash
# Tuning aid: DRAM bytes of one 1M-row launch for every kernel variant in cosmopower_b200/csrc/variants/
for lib in cosmopower_b200/csrc/variants/lib_*.so; do
  for m in ${1:-cmb_TT_NN cmb_PP_PCAplusNN}; do
    CPB200_LIB=$PWD/$lib timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fused_mlp -s 3 -c 1 --csv python tools/time_models.py 1048576 $m 2>/dev/null | grep dram__ | awk -F'","' -v l=$lib -v m=$m '{printf "%s %s %s %s %s\n", l, m, $(NF-2), $(NF-1), $NF}'
  done
done
