#!/bin/bash
# A/B of the job-table schedules on the GPU box (profiling build reads CPB200_SCHEDULE)
export CPB200_LIB=$PWD/cosmopower_b200/csrc/libcpb200_prof.so
for s in simple merge overlap default; do
  if [ "$s" = default ]; then unset CPB200_SCHEDULE; else export CPB200_SCHEDULE=$s; fi
  echo "== schedule $s"
  timeout 200 python tools/time_models.py 1048576 ${1:-cmb_PP_PCAplusNN,PKLIN_NN,cmb_TT_NN,cmb_TE_PCAplusNN} 2>&1 | grep rows=
done
