#!/bin/bash
# one `ncu --set full` capture of a 1 048 576-row launch per model (after the same command has run clean without ncu)
for m in ${1:-cmb_TT_NN cmb_EE_NN cmb_TE_PCAplusNN cmb_PP_PCAplusNN PKLIN_NN PKNLBOOST_NN}; do
  src=""; case $m in cmb_TT_NN|cmb_PP_PCAplusNN) src="--import-source on";; esac
  timeout 120 python tools/time_models.py 1048576 $m > gpurun_out/plain_$m.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none $src -k regex:fused_mlp -s 3 -c 1 -f -o gpurun_out/prof_$m python tools/time_models.py 1048576 $m > gpurun_out/ncu_$m.log 2>&1
  tail -1 gpurun_out/ncu_$m.log
done
ls -la gpurun_out/*.ncu-rep
