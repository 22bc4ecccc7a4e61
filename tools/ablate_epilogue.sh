#!/bin/bash
# Profiling ablations of the epilogue (results are garbage, only the cycle counts matter):
# 0 = normal, 1 = no epilogue work, 2 = TMEM reads only, 4 = no global stores, 8 = no constant loads, 12 = neither
for f in ${ABLATIONS:-0 1 2 4 8 12}; do
  CPB200_DEBUG_EPILOGUE=$f python tools/role_profile.py ${MODEL:-cmb_TT_NN} > gpurun_out/abl_$f.log 2>&1
done
