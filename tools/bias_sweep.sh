#!/bin/bash
# parity error of every model against the truncation-bias constant (profiling build reads CPB200_BIAS_ULPS_PER_STEP)
export CPB200_LIB=$PWD/cosmopower_b200/csrc/libcpb200_prof.so
for b in 0.16 0.21 0.26 0.31; do
  export CPB200_BIAS_ULPS_PER_STEP=$b
  echo "== bias $b ulp per MMA step"
  timeout 300 python tests/manual/accuracy_sweep.py ${1:-60000} 2>&1 | grep "max err"
done
