timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
for m in cmb_TT_NN PKLIN_NN cmb_PP_PCAplusNN cmb_TE_PCAplusNN; do timeout 120 python tools/role_profile.py $m 1048576 > gpurun_out/role_$m.log 2>&1; grep -E "rows=|mma_wait|epi_act|epi_out|mma_total" gpurun_out/role_$m.log; done
