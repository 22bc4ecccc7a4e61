# full single-GPU validation: tests, smoke, bench (both arms), ncu launch list + full capture
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-batch 65536 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-batch 65536 > gpurun_out/ncu1.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-batch 65536 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fused_mlp -s 6 -c 1 -o gpurun_out/prof_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-batch 65536 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
for m in cmb_TT_NN cmb_EE_NN cmb_TE_PCAplusNN cmb_PP_PCAplusNN PKLIN_NN PKNLBOOST_NN; do timeout 120 python tools/role_profile.py $m 1048576 2>&1 | grep "rows="; done | tee gpurun_out/all_models.log
