cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r2_e_pytest.log
(timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_step.csv python tools/profile_step.py > gpurun_out/r2_e_ncu.log 2>&1)
python tools/summarize_launches.py gpurun_out/r2_launches_step.csv 60 > gpurun_out/r2_launches_step_summary.txt 2>&1
(timeout 600 python bench.py --mode train --steps 6 --warmup 3 > gpurun_out/r2_e_train1.json) 2> gpurun_out/r2_e_train1.err
(timeout 900 python bench.py > gpurun_out/r2_e_bench.json) 2> gpurun_out/r2_e_bench.err
tail -12 gpurun_out/r2_e_pytest.log; head -30 gpurun_out/r2_launches_step_summary.txt; tail -c 800 gpurun_out/r2_e_train1.json; tail -c 300 gpurun_out/r2_e_train1.err; tail -c 300 gpurun_out/r2_e_bench.err
