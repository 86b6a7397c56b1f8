#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_step_ffn.csv python tools/profile_step.py > gpurun_out/r2_p_ncu.log 2>&1)
python tools/summarize_launches.py gpurun_out/r2_launches_step_ffn.csv 60 > gpurun_out/r2_launches_step_ffn_summary.txt 2>&1
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffn_fused_kernel -c 1 -o gpurun_out/r2_ffn_fused_full -f python tools/ffn_debug.py 22726 2048 1 > gpurun_out/r2_p_ncu2.log 2>&1)
ncu -i gpurun_out/r2_ffn_fused_full.ncu-rep --page details --csv > gpurun_out/r2_ffn_fused_ncu_full.csv 2>/dev/null
head -32 gpurun_out/r2_launches_step_ffn_summary.txt; grep -n "ffn" gpurun_out/r2_launches_step_ffn.csv | awk -F'","' '{print $5, $NF}' | cut -c1-30,140-220
