#!/bin/bash
mkdir -p gpurun_out
for cfg in "18181 2048 1" "22726 2048 4" "9090 2048 2"; do
timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,sm__cycles_elapsed.max --clock-control none -k regex:ffn --csv python tools/ffn_debug.py $cfg 2>&1 | grep -i "ffn_\|max err" | cut -c1-400
done > gpurun_out/r2_ffn_ncu_times.txt 2>&1
cat gpurun_out/r2_ffn_ncu_times.txt | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | cut -c1-200
