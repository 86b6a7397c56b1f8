#!/bin/bash
mkdir -p gpurun_out
for cfg in "256 256 1 1" "1000 2048 1 3" "18944 2048 1" "22726 2048 4" "22726 2048 1"; do
  echo "== $cfg"; timeout 100 python tools/ffn_debug.py $cfg 2>&1 | tail -2
done 2>&1 | tee gpurun_out/r2_n2_cfgs.log
