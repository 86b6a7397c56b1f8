#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/bench_predictor.py 1000 4096 20000 2>&1 | tail -6 | tee gpurun_out/r2_bench_predictor.txt
