#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ffn_fused" 2>&1 | tail -5
timeout 200 python tools/ffn_trace.py 22726 2048 1 > gpurun_out/r2_ffn_trace_v5.txt 2>&1
timeout 300 python tools/bench_ffn.py 2>&1 | tee gpurun_out/r2_n_bench_ffn_v5.txt
