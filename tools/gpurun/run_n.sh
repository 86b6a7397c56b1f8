#!/bin/bash
# fused FFN: parity test, then the micro-benchmark
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ffn_fused" 2>&1 | tail -15 | tee gpurun_out/r2_n_pytest.log
timeout 300 python tools/bench_ffn.py 2>&1 | tee gpurun_out/r2_n_bench_ffn.txt
