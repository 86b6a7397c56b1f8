cd $GRAFT_REPO_ROOT
(timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "f16x3" 2>&1 | tail -12) > gpurun_out/r2_j_tests.log
if grep -q "1 passed" gpurun_out/r2_j_tests.log; then
(timeout 400 python tools/bench_gemm2.py 2>&1 | tail -20) > gpurun_out/r2_j_gemm2.log
(timeout 200 python tools/gemm_trace2.py 22726 256 2048 > gpurun_out/r2_j_trace.log 2>&1)
fi
cat gpurun_out/r2_j_tests.log | tail -4; cat gpurun_out/r2_j_gemm2.log; tail -22 gpurun_out/r2_j_trace.log
