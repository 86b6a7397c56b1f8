cd $GRAFT_REPO_ROOT
(timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "f16x3" 2>&1 | tail -15) > gpurun_out/r2_c_tests.log
(timeout 400 python tools/bench_gemm2.py 2>&1 | tail -30) > gpurun_out/r2_c_gemm2.log
(timeout 600 python tools/msda_probe.py "g8:" "w32_1:msda_warp_per_item=1" "w32_2:msda_warp_per_item=2" "w32_1_c128:msda_warp_per_item=1,msda_chunk=128" "w32_1_c32:msda_warp_per_item=1,msda_chunk=32" "w32_2_c128:msda_warp_per_item=2,msda_chunk=128" 2>&1 | tail -25) > gpurun_out/r2_c_probe.log
(timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "config2" 2>&1 | tail -15) >> gpurun_out/r2_c_tests.log
(timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r2_c_bench.json) 2> gpurun_out/r2_c_bench.err
cat gpurun_out/r2_c_probe.log; cat gpurun_out/r2_c_gemm2.log; grep -E "passed|failed" gpurun_out/r2_c_tests.log; tail -c 400 gpurun_out/r2_c_bench.err
