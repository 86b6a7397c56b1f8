cd $GRAFT_REPO_ROOT
(timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "f16x3 or mask_plan" 2>&1 | tail -30) > gpurun_out/r2_b_f16test.log
if grep -q "2 passed" gpurun_out/r2_b_f16test.log; then K=f16x3; else K=tf32x3; fi
echo "kernel for the suite: $K" >> gpurun_out/r2_b_f16test.log
(timeout 400 python tools/bench_gemm2.py 2>&1 | tail -30) > gpurun_out/r2_b_gemm2.log
(SDETR_GEMM_KERNEL=$K timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60) > gpurun_out/r2_b_pytest.log
for t in 256 512 1024; do c=$((t/4)); (timeout 300 python tools/msda_probe.py msda_threads=$t msda_chunk=$c 2>&1 | tail -8) > gpurun_out/r2_b_probe_$t.log; done
(timeout 300 python tools/msda_probe.py msda_threads=1024 msda_chunk=512 2>&1 | tail -8) > gpurun_out/r2_b_probe_1024c512.log
(SDETR_GEMM_KERNEL=$K timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r2_b_bench.json) 2> gpurun_out/r2_b_bench.err
tail -3 gpurun_out/r2_b_f16test.log; tail -5 gpurun_out/r2_b_pytest.log; tail -c 600 gpurun_out/r2_b_bench.err
