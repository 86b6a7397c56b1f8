cd $GRAFT_REPO_ROOT
(timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "f16x3" 2>&1 | tail -12) > gpurun_out/r2_j_tests.log
if grep -q "1 passed" gpurun_out/r2_j_tests.log; then
(timeout 400 python tools/bench_gemm2.py 2>&1 | tail -20) > gpurun_out/r2_j_gemm2.log
(timeout 200 python tools/gemm_trace2.py 22726 256 2048 > gpurun_out/r2_j_trace.log 2>&1)
(timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r2_j_bench.json) 2> gpurun_out/r2_j_bench.err
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "config2 or golden or runner or c256 or gemm" 2>&1 | tail -5) >> gpurun_out/r2_j_tests.log
fi
cat gpurun_out/r2_j_tests.log | tail -6; cat gpurun_out/r2_j_gemm2.log; tail -22 gpurun_out/r2_j_trace.log; python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/r2_j_bench.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['fresh_masks_value'], j['gpu_launches_per_step'], j['roofline_gemm']['kernel_ms_per_step'])
except Exception as e: print('bench ERR', e)
PY
