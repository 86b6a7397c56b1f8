#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in 1 0; do
(SDETR_FUSED_FFN=$c timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r2_o_bench_$c.json) 2> gpurun_out/r2_o_bench_$c.err
done
(timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -6) > gpurun_out/r2_o_tests.log
python - <<'PY'
import json
for c in (1, 0):
    try:
        j=json.load(open(f'gpurun_out/r2_o_bench_{c}.json')); print(c, j['value'], j['ms_per_step'], j['e2e']['value'], j['gpu_launches_per_step'], j['roofline_gemm']['kernel_ms_per_step'], j['roofline_gemm']['frac'], j['clocks'])
    except Exception as e: print(c, 'ERR', e)
PY
tail -4 gpurun_out/r2_o_tests.log
