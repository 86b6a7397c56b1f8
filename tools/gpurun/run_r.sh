#!/bin/bash
# final validation on one GPU: full suite, smoke, default bench (with cpu baseline + comparator)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/r2_r_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -5) > gpurun_out/r2_r_smoke.log
(timeout 900 python bench.py > gpurun_out/r2_r_bench.json) 2> gpurun_out/r2_r_bench.err
cat gpurun_out/r2_r_pytest.log | tail -2; tail -2 gpurun_out/r2_r_smoke.log
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_r_bench.json'))
print(j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], 'fresh', j['e2e']['fresh_masks_value'], 'roofline', j['roofline']['frac'], 'gemm', j['roofline_gemm']['frac'], j['clocks']['sm_mhz'], 'cmp', j.get('gpu_comparator',{}).get('value'), 'cpu', j.get('cpu_baseline',{}).get('value'))
PY
