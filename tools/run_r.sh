#!/bin/bash
# final validation on one GPU: full suite, poisoned allocator, smoke, default bench (with cpu baseline + comparator), reference arm
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/r2_r_pytest.log
(SDETR_POISON=0xff timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/r2_r_pytest_poison.log
(timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -5) > gpurun_out/r2_r_smoke.log
(timeout 1200 python bench.py > gpurun_out/r2_r_bench.json) 2> gpurun_out/r2_r_bench.err
(timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_r_bench_ref.json) 2> gpurun_out/r2_r_bench_ref.err
tail -3 gpurun_out/r2_r_pytest.log gpurun_out/r2_r_pytest_poison.log gpurun_out/r2_r_smoke.log
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_r_bench.json'))
print(j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], 'roofline', j['roofline']['frac'], 'gemm', j['roofline_gemm']['frac'], j['clocks'], 'cmp', j.get('gpu_comparator',{}).get('value'), 'cpu', j.get('cpu_baseline',{}).get('value'))
try:
    r=json.load(open('gpurun_out/r2_r_bench_ref.json')); print('ref', r.get('value'), r.get('cpu_baseline'))
except Exception as e: print('ref ERR', e)
PY
