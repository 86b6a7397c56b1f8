#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python bench.py > gpurun_out/r2_final_bench.json) 2> gpurun_out/r2_final_bench.err
echo "rc=$? stdout lines: $(wc -l < gpurun_out/r2_final_bench.json)"
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_final_bench.json'))
print(j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], 'fresh', j['e2e']['fresh_masks_value'], 'roofline', j['roofline']['frac'], 'gemm', j['roofline_gemm']['frac'], j['clocks']['sm_mhz'], 'cmp', j.get('gpu_comparator',{}).get('value'), 'cpu', j.get('cpu_baseline',{}).get('value'), 'launches', j['gpu_launches'])
PY
