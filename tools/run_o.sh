#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/profile_eager_cpu.py 2>&1 | tail -45 > gpurun_out/r2_eager_cpu_profile2.txt; head -36 gpurun_out/r2_eager_cpu_profile2.txt | cut -c1-150
(timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r2_o_bench_f.json) 2> gpurun_out/r2_o_bench_f.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_o_bench_f.json')); print(j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], 'serial', j['e2e']['serial_value'], 'fresh', j['e2e']['fresh_masks_value'], 'fresh serial', j['e2e']['fresh_masks_serial_value'])
PY
