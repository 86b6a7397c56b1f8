#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3) > gpurun_out/r2_o_pytest.log; tail -2 gpurun_out/r2_o_pytest.log
(timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r2_o_bench.json) 2> gpurun_out/r2_o_bench.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_o_bench.json'))
print(j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], 'serial', j['e2e']['serial_value'], 'fresh', j['e2e']['fresh_masks_value'], 'fresh serial', j['e2e']['fresh_masks_serial_value'])
PY
