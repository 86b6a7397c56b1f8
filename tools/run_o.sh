#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in 0 140 132 120; do
(timeout 600 python bench.py --skip-cpu-baseline --e2e-persistent-ctas $c > gpurun_out/r2_o_bench_c$c.json) 2> gpurun_out/r2_o_bench_c$c.err
done
python - <<'PY'
import json
for c in (0, 140, 132, 120):
    try:
        j=json.load(open(f'gpurun_out/r2_o_bench_c{c}.json')); print(c, j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['serial_value'], j['e2e']['fresh_masks_value'])
    except Exception as e: print(c, 'ERR', e)
PY
