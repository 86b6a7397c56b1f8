#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3) > gpurun_out/r2_o_pytest.log; tail -2 gpurun_out/r2_o_pytest.log
(timeout 300 python bench.py --skip-cpu-baseline --steps 20 --warmup 3 > gpurun_out/r2_o_bench.json) 2> gpurun_out/r2_o_bench.err
echo "rc=$? lines=$(wc -l < gpurun_out/r2_o_bench.json)"; python -c "
import json; j=json.load(open('gpurun_out/r2_o_bench.json')); print(j['value'], j['e2e']['value'], j['e2e']['fresh_masks_value'])"
