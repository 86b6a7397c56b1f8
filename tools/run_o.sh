#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ffn_fused" 2>&1 | tail -3
timeout 300 python tools/bench_ffn.py 2>&1 | tee gpurun_out/r2_n_bench_ffn_v6.txt
for d in 3 4; do
(timeout 600 python bench.py --skip-cpu-baseline --pipeline-depth $d > gpurun_out/r2_o_bench_d$d.json) 2> gpurun_out/r2_o_bench_d$d.err
done
python - <<'PY'
import json
for c in (3, 4):
    try:
        j=json.load(open(f'gpurun_out/r2_o_bench_d{c}.json')); print(c, j['value'], j['ms_per_step'], j['e2e']['value'], j['gpu_launches_per_step'], j['roofline_gemm']['kernel_ms_per_step'], j['roofline_gemm']['frac'], j['clocks'])
    except Exception as e: print(c, 'ERR', e)
PY
