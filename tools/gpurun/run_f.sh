cd $GRAFT_REPO_ROOT
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "select or config2 or five_scale or golden_even or ragged" 2>&1 | tail -8) > gpurun_out/r2_f_tests.log
(timeout 600 python bench.py --skip-cpu-baseline --pipeline-depth 2 > gpurun_out/r2_f_bench_d2.json) 2> gpurun_out/r2_f_bench_d2.err
(timeout 600 python bench.py --skip-cpu-baseline --pipeline-depth 3 > gpurun_out/r2_f_bench_d3.json) 2> gpurun_out/r2_f_bench_d3.err
tail -3 gpurun_out/r2_f_tests.log; python - <<'PY'
import json
for d in (2,3):
    try:
        j=json.load(open(f'gpurun_out/r2_f_bench_d{d}.json')); print(d, j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['fresh_masks_value'], j['gpu_launches_per_step'])
    except Exception as e: print(d, 'ERR', e)
PY
tail -c 300 gpurun_out/r2_f_bench_d2.err
