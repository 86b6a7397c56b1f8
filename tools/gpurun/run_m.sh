cd $GRAFT_REPO_ROOT
for c in 0 6144 4096 9216; do
(SDETR_FFN_CHUNK=$c timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r2_m_bench_$c.json) 2> gpurun_out/r2_m_bench_$c.err
done
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "config2 or golden or runner or c256" 2>&1 | tail -4) > gpurun_out/r2_m_tests.log
python - <<'PY'
import json
for c in (0,6144,4096,9216):
    try:
        j=json.load(open(f'gpurun_out/r2_m_bench_{c}.json')); print(c, j['value'], j['ms_per_step'], j['e2e']['value'], j['gpu_launches_per_step'], j['roofline_gemm']['kernel_ms_per_step'])
    except Exception as e: print(c, 'ERR', e)
PY
tail -3 gpurun_out/r2_m_tests.log
