cd $GRAFT_REPO_ROOT
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/r2_k_smoke.log
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/r2_k_pytest.log
(timeout 900 python bench.py > gpurun_out/r2_k_bench.json) 2> gpurun_out/r2_k_bench.err
(timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_k_ref.json) 2> gpurun_out/r2_k_ref.err
cat gpurun_out/r2_k_smoke.log; tail -4 gpurun_out/r2_k_pytest.log; python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_k_bench.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['fresh_masks_value'], j['gpu_launches_per_step'], j['roofline']['frac'], j['roofline_gemm']['frac'], j['gpu_comparator']['value'], j['cpu_baseline'])
r=json.load(open('gpurun_out/r2_k_ref.json')); print(r['value'], r['cpu_baseline'], r['config'])
PY
tail -c 300 gpurun_out/r2_k_bench.err
