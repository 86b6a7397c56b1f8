cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/r2_g_tests.log
(timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/r2_g_bench.json) 2> gpurun_out/r2_g_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3_kernel -s 2 -c 1 -o gpurun_out/r2_gemm_f16x3_ffn1 python tools/one_gemm.py 22726 256 2048 4 > gpurun_out/r2_g_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3_kernel -s 2 -c 1 -o gpurun_out/r2_gemm_f16x3_ffn2 python tools/one_gemm.py 22726 2048 256 4 > gpurun_out/r2_g_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd_kernel -c 2 --profile-from-start off -o gpurun_out/r2_msda_fwd python tools/profile_step.py > gpurun_out/r2_g_ncu3.log 2>&1
tail -3 gpurun_out/r2_g_tests.log; python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_g_bench.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['fresh_masks_value'], j['gpu_launches_per_step'])
PY
tail -c 300 gpurun_out/r2_g_bench.err; ls -la gpurun_out/r2_*.ncu-rep
