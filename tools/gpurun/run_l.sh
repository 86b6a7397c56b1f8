cd $GRAFT_REPO_ROOT
(SDETR_POISON=ff timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/r2_l_poison.log
(timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "mask_plan or salience_targets or variants_agree or non_finite or topk_desc" 2>&1 | tail -15) > gpurun_out/r2_l_memcheck.log
(timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_decoder_half.py -q -x -k "nms or neck" 2>&1 | tail -15) >> gpurun_out/r2_l_memcheck.log
tail -5 gpurun_out/r2_l_poison.log; tail -30 gpurun_out/r2_l_memcheck.log
