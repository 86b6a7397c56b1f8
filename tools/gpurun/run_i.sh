cd $GRAFT_REPO_ROOT
(timeout 200 python tools/gemm_trace2.py 22726 256 2048 | tail -24) > gpurun_out/r2_i_trace2.log 2>&1
cat gpurun_out/r2_i_trace2.log
