cd $GRAFT_REPO_ROOT
(timeout 600 python tools/msda_probe.py "g8:" "tma_c64:msda_tma=1" "tma_c32:msda_tma=1,msda_chunk=32" "tma_c128:msda_tma=1,msda_chunk=128" "tma_c96:msda_tma=1,msda_chunk=96" 2>&1 | tail -25) > gpurun_out/r2_d_probe.log
cat gpurun_out/r2_d_probe.log
