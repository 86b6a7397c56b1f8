#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "linear_train or training_path" 2>&1 | tail -8
for c in 1 0; do
(SDETR_TRAIN_TENSOR_CORE=$c timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r2_o_train_$c.json) 2> gpurun_out/r2_o_train_$c.err
done
python - <<'PY'
import json
for c in (1, 0):
    try:
        j=json.loads([l for l in open(f'gpurun_out/r2_o_train_{c}.json') if l.startswith('{')][-1]); print(c, j['value'], j['ms_per_step'], j['loss'], j['gpu_launches_per_step'])
    except Exception as e: print(c, 'ERR', e)
PY
tail -3 gpurun_out/r2_o_train_1.err
