cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --mode train --steps 10 --warmup 3 > gpurun_out/r2_train_n8_final.json) 2> gpurun_out/r2_train_n8_final.err
echo "rc=$?"
python - <<'PY'
import json
for f in ('r2_train_n8_final',):
    try:
        j=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1]); print(f, j['value'], j['ms_per_step'], j['config'].get('cuda_graph'), j.get('allreduce'))
    except Exception as e: print(f, 'ERR', e)
PY
grep -v "Warn\|^$\|\*\*\*\|OMP_NUM\|Consider\|gpu_launches\|run_backward" gpurun_out/r2_train_n8_final.err | tail -4
