cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --mode train --steps 10 --warmup 3 > gpurun_out/r2_train_n2_final.json) 2> gpurun_out/r2_train_n2_final.err
echo "rc=$?"
python - <<'PY'
import json
for f in ('r2_train_n2_final',):
    try:
        j=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1]); print(f, j['value'], j['ms_per_step'], j['config'].get('cuda_graph'), j.get('allreduce'), j['gpu_launches_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
grep -v "Warn\|^$\|\*\*\*\|OMP_NUM\|Consider\|gpu_launches" gpurun_out/r2_train_n2_final.err | tail -5
