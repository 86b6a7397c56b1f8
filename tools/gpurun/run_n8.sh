cd $GRAFT_REPO_ROOT
nvidia-smi topo -m > gpurun_out/r2_n8_topo.txt 2>&1
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_scale_n8.json) 2> gpurun_out/r2_scale_n8.err
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --mode train --steps 6 --warmup 3 > gpurun_out/r2_train_n8.json) 2> gpurun_out/r2_train_n8.err
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --mode train --steps 6 --warmup 3 > gpurun_out/r2_train_n2.json) 2> gpurun_out/r2_train_n2.err
python - <<'PY'
import json
for f in ('r2_scale_n8','r2_train_n8','r2_train_n2'):
    try:
        j=json.load(open(f'gpurun_out/{f}.json')); print(f, j['value'], j['ms_per_step'], j.get('e2e',{}).get('value'), j.get('allreduce'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -c 600 gpurun_out/r2_scale_n8.err; tail -c 400 gpurun_out/r2_train_n8.err
