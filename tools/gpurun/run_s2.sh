cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_scale_n2_final.json) 2> gpurun_out/r2_scale_n2_final.err
echo "rc=$?"
python - <<'PY'
import json
lines=open('gpurun_out/r2_scale_n2_final.json').read().splitlines()
print('stdout lines:', len(lines))
j=json.loads(lines[-1]); print(j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], 'fresh', j['e2e']['fresh_masks_value'], j['clocks'])
PY
grep -v "Warn\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r2_scale_n2_final.err | tail -4
