# config-5 read-chunk sharding at N GPUs (run: gpurun --gpus N -- bash tools/gpu_sharded_bench.sh N)
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --legs config5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/q_bench$N.json 2> gpurun_out/q_bench$N.err; echo "bench rc=$?"
python - <<PY
import json
txt=open('gpurun_out/q_bench$N.json').read()
line=[l for l in txt.split('\n') if l.startswith('{')][-1]
d=json.loads(line); c=d['configs']['config5_sharded']
for m in ('reduce_scatter','all_reduce','halo','peer'): print(m, c.get(m))
print('main', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
tail -3 gpurun_out/q_bench$N.err
