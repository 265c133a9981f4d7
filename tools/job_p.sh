python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/p_tests2.log 2>&1; echo "rc=$?" >> gpurun_out/p_tests2.log
tail -5 gpurun_out/p_tests2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --legs config5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/p_bench2.json 2> gpurun_out/p_bench2.err; echo "bench rc=$?"
python - <<'PY'
import json
txt=open('gpurun_out/p_bench2.json').read()
line=[l for l in txt.split('\n') if l.startswith('{')][-1]
d=json.loads(line); c=d['configs']['config5_sharded']
for m in ('reduce_scatter','all_reduce','halo','peer'): print(m, c.get(m))
print('main', d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac_live'])
PY
tail -3 gpurun_out/p_bench2.err
