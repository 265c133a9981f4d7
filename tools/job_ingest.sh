# native ingest on the GPU box's host: thread sweep with the library's own inflate and with zlib's, phase times, then e2e_api
LVC_INGEST_TIMING=1 python tools/bench_ingest.py --pairs 996767 --sweep 8,16,32 --reps 3 > gpurun_out/i_sweep.jsonl 2> gpurun_out/i_sweep.err; echo "sweep rc=$?"
cat gpurun_out/i_sweep.jsonl
nproc; grep -m1 "model name" /proc/cpuinfo; cat /sys/kernel/mm/transparent_hugepage/enabled
python bench.py --legs e2e_api --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/i_api.json 2> gpurun_out/i_api.err; echo "api rc=$?"
python -c "import json; d=json.load(open('gpurun_out/i_api.json')); print(json.dumps(d['e2e_api'])); print(d['e2e'])"
