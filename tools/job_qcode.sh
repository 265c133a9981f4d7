# quality-code batches + admitted-only batches on one B200: the GPU tests, then config 2 end to end in both transfer modes
python -m pytest tests -m gpu -x -q > gpurun_out/q_tests.log 2>&1; echo "rc=$?" >> gpurun_out/q_tests.log
tail -5 gpurun_out/q_tests.log
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 8"
LVC_INGEST_TIMING=1 $B --legs e2e_api > gpurun_out/q_codes.json 2> gpurun_out/q_codes.err; echo "codes rc=$?"
LVC_ZERO_COPY=0 $B --legs main > gpurun_out/q_codes_nozc.json 2> gpurun_out/q_codes_nozc.err; echo "codes nozc rc=$?"
