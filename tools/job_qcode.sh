# quality-code batches + admitted-only batches on one B200: the new parity tests, then config 2 in both forms
python -m pytest tests/test_gpu_qcode.py -x -q > gpurun_out/q_tests.log 2>&1; echo "rc=$?" >> gpurun_out/q_tests.log
tail -5 gpurun_out/q_tests.log
B="python bench.py --legs main --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 8"
$B --quality-form bytes > gpurun_out/q_bytes.json 2> gpurun_out/q_bytes.err; echo "bytes rc=$?"
$B --quality-form codes > gpurun_out/q_codes.json 2> gpurun_out/q_codes.err; echo "codes rc=$?"
