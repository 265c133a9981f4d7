# quick check of a kernel change: the tiled-kernel parity tests, then config 2 and config 5
python -m pytest tests/test_gpu_qcode.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/w_tests.log 2>&1; echo "rc=$?" >> gpurun_out/w_tests.log
tail -4 gpurun_out/w_tests.log
python bench.py --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 3 --legs config5 > gpurun_out/w_new.json 2> gpurun_out/w_new.err; echo "rc=$?"
