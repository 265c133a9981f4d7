# quick check of a change to the base-code path: its GPU tests, then the no-call benchmark
python -m pytest tests/test_gpu_qcode.py -m gpu -x -q > gpurun_out/b_new.log 2>&1; echo "rc=$?" >> gpurun_out/b_new.log; tail -2 gpurun_out/b_new.log
python tools/bench_nocalls.py > gpurun_out/b_nocalls.jsonl 2> gpurun_out/b_nocalls.err; echo "nocalls rc=$?"; cat gpurun_out/b_nocalls.jsonl; tail -3 gpurun_out/b_nocalls.err
