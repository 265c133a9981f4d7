# full GPU validation of a commit on one B200: pytest -m gpu, bench.py (all legs), the reference arm, ncu launch lists (run: gpurun -- bash tools/gpu_validate.sh)
python -m pytest tests -m gpu -x -q > gpurun_out/o_tests.log 2>&1; echo "rc=$?" >> gpurun_out/o_tests.log
python bench.py > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/o_ref.json 2> gpurun_out/o_ref.err; echo "ref rc=$?"
C="python bench.py --legs main --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$C > gpurun_out/o_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $C > gpurun_out/o_ncu.log 2>&1
C3="python tools/bench_configs.py --config 3 --distinct 2 --batches 10"
$C3 > gpurun_out/o_plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg3.csv $C3 > gpurun_out/o_ncu3.log 2>&1
tail -4 gpurun_out/o_tests.log
