# the default build (LVC5_WAIT 1): GPU tests, then config 2 with the admitted-only and the keep-masked device batch
python -m pytest tests -m gpu -x -q > gpurun_out/w_tests.log 2>&1; echo "rc=$?" >> gpurun_out/w_tests.log
tail -4 gpurun_out/w_tests.log
B="python bench.py --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 3 --legs main"
$B --device-batch admitted > gpurun_out/w_adm.json 2> gpurun_out/w_adm.err; echo "adm rc=$?"
$B --device-batch masked > gpurun_out/w_msk.json 2> gpurun_out/w_msk.err; echo "msk rc=$?"
