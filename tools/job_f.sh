python -m pytest tests -m gpu -x -q > gpurun_out/f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/f_tests.log
for lpp in 1 2 4 8; do LVC_GENO_LPP=$lpp python tools/bench_configs.py --config 3 --distinct 4 > gpurun_out/f_cfg3_lpp$lpp.json 2> gpurun_out/f_cfg3_lpp$lpp.err; done
LVC_LONG_IMPL=3 python tools/bench_configs.py --config 3 --distinct 4 > gpurun_out/f_cfg3_impl3.json 2>&1
python tools/bench_configs.py --config 3 --distinct 4 --min-bq 10 > gpurun_out/f_cfg3_bq10.json 2>&1
tail -3 gpurun_out/f_tests.log; cat gpurun_out/f_cfg3_*.json
