python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/j_tests.log
python tools/bench_configs.py --config 3 --distinct 4 > gpurun_out/j_cfg3.json 2> gpurun_out/j_cfg3.err
LVC_GENO_LPP=1 python tools/bench_configs.py --config 3 --distinct 4 > gpurun_out/j_cfg3_lpp1.json 2>&1
tail -3 gpurun_out/j_tests.log
for f in gpurun_out/j_cfg3.json gpurun_out/j_cfg3_lpp1.json; do python -c "
import json,sys; d=json.load(open('$f')); print(d['batch_ms_p50'], d['kernel_avg_ms'])"; done
