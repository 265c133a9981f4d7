for v in ont_c5 ont_pf ont_pfc5; do
  LVC_LIB_PATH=$PWD/exp/$v.so python tools/bench_configs.py --config 3 --distinct 4 > gpurun_out/i_$v.json 2> gpurun_out/i_$v.err
done
python tools/bench_configs.py --config 3 --distinct 4 > gpurun_out/i_base.json 2> gpurun_out/i_base.err
for f in gpurun_out/i_*.json; do echo $f; python -c "
import json,sys; d=json.load(open('$f')); print(d['batch_ms_p50'], d['kernel_avg_ms'])"; done
