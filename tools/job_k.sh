for v in x_nored x_noflush x_nop2; do
  LVC_LIB_PATH=$PWD/exp/$v.so python tools/bench_configs.py --config 3 --distinct 4 > gpurun_out/k_$v.json 2> gpurun_out/k_$v.err
  LVC_LIB_PATH=$PWD/exp/$v.so python tools/bench_configs.py --config 3 --distinct 4 --min-bq 10 > gpurun_out/k_${v}_bq10.json 2> gpurun_out/k_$v.err
done
for f in gpurun_out/k_*.json; do echo $f; python -c "
import json,sys; d=json.load(open('$f')); print(d['batch_ms_p50'], d['kernel_avg_ms'])"; done
