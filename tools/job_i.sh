for v in t128c9 t64c16 t192c5 t128c8; do
  L=$PWD/exp/ont_$v.so
  LVC_LIB_PATH=$L python tools/bench_configs.py --config 3 --distinct 4 > gpurun_out/i_$v.json 2> gpurun_out/i_$v.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/i_$v.json')); print('$v', d['batch_ms_p50'], d['kernel_avg_ms'])"
done
