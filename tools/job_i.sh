python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for v in base g16; do
  L=$PWD/exp/$v.so; [ $v = base ] && L=$PWD/covid-spings-variant-caller_b200/lvc_b200/liblvc_b200.so
  LVC_LIB_PATH=$L python tools/bench_configs.py --config 3 --distinct 4 > gpurun_out/i_$v.json 2> gpurun_out/i_$v.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/i_$v.json')); print('$v', d['batch_ms_p50'], d['kernel_avg_ms'])"
  LVC_LIB_PATH=$L python bench.py --legs main --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v cfg2', d['ms_per_step'], d['roofline']['other_kernels_ms_per_step'])"
done
