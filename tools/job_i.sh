python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
python tools/bench_configs.py --config 3 --distinct 4 > gpurun_out/i_base.json 2> gpurun_out/i_base.err
python -c "
import json,sys; d=json.load(open('gpurun_out/i_base.json')); print(d['batch_ms_p50'], d['kernel_avg_ms'])"
