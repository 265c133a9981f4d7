python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/m_tests.log 2>&1; echo "rc=$?" >> gpurun_out/m_tests.log
for p in 1 0; do
LVC_TILE5_PERSIST=$p python bench.py --legs config5 --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/m_bench_p$p.json 2> gpurun_out/m_bench_p$p.err
done
tail -3 gpurun_out/m_tests.log
for f in gpurun_out/m_bench_p1.json gpurun_out/m_bench_p0.json; do python -c "
import json,sys; d=json.load(open('$f')); r=d['roofline']; c=d['configs']['config5']; print(d['ms_per_step'], r['avg_launch_ms'], r['frac_live'], c['deposit_kernel_ms'], c['frac_live'])"; done
