# genotype pass on config 3 (61 quality planes): 16-byte shared loads of the per-plane constants (LVC_GENO_BANK=0), constants and plane
# pointers as a kernel parameter (default), counts of the next batch in flight (LVC_GENO_PIPE)
C3="python tools/bench_configs.py --config 3 --distinct 2 --batches 40"
show() { tail -1 $1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['batch_ms_p50'], d['kernel_avg_ms'])"; }
for v in v128 v128p bank bankp bankb16; do LVC_LIB_PATH=$PWD/exp/lvc_$v.so $C3 > gpurun_out/g_$v.log 2>&1; echo $v; show gpurun_out/g_$v.log; done
LVC_GENO_BANK=0 LVC_LIB_PATH=$PWD/exp/lvc_bank.so $C3 > gpurun_out/g_bank0.log 2>&1; echo bank-off; show gpurun_out/g_bank0.log
LVC_LIB_PATH=$PWD/exp/lvc_bank.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q > gpurun_out/g_tests.log 2>&1; tail -2 gpurun_out/g_tests.log
LVC_LIB_PATH=$PWD/exp/lvc_bank.so python bench.py --legs main --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg2 step', d['ms_per_step'], d['roofline']['frac'])"
python bench.py --legs main --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg2 step (repo lib)', d['ms_per_step'], d['roofline']['frac'])"
