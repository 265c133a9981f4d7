# genotype pass on config 3 (61 quality planes): plane counts requested together per lane (LVC_GENO_BATCH 4 / 8 / 16)
C3="python tools/bench_configs.py --config 3 --distinct 2 --batches 40"
$C3 > gpurun_out/g_b8.log 2>&1; tail -1 gpurun_out/g_b8.log
LVC_LIB_PATH=$PWD/exp/lvc_gb16.so $C3 > gpurun_out/g_b16.log 2>&1; tail -1 gpurun_out/g_b16.log
LVC_LIB_PATH=$PWD/exp/lvc_gb4.so $C3 > gpurun_out/g_b4.log 2>&1; tail -1 gpurun_out/g_b4.log
