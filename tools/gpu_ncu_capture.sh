# ncu --set full captures of the deposit / genotype kernels on configs 2, 3, 5 (run: gpurun -- bash tools/gpu_ncu_capture.sh)
set -x
M=lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum,lts__t_sectors_srcunit_tex_op_read.sum
C2="python bench.py --legs main --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$C2 > gpurun_out/e_cfg2_plain.log 2>&1 && ncu --set full --metrics $M --clock-control none --import-source on -k regex:k_deposit_tile5 -s 4 -c 1 -o gpurun_out/prof_r2e_cfg2 $C2 > gpurun_out/e_cfg2_ncu.log 2>&1
C3="python tools/bench_configs.py --config 3 --distinct 2 --batches 8"
$C3 > gpurun_out/e_cfg3_plain.log 2>&1 && ncu --set full --metrics $M --clock-control none --import-source on -k regex:"k_deposit_ont|k_genotype" -s 8 -c 2 -o gpurun_out/prof_r2e_cfg3 $C3 > gpurun_out/e_cfg3_ncu.log 2>&1
C5="python tools/bench_configs.py --config 5 --steps 3"
$C5 > gpurun_out/e_cfg5_plain.log 2>&1 && ncu --set full --metrics $M --clock-control none --import-source on -k regex:k_deposit_tile5 -s 2 -c 1 -o gpurun_out/prof_r2e_cfg5 $C5 > gpurun_out/e_cfg5_ncu.log 2>&1
ls -la gpurun_out
