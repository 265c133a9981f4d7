python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/h_tests.log 2>&1; echo "rc=$?" >> gpurun_out/h_tests.log
python tools/bench_configs.py --config 3 --distinct 4 > gpurun_out/h_cfg3.json 2> gpurun_out/h_cfg3.err
python tools/bench_configs.py --config 3 --distinct 4 --min-bq 10 > gpurun_out/h_cfg3_bq10.json 2>&1
M=lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum,lts__t_sectors_srcunit_tex_op_read.sum
C3="python tools/bench_configs.py --config 3 --distinct 2 --batches 8"
$C3 > gpurun_out/h_cfg3_plain.log 2>&1 && ncu --set full --metrics $M --clock-control none --import-source on -k regex:"k_deposit_ont" -s 4 -c 1 -o gpurun_out/prof_r2h_cfg3 $C3 > gpurun_out/h_cfg3_ncu.log 2>&1
tail -3 gpurun_out/h_tests.log; cat gpurun_out/h_cfg3.json gpurun_out/h_cfg3_bq10.json
