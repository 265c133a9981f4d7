# ncu --set full of the tiled deposit kernel on config 2 (default bench: admitted-only device batch, 2-bit quality codes) and config 5
M=lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum,lts__t_sectors_srcunit_tex_op_read.sum
C2="python bench.py --legs main --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$C2 > gpurun_out/f_cfg2_plain.log 2>&1 && ncu --set full --metrics $M --clock-control none --import-source on -k regex:k_deposit_tile5 -s 4 -c 1 -o gpurun_out/prof_r2f_cfg2 $C2 > gpurun_out/f_cfg2_ncu.log 2>&1
echo "cfg2 rc=$?"
