# ncu --set full of the long-read deposit kernel on config 3, with sources
C3="python tools/bench_configs.py --config 3 --distinct 2 --batches 8"
$C3 > gpurun_out/n_ont_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_deposit_ont -s 8 -c 1 -o gpurun_out/prof_r2i_ont $C3 > gpurun_out/n_ont_ncu.log 2>&1
echo "rc=$?"
