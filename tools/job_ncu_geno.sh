# ncu --set full of the genotype pass on config 3 (61 quality planes), with sources
C3="python tools/bench_configs.py --config 3 --distinct 2 --batches 8"
$C3 > gpurun_out/n_cfg3_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_genotype -s 8 -c 1 -o gpurun_out/prof_r2i_geno $C3 > gpurun_out/n_cfg3_ncu.log 2>&1
echo "rc=$?"; ls -la gpurun_out | tail -5
