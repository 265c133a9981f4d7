# smoke() and the ncu launch lists of the round's final commit
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/o_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/o_smoke.log
C="python bench.py --legs main --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$C > gpurun_out/o_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $C > gpurun_out/o_ncu.log 2>&1; echo "ncu2 rc=$?"
C3="python tools/bench_configs.py --config 3 --distinct 2 --batches 10"
$C3 > gpurun_out/o_plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg3.csv $C3 > gpurun_out/o_ncu3.log 2>&1; echo "ncu3 rc=$?"
