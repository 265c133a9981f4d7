for v in ee eepf pf base; do
L=$PWD/exp/t5_$v.so; [ $v = base ] && L=$PWD/covid-spings-variant-caller_b200/lvc_b200/liblvc_b200.so
LVC_LIB_PATH=$L python bench.py --legs config5 --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/n_$v.json 2> gpurun_out/n_$v.err
python -c "
import json,sys; d=json.load(open('gpurun_out/n_$v.json')); r=d['roofline']; c=d['configs']['config5']; print('$v', d['ms_per_step'], r['avg_launch_ms'], r['frac_live'], c['deposit_kernel_ms'], c['frac_live'])"
done
