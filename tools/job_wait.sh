# where the tiled kernel waits for the previous kernel (LVC5_WAIT 0 / 1 / 2): parity of the default build, then config 2 / 5 / 4 per variant
python -m pytest tests -m gpu -x -q > gpurun_out/w_tests.log 2>&1; echo "rc=$?" >> gpurun_out/w_tests.log
tail -4 gpurun_out/w_tests.log
B="python bench.py --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 3 --legs config5,config4"
$B > gpurun_out/w2.json 2> gpurun_out/w2.err; echo "w2 rc=$?"
LVC_LIB_PATH=$PWD/exp/lvc_w1.so $B > gpurun_out/w1.json 2> gpurun_out/w1.err; echo "w1 rc=$?"
LVC_LIB_PATH=$PWD/exp/lvc_w0.so $B > gpurun_out/w0.json 2> gpurun_out/w0.err; echo "w0 rc=$?"
