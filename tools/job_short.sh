python -m pytest tests/test_gpu_qcode.py -m gpu -x -q -k "base_code or ships_base" > gpurun_out/b_new.log 2>&1; echo "rc=$?" >> gpurun_out/b_new.log; tail -2 gpurun_out/b_new.log
( time LVC_INGEST_TIMING=1 python bench.py > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err ) 2> gpurun_out/o_bench.time; echo "bench rc=$?"; grep real gpurun_out/o_bench.time
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/o_smoke.log 2>&1; tail -1 gpurun_out/o_smoke.log
