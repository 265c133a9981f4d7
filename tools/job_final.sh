# validation of the round's final commit on one B200: pytest -m gpu, bench.py (all legs), the reference arm
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/o_tests.log 2>&1; echo "rc=$?" >> gpurun_out/o_tests.log
( time LVC_INGEST_TIMING=1 python bench.py > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err ) 2> gpurun_out/o_bench.time; echo "bench rc=$?"
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/o_ref.json 2> gpurun_out/o_ref.err ) 2> gpurun_out/o_ref.time; echo "ref rc=$?"
tail -4 gpurun_out/o_tests.log; cat gpurun_out/o_bench.time gpurun_out/o_ref.time | grep real
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/o_smoke.log 2>&1; tail -1 gpurun_out/o_smoke.log
