python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/l_tests2.log 2>&1; echo "rc=$?" >> gpurun_out/l_tests2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/l_bench2.json 2> gpurun_out/l_bench2.err; echo "bench rc=$?"
tail -5 gpurun_out/l_tests2.log; tail -c 3000 gpurun_out/l_bench2.json
